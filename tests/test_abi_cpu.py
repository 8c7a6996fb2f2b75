"""The C-ABI shared library loads without a GPU and exports every symbol include/clipcap_b200.h declares; the ctypes
mirrors of its structs have the C layout; error paths that need no device work."""
import ctypes as C
import os
import re

from clipcap_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "clipcap_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_ffi.LIB_PATH), "run __graft_entry__.build() first"
    handle = C.CDLL(_ffi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/clipcap_b200.h but not exported"
    assert set(names) == set(_ffi.PROTOTYPES), set(names) ^ set(_ffi.PROTOTYPES)


def test_version_and_struct_layout():
    assert "sm_100a" in _ffi.version()
    assert C.sizeof(_ffi.cc_tensor) == 8 + 8 + 4 + 4 + 32
    assert C.sizeof(_ffi.cc_vit_cfg) == 32
    assert C.sizeof(_ffi.cc_mapper_cfg) == 40
    assert C.sizeof(_ffi.cc_gpt2_cfg) == 24
    assert C.sizeof(_ffi.cc_gen_cfg) == 64  # 11 x 4-byte fields, padding, history pointer, 64-bit seed
    assert _ffi.cc_gen_cfg.history.offset == 48 and _ffi.cc_gen_cfg.seed.offset == 56


def test_null_arguments_are_rejected_without_a_device():
    lib = _ffi.lib()
    h = C.c_void_p()
    st = lib.cc_mapper_create(C.byref(h), None, None, 0, 1)
    assert st == -1 and "null" in _ffi.last_error()
    assert lib.cc_generate(None, None, 0, 1, 1, None, None, None, None, None) == -1
    assert lib.cc_vit_forward(None, None, 0, 1, 0, None, 0, None) == -1


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        return
    lib = _ffi.lib()
    x = (C.c_float * 8)()
    st = lib.cc_op_layernorm(x, 8, x, x, x, 8, 1, 8, 1e-5, None)
    assert st in (-4, -6), _ffi.last_error()  # CC_ECUDA / CC_EARCH


def test_struct_layouts_match_the_header_compiled_as_c(tmp_path):
    """include/clipcap_b200.h is plain C: gcc compiles it, and the sizes and field offsets it gives every struct are the
    ones the ctypes mirrors in clipcap_b200/_ffi.py have (so a cgo / JNI / ctypes binding sees the same layout)."""
    import shutil
    import subprocess
    import pytest
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    structs = {n: getattr(_ffi, n) for n in ("cc_tensor", "cc_vit_cfg", "cc_clap_cfg", "cc_mapper_cfg", "cc_gpt2_cfg",
                                             "cc_gen_cfg")}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "clipcap_b200.h"', 'int main(void) {']
    for name, st in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in st._fields_:
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True,
                                                       text=True).stdout.splitlines())
    for name, st in structs.items():
        assert int(got[name]) == C.sizeof(st), name
        for field, _ in st._fields_:
            assert int(got[f"{name}.{field}"]) == getattr(st, field).offset, f"{name}.{field}"
