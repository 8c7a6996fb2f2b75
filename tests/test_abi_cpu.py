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
