"""ctypes binding of libclipcap_b200.so (the C ABI declared in include/clipcap_b200.h).

PyTorch is used only for device memory and streams: tensors cross the boundary as ``data_ptr()`` + sizes.
There is no CPU fallback: if the shared library is missing, or the device is not sm_100, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Sequence, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclipcap_b200.so")

CC_OK = 0
CC_F32, CC_F16 = 0, 1
STATUS_NAMES = {0: "CC_OK", -1: "CC_EINVAL", -2: "CC_ESHAPE", -3: "CC_EALIGN", -4: "CC_ECUDA", -5: "CC_ENCCL",
                -6: "CC_EARCH", -7: "CC_ENOMEM"}

# GEMM epilogues (csrc/common.h `enum Epi`)
EPI_F16_NONE, EPI_F16_RELU, EPI_F16_QUICKGELU, EPI_F16_GELU_NEW, EPI_F16_TANH, EPI_F32, EPI_RESID_F32, EPI_ARGMAX = range(8)
EPI_F16_GELU_ERF = 10  # after the two internal plan kinds (split-K partials, head-major QKV)

CC_MAPPER_TRANSFORMER, CC_MAPPER_WINDOWED, CC_MAPPER_MLP = 0, 1, 2
CC_GEN_GREEDY, CC_GEN_BEAM, CC_GEN_NUCLEUS, CC_GEN_SAMPLE = 0, 1, 2, 3


class CCError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class cc_tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 4)]


class cc_vit_cfg(C.Structure):
    _fields_ = [("image_size", C.c_int32), ("patch", C.c_int32), ("width", C.c_int32), ("layers", C.c_int32),
                ("heads", C.c_int32), ("mlp_dim", C.c_int32), ("out_dim", C.c_int32), ("eps", C.c_float)]


class cc_clap_cfg(C.Structure):
    _fields_ = [("num_mel_bins", C.c_int32), ("spec_size", C.c_int32), ("patch", C.c_int32), ("embed", C.c_int32),
                ("depths", C.c_int32 * 4), ("heads", C.c_int32 * 4), ("window", C.c_int32),
                ("projection_dim", C.c_int32), ("eps", C.c_float)]


class cc_mapper_cfg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("E", C.c_int32), ("d", C.c_int32), ("P", C.c_int32), ("K", C.c_int32),
                ("H", C.c_int32), ("L", C.c_int32), ("W", C.c_int32), ("use_pos", C.c_int32), ("eps", C.c_float)]


class cc_gpt2_cfg(C.Structure):
    _fields_ = [("d", C.c_int32), ("L", C.c_int32), ("H", C.c_int32), ("V", C.c_int32), ("n_pos", C.c_int32),
                ("eps", C.c_float)]


class cc_gen_cfg(C.Structure):
    _fields_ = [("mode", C.c_int32), ("beam", C.c_int32), ("entry_length", C.c_int32), ("temperature", C.c_float),
                ("stop_token", C.c_int32), ("top_p", C.c_float), ("top_k", C.c_int32), ("repetition_penalty", C.c_float),
                ("desired_sentence_length", C.c_int32), ("sentence_length_factor", C.c_float), ("n_history", C.c_int32),
                ("history", C.c_void_p), ("seed", C.c_uint64)]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/clipcap_b200.h declares.
PROTOTYPES: Dict[str, Tuple[object, List[object]]] = {
    "cc_last_error": (C.c_char_p, []),
    "cc_version": (C.c_char_p, []),
    "cc_vit_create": (_i, [_pp, C.POINTER(cc_vit_cfg), C.POINTER(cc_tensor), _i, _i]),
    "cc_vit_forward": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "cc_vit_last_launches": (_i, [_vp]),
    "cc_vit_destroy": (None, [_vp]),
    "cc_clap_create": (_i, [_pp, C.POINTER(cc_clap_cfg), C.POINTER(cc_tensor), _i, _i]),
    "cc_clap_forward": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "cc_clap_last_launches": (_i, [_vp]),
    "cc_clap_destroy": (None, [_vp]),
    "cc_mapper_create": (_i, [_pp, C.POINTER(cc_mapper_cfg), C.POINTER(cc_tensor), _i, _i]),
    "cc_mapper_forward": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp]),
    "cc_mapper_last_launches": (_i, [_vp]),
    "cc_mapper_destroy": (None, [_vp]),
    "cc_gpt2_create": (_i, [_pp, C.POINTER(cc_gpt2_cfg), C.POINTER(cc_tensor), _i, _i, _i]),
    "cc_gpt2_logits": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "cc_gpt2_embed": (_i, [_vp, _vp, _i, _vp, _i, _vp]),
    "cc_generate": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(cc_gen_cfg), _vp, _vp, _vp, _vp]),
    "cc_generate_prefill": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(cc_gen_cfg), _vp]),
    "cc_generate_decode": (_i, [_vp, _i, _i, C.POINTER(cc_gen_cfg), _vp, _vp, _vp, _vp]),
    "cc_gpt2_last_launches": (_i, [_vp]),
    "cc_gpt2_set_prefill_defer": (_i, [_vp, _i]),
    "cc_gpt2_destroy": (None, [_vp]),
    "cc_comm_unique_id": (_i, [_vp]),
    "cc_comm_create": (_i, [_pp, _vp, _i, _i]),
    "cc_allgather_prefix": (_i, [_vp, _vp, C.c_size_t, _vp]),
    "cc_comm_rank": (_i, [_vp]),
    "cc_comm_nranks": (_i, [_vp]),
    "cc_nccl_version": (_i, []),
    "cc_comm_destroy": (None, [_vp]),
    "cc_partition_create": (_i, [_pp, _i, _i]),
    "cc_partition_stream": (_vp, [_vp, _i]),
    "cc_partition_sms": (_i, [_vp, _i]),
    "cc_partition_destroy": (None, [_vp]),
    "cc_set_sm_budget": (None, [_i]),
    "cc_get_sm_budget": (_i, []),
    "cc_prof_enable": (None, [_i]),
    "cc_prof_read": (None, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "cc_prof_read_family": (None, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "cc_op_gemm": (_i, [_vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "cc_op_layernorm": (_i, [_vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _f, _vp]),
    "cc_op_attention": (_i, [_vp, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _f, _vp]),
    "cc_op_skinny_gemm": (_i, [_vp, _vp, _vp, _f, _vp, _i64, _i, _vp, _i, _i, _i, _vp, _vp, _i64, _vp]),
    "cc_op_attention_bwd": (_i, [_vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _f, _vp]),
    "cc_op_decode_attention": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "cc_op_decode_attention_beam": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "cc_op_tile_image": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "cc_op_sample": (_i, [_vp, _i, _i, C.POINTER(cc_gen_cfg), _i, _vp, _vp, _vp, _vp]),
    "cc_train_create": (_i, [_pp, C.POINTER(cc_mapper_cfg), C.POINTER(cc_gpt2_cfg), C.POINTER(cc_tensor), _i, _i, _i]),
    "cc_train_step": (_i, [_vp, C.POINTER(cc_tensor), _i, C.POINTER(cc_tensor), _i, _vp, _i, _vp, _i, _i, _f, _vp, _vp]),
    "cc_train_last_nonfinite": (_i, [_vp, _vp]),
    "cc_train_last_launches": (_i, [_vp]),
    "cc_train_destroy": (None, [_vp]),
    "cc_op_adamw": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Loads the shared library once. Raises (never falls back) when it is missing or a symbol is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(clipcap_b200 has no CPU / PyTorch fallback path)")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(handle, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(status: int) -> None:
    if status != CC_OK:
        raise CCError(status, lib().cc_last_error().decode("utf-8", "replace"))


def last_error() -> str:
    return lib().cc_last_error().decode("utf-8", "replace")


def version() -> str:
    return lib().cc_version().decode()


def torch_dtype_code(t) -> int:
    import torch
    if t.dtype == torch.float32:
        return CC_F32
    if t.dtype == torch.float16:
        return CC_F16
    raise TypeError(f"clipcap_b200 accepts float32 / float16 tensors at the boundary, got {t.dtype}")


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def make_tensor_table(named: Sequence[Tuple[str, "object"]]):
    """[(name, fp32 contiguous torch tensor)] -> (cc_tensor array, keep-alive list)."""
    import torch
    arr = (cc_tensor * len(named))()
    keep = []
    for i, (name, t) in enumerate(named):
        t = t.detach()
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(torch.float32).contiguous()
        if t.dim() > 4:
            t = t.reshape(t.shape[0], -1)
        bname = name.encode()
        keep.append((bname, t))
        arr[i].name = bname
        arr[i].data = t.data_ptr()
        arr[i].dtype = CC_F32
        arr[i].ndim = max(1, t.dim())
        shape = list(t.shape) if t.dim() > 0 else [1]
        for k in range(4):
            arr[i].shape[k] = shape[k] if k < len(shape) else 1
    return arr, keep
