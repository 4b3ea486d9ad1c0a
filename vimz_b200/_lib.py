"""ctypes binding of libvimz_gpu.so (the C ABI in include/vimz_gpu.h).

This is the same binding a reference-side FFI stub would make (see INTEGRATION.md for the Rust
`extern "C"` block); Python is used here only because the image has no Rust toolchain.  There is
no fallback: if the shared library is missing the import fails, and if no CUDA device is visible
`vimz_ctx_create` returns VIMZ_ERR_NO_DEVICE which is raised as VimzError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VIMZ_GPU_LIB") or os.path.join(_HERE, "libvimz_gpu.so")  # override only for A/B experiments

VIMZ_OK = 0
VIMZ_ERR_CUDA = -1
VIMZ_ERR_ARG = -2
VIMZ_ERR_LENGTH = -3
VIMZ_ERR_NO_DEVICE = -4
VIMZ_ERR_INDEX = -5

CURVE_IDS = {"pallas": 0, "vesta": 1, "bn254": 2, "grumpkin": 3}


class VimzError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[vimz_gpu {code}] {msg}")
        self.code = code


class InvalidWitnessLength(VimzError):
    """nova-snark's NovaError::InvalidWitnessLength."""


class InvalidIndex(VimzError):
    """nova-snark's NovaError::InvalidIndex."""


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "vimz_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    pp = C.POINTER(C.c_void_p)
    sigs = {
        "vimz_last_error": (C.c_char_p, []),
        "vimz_version": (i32, []),
        "vimz_device_count": (i32, []),
        "vimz_host_alloc": (vp, [sz]),
        "vimz_host_free": (None, [vp]),
        "vimz_ctx_create": (i32, [i32, i32, pp]),
        "vimz_ctx_destroy": (None, [vp]),
        "vimz_ctx_sync": (i32, [vp]),
        "vimz_ctx_set_option": (i32, [vp, C.c_char_p, C.c_long]),
        "vimz_ctx_profile": (i32, [vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64), i32]),
        "vimz_ctx_stream": (vp, [vp]),
        "vimz_ctx_launch_count": (u64, [vp]),
        "vimz_ck_upload": (i32, [vp, vp, sz, pp]),
        "vimz_ck_upload_dev": (i32, [vp, vp, sz, pp]),
        "vimz_ck_destroy": (None, [vp]),
        "vimz_ck_len": (sz, [vp]),
        "vimz_ck_window_bits": (i32, [vp]),
        "vimz_ck_num_windows": (i32, [vp]),
        "vimz_msm": (i32, [vp, vp, vp, sz, vp]),
        "vimz_msm_dev": (i32, [vp, vp, vp, sz, vp]),
        "vimz_msm_range_dev": (i32, [vp, vp, sz, vp, sz, vp]),
        "vimz_msm_async_dev": (i32, [vp, vp, sz, vp, sz, vp]),
        "vimz_point_sum": (i32, [vp, vp, sz, vp]),
        "vimz_point_to_affine": (i32, [vp, vp, vp]),
        "vimz_point_scale_add": (i32, [vp, vp, vp, vp, vp]),
        "vimz_shape_upload": (i32, [vp, sz, sz, sz, vp, vp, vp, sz, vp, vp, vp, sz, vp, vp, vp, sz, pp]),
        "vimz_shape_destroy": (None, [vp]),
        "vimz_multiply_vec": (i32, [vp, vp, vp, sz, vp, vp, vp]),
        "vimz_commit_T": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "vimz_fold_witness": (i32, [vp, vp, vp, vp, sz, vp, vp, sz, vp, vp]),
        "vimz_acc_init": (i32, [vp, vp, vp, pp]),
        "vimz_acc_init_sharded": (i32, [vp, vp, vp, vp, sz, sz, pp]),
        "vimz_acc_step_begin_dev_async": (i32, [vp, vp, vp, pp]),
        "vimz_acc_step_combine_dev": (i32, [vp, vp, sz, vp, vp]),
        "vimz_acc_reset": (i32, [vp]),
        "vimz_acc_load": (i32, [vp, vp, vp, vp, vp, vp, vp]),
        "vimz_acc_step_begin": (i32, [vp, vp, vp, vp, vp]),
        "vimz_acc_step_begin_dev": (i32, [vp, vp, vp, vp, vp]),
        "vimz_acc_commit_fresh": (i32, [vp, vp, vp, vp]),
        "vimz_acc_cross_begin": (i32, [vp, vp]),
        "vimz_acc_fresh_witness": (i32, [vp, vp, vp]),
        "vimz_acc_stage_fresh": (i32, [vp, vp, sz, sz]),
        "vimz_acc_step_begin_async": (i32, [vp, vp, vp]),
        "vimz_acc_step_wait": (i32, [vp, vp, vp]),
        "vimz_acc_step_begin_staged": (i32, [vp, vp, sz, sz, vp, vp, vp]),
        "vimz_acc_step_end": (i32, [vp, vp]),
        "vimz_acc_download": (i32, [vp, vp, vp, vp, vp, vp, vp]),
        "vimz_acc_last_T": (i32, [vp, vp]),
        "vimz_acc_destroy": (None, [vp]),
        "vimz_comm_unique_id": (i32, [vp]),
        "vimz_comm_create": (i32, [i32, vp, i32, i32, pp]),
        "vimz_comm_destroy": (None, [vp]),
        "vimz_comm_rank": (i32, [vp]),
        "vimz_comm_world": (i32, [vp]),
        "vimz_comm_nccl_version": (i32, []),
        "vimz_comm_broadcast_dev": (i32, [vp, vp, vp, sz, i32]),
        "vimz_msm_sharded_dev": (i32, [vp, vp, vp, sz, vp, sz, vp]),
        "vimz_acc_step_begin_sharded_dev": (i32, [vp, vp, vp, vp, vp, vp]),
        "vimz_acc_step_begin_sharded": (i32, [vp, vp, vp, i32, vp, vp, vp]),
        "vimz_gen_bases_dev": (i32, [vp, u64, u64, sz, vp]),
        "vimz_field_op": (i32, [vp, i32, i32, vp, vp, sz, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    lib._vimz_symbols = tuple(sigs)
    return lib


lib = _load()
EXPORTED_SYMBOLS = lib._vimz_symbols


def check(rc: int) -> None:
    if rc == VIMZ_OK:
        return
    msg = (lib.vimz_last_error() or b"").decode("utf-8", "replace")
    if rc == VIMZ_ERR_LENGTH:
        raise InvalidWitnessLength(rc, msg)
    if rc == VIMZ_ERR_INDEX:
        raise InvalidIndex(rc, msg)
    raise VimzError(rc, msg)
