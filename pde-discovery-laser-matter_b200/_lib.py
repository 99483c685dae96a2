"""ctypes binding of libpdegram.so (include/pdegram.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is visible,
every compute entry point raises.  PyTorch is used only as the device allocator / stream
provider; no torch type crosses the C ABI (plain pointers and sizes).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("PG_LIBPDEGRAM") or PKG_DIR / "libpdegram.so")   # env: experiment builds (tools/)
CSRC_DIR = PKG_DIR / "csrc"

# ---- constants mirrored from include/pdegram.h
PG_MAX_P = 16
PG_MAX_FOLDS = 8
PG_COMM_MAX_RANKS, PG_COMM_MAX_LEN = 16, 1536
FD_KS_PERIODIC, FD_BASIC_TRIM, FD_SLICE_CENTRAL = 0, 1, 2
LIB_KS_TRUE, LIB_KS_TRUE_ADV, LIB_KS_RICH, LIB_KS_RICH_NOADV, LIB_BASIC = 0, 1, 2, 3, 4
LIB_KS_GRAD, LIB_KS_LAP, LIB_PATCH_MODEL4, LIB_PATCH_FULL, LIB_PATCH_DERIVS, LIB_AR_FULL = 5, 6, 7, 8, 9, 10
STRIDGE_KS, STRIDGE_SKLEARN, STRIDGE_BASIC = 0, 1, 2
STRIDGE_RMS_PRESCALE, STRIDGE_NO_EPS = 1, 2
VARIANT_AUTO, VARIANT_GENERIC, VARIANT_TILED = 0, 1, 2
LIB_WIDTH = {LIB_KS_TRUE: 3, LIB_KS_TRUE_ADV: 5, LIB_KS_RICH: 9, LIB_KS_RICH_NOADV: 7, LIB_BASIC: 6,
             LIB_KS_GRAD: 2, LIB_KS_LAP: 1, LIB_PATCH_MODEL4: 6, LIB_PATCH_FULL: 8, LIB_PATCH_DERIVS: 6, LIB_AR_FULL: 13}


def stats_len(p: int) -> int:
    return 3 + 2 * p + p * (p + 1) // 2


class PdeGramError(RuntimeError):
    """A libpdegram call failed (message from pg_last_error)."""


_i64, _i32, _dbl, _ptr, _u64 = C.c_int64, C.c_int, C.c_double, C.c_void_p, C.c_uint64

_PROTOS = {
    "pg_version": (C.c_int, []),
    "pg_last_error": (C.c_char_p, []),
    "pg_launch_count": (C.c_int64, []),
    "pg_shutdown": (C.c_int, []),
    "pg_library_width": (C.c_int, [_i32]),
    "pg_fd_lib_gram": (C.c_int, [_ptr, _i64, _i64, _i64, _dbl, _dbl, _dbl, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr,
                                 _i32, _ptr, _ptr, _i32, _ptr]),
    "pg_fd_lib_gram_tail": (C.c_int, [_ptr, _i64, _i64, _i64, _dbl, _dbl, _dbl, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr,
                                      _i32, _ptr, _ptr, _ptr, _i32, _ptr]),
    "pg_fd_lib_gram_halo": (C.c_int, [_ptr, _i64, _i64, _i64, _dbl, _dbl, _dbl, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr,
                                      _i32, _ptr, C.c_uint32, _ptr, _ptr, _i32, _ptr]),
    "pg_comm_workspace_bytes": (C.c_size_t, []),
    "pg_comm_init": (C.c_int, [_i32, _i32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "pg_comm_destroy": (C.c_int, [_ptr]),
    "pg_comm_errors": (C.c_int64, [_ptr]),
    "pg_comm_barrier": (C.c_int, [_ptr, _ptr]),
    "pg_allreduce_stats": (C.c_int, [_ptr, _ptr, _i32, _ptr]),
    "pg_halo_exchange": (C.c_int, [_ptr, _ptr, _ptr, C.c_size_t, _ptr, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]),
    "pg_fd_lib_gram_two": (C.c_int, [_ptr, _ptr, _i64, _i64, _i64, _dbl, _dbl, _dbl, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr,
                                     _i32, _ptr, _ptr, _i32, _ptr]),
    "pg_fd_gather_rows_two": (C.c_int, [_ptr, _ptr, _i64, _i64, _i64, _dbl, _dbl, _dbl, _i32, _i32, _ptr, _i64, _ptr, _ptr, _ptr]),
    "pg_fd_terms": (C.c_int, [_ptr, _i64, _i64, _i64, _dbl, _dbl, _dbl, _i32, _i32, _ptr, _ptr]),
    "pg_fd_gather_rows": (C.c_int, [_ptr, _i64, _i64, _i64, _dbl, _dbl, _dbl, _i32, _i32, _ptr, _i64, _ptr, _ptr, _ptr]),
    "pg_block_means": (C.c_int, [_ptr, _i32, _i64, _i64, _i64, _i32, _i32, _i32, _ptr, _ptr]),
    "pg_rows_gram": (C.c_int, [_ptr, _ptr, _i64, _i64, _i32, _i64, _ptr, _i32, _ptr, _ptr, _ptr, _ptr]),
    "pg_rows_gram_weighted": (C.c_int, [_ptr, _ptr, _i64, _i32, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _ptr]),
    "pg_poly_rows": (C.c_int, [_ptr, _i32, _i64, _i64, _i64, _ptr, _i64, _ptr, _i32, _i32, _i32, _ptr, _ptr, _ptr]),
    "pg_stridge_batched": (C.c_int, [_ptr, _i64, _i32, _i32, _i32, _ptr, _i32, _ptr, _i32, _i32, _ptr, _ptr, _ptr, _ptr,
                                     _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "pg_sindy_rows": (C.c_int, [_ptr, _i64, _i64, _i64, _ptr, _i64, _i32, _i32, _i32, _dbl, _dbl, _dbl, _i32, _ptr, _ptr,
                                C.POINTER(C.c_int64), _ptr]),
    "pg_basic_library_rows": (C.c_int, [_ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr]),
    "pg_stats_accumulate": (C.c_int, [_ptr, _ptr, _i64, _ptr]),
    "pg_fd_block_rows": (C.c_int, [_ptr, _i64, _i64, _i64, _dbl, _dbl, _dbl, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr]),
    "pg_rows_residual_ss": (C.c_int, [_ptr, _ptr, _i64, _i32, _i64, _ptr, _i32, _ptr, _i32, _ptr, _ptr]),
    "pg_fd_residual_ss": (C.c_int, [_ptr, _i64, _i64, _i64, _dbl, _dbl, _dbl, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr,
                                    _i32, _i32, _ptr, _i32, _ptr, _ptr]),
    "pg_ks_rollout": (C.c_int, [_ptr, _i64, _i64, _i64, _dbl, _dbl, _dbl, _i32, _ptr, _i32, _ptr, _ptr, _ptr]),
    "pg_ar_rollout": (C.c_int, [_ptr, _i64, _i64, _i64, _dbl, _dbl, _dbl, _ptr, _ptr, _i32, _i32, _i64, _i64, _ptr, _ptr, _ptr, _ptr]),
    "pg_one_step_ss": (C.c_int, [_ptr, _ptr, _i64, _i64, _dbl, _ptr, _ptr, _ptr]),
    "pg_fit_metrics": (C.c_int, [_ptr, _ptr, _i64, _ptr, _ptr]),
    "pg_rows_metrics_batched": (C.c_int, [_ptr, _ptr, _ptr, _i64, _i64, _i32, _i64, _ptr, _ptr, _ptr]),
    "pg_reflect_conv": (C.c_int, [_ptr, _i32, _i64, _i64, _i64, _i32, _ptr, _i32, _ptr, _ptr]),
    "pg_reflect_gauss2d": (C.c_int, [_ptr, _i32, _i64, _i64, _i64, _ptr, _i32, _ptr, _ptr]),
    "pg_time_moving_average": (C.c_int, [_ptr, _i64, _i64, _i64, _i32, _ptr, _ptr]),
    "pg_periodic_conv": (C.c_int, [_ptr, _i64, _i64, _i64, _i32, _ptr, _ptr, _i32, _ptr, _ptr]),
    "pg_periodic_gaussian_fft": (C.c_int, [_ptr, _i64, _i64, _i64, _dbl, _ptr, _ptr]),
    "pg_synth_field": (C.c_int, [_ptr, _i64, _i64, _i64, _i64, _i64, _u64, _i32, _dbl, _ptr]),
}
EXPORTS = tuple(_PROTOS)

_lib = None


def build(verbose: bool = False) -> Path:
    """Compile libpdegram.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", str(CSRC_DIR), f"-j{os.cpu_count() or 4}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode:
        raise PdeGramError(f"building libpdegram.so failed (exit {res.returncode})")
    return LIB_PATH


def load():
    """Load the shared library and bind every symbol of include/pdegram.h (no compute call)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise PdeGramError(
            f"{LIB_PATH} is missing: build it with `make -C {CSRC_DIR}` (or __graft_entry__.build()); "
            "there is no CPU fallback")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().pg_last_error()
        raise PdeGramError(f"libpdegram error {rc}: {msg.decode() if msg else ''}")


def torch_cuda():
    """Return torch after checking that a CUDA device is usable; raise loudly otherwise."""
    import torch

    if not torch.cuda.is_available():
        raise PdeGramError("no CUDA device visible: pde_b200 has no CPU fallback (the hot path is CUDA-only)")
    return torch


def stream_ptr() -> int:
    torch = torch_cuda()
    return int(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> int:
    """Device pointer of a torch CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise PdeGramError("expected a contiguous CUDA tensor")
    return int(t.data_ptr())
