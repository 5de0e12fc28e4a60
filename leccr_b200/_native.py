"""ctypes binding of include/leccr_b200.h.

There is no CPU path: if the library is missing or a call fails, this module raises.
"""
import ctypes

from . import build as _build

i64 = ctypes.c_int64
c_int = ctypes.c_int
c_float = ctypes.c_float
vp = ctypes.c_void_p
sz = ctypes.c_size_t

OK = 0
F32, F16, BF16 = 0, 1, 2
FMT_F16, FMT_BF16 = 0, 1
LAYOUT_HI, LAYOUT_X3_ROWS, LAYOUT_X3_COLS = 0, 1, 2
FUSE_NORM, FUSE_RAW = 1, 2
STAT_WORDS = 4
TOPK_KP = 16
RANK_CAP = 10


class TopkProblem(ctypes.Structure):
    """Mirror of struct leccr_topk_problem."""

    _fields_ = [
        ("rows16", vp), ("cols16", vp),
        ("ld_rows16", i64), ("ld_cols16", i64),
        ("n_rows", i64), ("n_cols", i64),
        ("topk_val", vp), ("topk_idx", vp),
        ("gt_off", vp), ("gt_ids", vp),
        ("rows_x", vp), ("cols_x", vp),
        ("ld_rows_x", i64), ("ld_cols_x", i64),
        ("x_dtype", c_int),
        ("rn_hi", vp), ("rn_lo", vp), ("col_stats", vp),
        ("rank", vp), ("recall_counts", vp), ("gt_score", vp),
    ]


class TopkStream(ctypes.Structure):
    """Mirror of struct leccr_topk_stream."""

    _fields_ = [
        ("phases", ctypes.c_int32), ("sub_begin", ctypes.c_int32), ("sub_count", ctypes.c_int32),
        ("sub_total", ctypes.c_int32), ("col_begin", i64), ("n_cols_total", i64),
        ("workspace", vp), ("workspace_bytes", sz),
    ]


TOPK_INIT, TOPK_GEMM, TOPK_FINALIZE, TOPK_LONG = 1, 2, 4, 8

_SIGNATURES = {
    "leccr_strerror": (ctypes.c_char_p, [c_int]),
    "leccr_last_cuda_error": (ctypes.c_char_p, []),
    "leccr_abi_version": (c_int, []),
    "leccr_check_device": (c_int, []),
    "leccr_profile_enable": (None, [c_int]),
    "leccr_profile_read": (c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_int)]),
    "leccr_prep": (c_int, [vp, i64, c_int, i64, c_int, c_int, c_int, vp, i64, vp, vp, vp, vp]),
    "leccr_prep_pair": (c_int, [vp, i64, i64, vp, i64, vp, vp, vp, vp, i64, i64, vp, i64, vp, vp, vp, c_int, c_int, c_int,
                                c_int, vp]),
    "leccr_prep_push": (c_int, [vp, i64, c_int, i64, c_int, c_int, vp, c_int, i64, i64, i64, vp]),
    "leccr_push_words": (c_int, [vp, i64, vp, c_int, i64, vp]),
    "leccr_sim_topk_stream_workspace": (sz, [i64, c_int]),
    "leccr_sim_topk_stream": (c_int, [ctypes.POINTER(TopkProblem), ctypes.POINTER(TopkStream), c_int, c_int, c_int,
                                      c_int, vp]),
    "leccr_itc_fwd_workspace": (sz, [i64, c_int]),
    "leccr_itc_forward": (c_int, [vp, i64, vp, i64, vp, i64, c_int, c_int, c_int, c_int, vp, vp, vp, ctypes.c_uint32,
                                  vp, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "leccr_itc_bwd_workspace": (sz, [i64, i64, c_int]),
    "leccr_itc_backward": (c_int, [vp, vp, i64, c_int, c_int, vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, c_int, vp, sz,
                                   vp]),
    "leccr_caploss_fwd_workspace": (sz, [c_int, i64]),
    "leccr_caploss_fwd": (c_int, [vp, i64, vp, i64, c_int, i64, c_int, c_int, vp, vp, vp, vp, vp, vp, sz, vp]),
    "leccr_caploss_bwd_workspace": (sz, [c_int, i64, c_int]),
    "leccr_caploss_bwd": (c_int, [vp, vp, vp, vp, i64, vp, i64, c_int, i64, c_int, c_int, vp, vp, vp, vp, vp, vp, vp,
                                  sz, vp]),
    "leccr_topk_dense": (c_int, [vp, i64, i64, i64, c_int, c_int, vp, vp, vp]),
    "leccr_comm_unique_id": (c_int, [vp]),
    "leccr_comm_init": (c_int, [vp, c_int, c_int, ctypes.POINTER(vp)]),
    "leccr_comm_destroy": (c_int, [vp]),
    "leccr_allgather": (c_int, [vp, vp, vp, sz, vp]),
    "leccr_dstl_fwd_workspace": (sz, [c_int, i64]),
    "leccr_dstl_fwd": (c_int, [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, c_int, i64, c_int, c_int, c_float, vp, vp, vp,
                               vp, vp, sz, vp]),
    "leccr_dstl_bwd_workspace": (sz, [i64, i64, c_int]),
    "leccr_dstl_bwd": (c_int, [vp, vp, vp, vp, i64, vp, i64, i64, c_int, c_int, i64, i64, vp, vp, vp, vp, sz, vp]),
    "leccr_peer_barrier": (c_int, [vp, c_int, c_int, ctypes.c_uint32, vp]),
    "leccr_double_sim_topk_workspace": (sz, [i64, i64, c_int]),
    "leccr_double_sim_topk": (c_int, [vp, vp, i64, i64, c_int, c_int, c_int, c_int, c_float, c_float, c_int, vp, vp, vp,
                                      c_int, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "leccr_normalize_fwd": (c_int, [vp, i64, c_int, i64, vp, i64, vp, vp, i64, c_int, vp]),
    "leccr_normalize_bwd": (c_int, [vp, i64, vp, vp, i64, i64, c_int, vp, i64, vp]),
    "leccr_memcpy_peer_async": (c_int, [vp, vp, sz, vp]),
    "leccr_topk_merge_peers": (c_int, [vp, vp, c_int, c_int, i64, i64, ctypes.POINTER(i64), c_int, vp, vp, vp]),
    "leccr_stats16": (c_int, [vp, c_int, i64, c_int, i64, vp, vp, vp, vp]),
    "leccr_transpose16": (c_int, [vp, i64, c_int, i64, vp, i64, vp]),
    "leccr_sim_f32": (c_int, [vp, i64, vp, i64, i64, i64, c_int, c_int, vp, i64, c_float, vp, vp]),
    "leccr_sim_rank_workspace": (sz, [ctypes.POINTER(TopkProblem), c_int]),
    "leccr_sim_rank": (c_int, [ctypes.POINTER(TopkProblem), c_int, c_int, c_int, vp, sz, vp]),
    "leccr_sim_topk_workspace": (sz, [ctypes.POINTER(TopkProblem), c_int, c_int]),
    "leccr_sim_topk": (c_int, [ctypes.POINTER(TopkProblem), c_int, c_int, c_int, c_int, c_int, vp, sz, vp]),
    "leccr_infonce_fwd_workspace": (sz, [i64, c_int]),
    "leccr_infonce_fwd": (c_int, [vp, vp, i64, vp, i64, c_int, c_int, vp, vp, vp, vp, c_int, vp, sz, vp]),
    "leccr_infonce_bwd_workspace": (sz, [i64, i64, c_int]),
    "leccr_infonce_bwd": (c_int, [vp, vp, i64, vp, vp, i64, vp, i64, c_int, c_int, vp, vp, vp, i64, i64,
                                  vp, vp, vp, vp, sz, vp]),
    "leccr_rank_rows": (c_int, [vp, i64, i64, i64, vp, vp, vp, vp]),
    "leccr_rank_cols": (c_int, [vp, i64, i64, i64, vp, vp, vp, vp, vp]),
    "leccr_recall_counts": (c_int, [vp, i64, vp, vp]),
    "leccr_double_sim_fuse": (c_int, [vp, vp, c_int, i64, vp, vp, c_float, c_float, c_int, vp]),
}

EXPORTS = tuple(_SIGNATURES)

_lib = None


class LeccrError(RuntimeError):
    pass


def lib_path():
    return _build.LIB


def load():
    """Load (building first if the .so is absent) and type the C ABI. Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if _build.is_stale():  # missing, or built from other sources than the ones in the tree
        _build.build()
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != OK:
        lib = load()
        msg = lib.leccr_strerror(rc).decode()
        detail = lib.leccr_last_cuda_error().decode()
        raise LeccrError(f"{what} failed: {msg}" + (f" [{detail}]" if detail else ""))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    """Raw cudaStream_t of torch's current stream on the current device (the per-call host cost matters: the
    training losses are host-issue-bound)."""
    import torch

    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:  # older / newer torch without the private accessor
        return torch.cuda.current_stream().cuda_stream
