"""Device-level operations: torch CUDA tensors in, C-ABI calls on the current stream, tensors out.

torch is used for device memory and streams only; all arithmetic happens in libleccr_b200.so.
"""
from dataclasses import dataclass
from typing import Optional

import torch

from . import _native as N


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise N.LeccrError(f"{name} must be a CUDA tensor: leccr_b200 has no CPU path")


def fmt_of(precision: str) -> int:
    if precision in ("f16", "f16x3"):
        return N.FMT_F16
    if precision in ("bf16", "bf16x3"):
        return N.FMT_BF16
    raise ValueError(f"unknown precision {precision!r}")


@dataclass
class Operand:
    """A tensor-core operand: 16-bit rows [n, K] plus the statistics the exactness logic needs."""

    t16: torch.Tensor            # [n, K] fp16 / bf16 (K = D or 3D)
    fmt: int
    layout: int
    n: int
    D: int
    rn_hi: Optional[torch.Tensor]  # [n] ||hi_i||
    rn_lo: Optional[torch.Tensor]  # [n] ||x_i - hi_i||
    stats: Optional[torch.Tensor]  # [STAT_WORDS]
    x: torch.Tensor              # the original tensor (exact re-scoring reads it)

    @property
    def K(self):
        return self.t16.shape[1]

    @property
    def x_dtype(self):
        return {torch.float32: N.F32, torch.float16: N.F16, torch.bfloat16: N.BF16}[self.x.dtype]

    def rows(self, begin: int, end: int) -> "Operand":
        """Rows [begin, end) as an operand of their own (views; the per-tensor stats stay the whole tensor's)."""
        return Operand(self.t16[begin:end], self.fmt, self.layout, end - begin, self.D,
                       None if self.rn_hi is None else self.rn_hi[begin:end],
                       None if self.rn_lo is None else self.rn_lo[begin:end], self.stats, self.x[begin:end])


def prep(x: torch.Tensor, fmt: int = N.FMT_F16, layout: int = N.LAYOUT_HI, normalize: bool = False,
         want_stats: bool = True) -> Operand:
    """fp32 [n, D] -> 16-bit operand (leccr_prep); fp16/bf16 inputs are used in place (leccr_stats16)."""
    _require_cuda(x, "x")
    if x.dim() != 2:
        raise ValueError("operand must be 2-D [n, D]")
    lib = N.load()
    n, D = x.shape
    if D % 8 != 0:
        raise N.LeccrError("embedding dimension must be a multiple of 8 (TMA 16-byte rows)")
    dev = x.device
    rn_hi = torch.empty(n, dtype=torch.float32, device=dev) if want_stats else None
    rn_lo = torch.empty(n, dtype=torch.float32, device=dev) if want_stats else None
    stats = torch.zeros(N.STAT_WORDS, dtype=torch.float32, device=dev) if want_stats else None
    if x.dtype in (torch.float16, torch.bfloat16):
        if normalize or layout != N.LAYOUT_HI:
            raise N.LeccrError("16-bit inputs are consumed as they are (no normalise / split)")
        x = x.contiguous()
        f = N.FMT_F16 if x.dtype == torch.float16 else N.FMT_BF16
        if want_stats:
            N.check(lib.leccr_stats16(N.ptr(x), f, n, D, x.stride(0), N.ptr(rn_hi), N.ptr(rn_lo), N.ptr(stats),
                                      N.stream_ptr()), "leccr_stats16")
        return Operand(x, f, N.LAYOUT_HI, n, D, rn_hi, rn_lo, stats, x)
    if x.dtype != torch.float32:
        raise N.LeccrError(f"unsupported dtype {x.dtype}")
    if x.stride(1) != 1:
        x = x.contiguous()
    K = D if layout == N.LAYOUT_HI else 3 * D
    t16 = torch.empty((n, K), dtype=torch.float16 if fmt == N.FMT_F16 else torch.bfloat16, device=dev)
    N.check(lib.leccr_prep(N.ptr(x), n, D, x.stride(0), int(normalize), fmt, layout, N.ptr(t16), K,
                           N.ptr(rn_hi), N.ptr(rn_lo), N.ptr(stats), N.stream_ptr()), "leccr_prep")
    return Operand(t16, fmt, layout, n, D, rn_hi, rn_lo, stats, x)


def prep_into(x: torch.Tensor, dst: torch.Tensor, fmt: int, normalize: bool = False) -> torch.Tensor:
    """Cast fp32 [n, D] into the (possibly strided) 16-bit view `dst` [n, D]; no statistics."""
    _require_cuda(x, "x")
    lib = N.load()
    n, D = x.shape
    if x.dtype != torch.float32 or dst.shape != x.shape or dst.stride(1) != 1:
        raise N.LeccrError("prep_into needs fp32 [n, D] input and a row-major 16-bit destination view")
    if x.stride(1) != 1:
        x = x.contiguous()
    N.check(lib.leccr_prep(N.ptr(x), n, D, x.stride(0), int(normalize), fmt, N.LAYOUT_HI, N.ptr(dst), dst.stride(0),
                           None, None, None, N.stream_ptr()), "leccr_prep")
    return dst


def transpose16(op: Operand) -> torch.Tensor:
    """[n, D] 16-bit -> [D, ld] with ld = n rounded up to 8 (zero padded)."""
    lib = N.load()
    ld = (op.n + 7) // 8 * 8
    out = torch.empty((op.D, ld), dtype=op.t16.dtype, device=op.t16.device)
    N.check(lib.leccr_transpose16(N.ptr(op.t16), op.n, op.D, op.t16.stride(0), N.ptr(out), ld, N.stream_ptr()),
            "leccr_transpose16")
    return out


def sim_matrix(rows: Operand, cols: Operand, scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """S = scale * rows . cols^T, fp32, materialised (image_Retrieval_caption.py:151)."""
    lib = N.load()
    if rows.fmt != cols.fmt or rows.K != cols.K:
        raise N.LeccrError("operands disagree in format or K")
    if (rows.layout, cols.layout) not in ((N.LAYOUT_HI, N.LAYOUT_HI), (N.LAYOUT_X3_ROWS, N.LAYOUT_X3_COLS)):
        raise N.LeccrError("operand layouts must be (HI, HI) or (X3_ROWS, X3_COLS)")
    if out is None:
        out = torch.empty((rows.n, cols.n), dtype=torch.float32, device=rows.t16.device)
    N.check(lib.leccr_sim_f32(N.ptr(rows.t16), rows.t16.stride(0), N.ptr(cols.t16), cols.t16.stride(0), rows.n,
                              cols.n, rows.K, rows.fmt, N.ptr(out), out.stride(0), float(scale), None,
                              N.stream_ptr()), "leccr_sim_f32")
    return out


def csr_from_lists(lists, device):
    """Ground-truth lists (one per row) -> CSR (offsets int32 [n+1], ids int32 [nnz]) on `device`."""
    off = [0]
    ids = []
    for l in lists:
        ids.extend(int(v) for v in l)
        off.append(len(ids))
    return (torch.tensor(off, dtype=torch.int32, device=device),
            torch.tensor(ids, dtype=torch.int32, device=device))


@dataclass
class TopkResult:
    val: torch.Tensor                    # [n_rows, k] approximate scores, descending
    idx: torch.Tensor                    # [n_rows, k] int32 columns
    rank: Optional[torch.Tensor]         # [n_rows] int32 (exact when < RANK_CAP)
    recall_counts: Optional[torch.Tensor]  # [3] int32 #{rank < 1, 5, 10}
    gt_score: Optional[torch.Tensor]     # [nnz] exact fp32 ground-truth scores


def sim_topk(problems, k: int = 10, tiles_per_chunk: int = 0):
    """Fused similarity + per-row top-k (+ exact Recall ranks) for 1 or 2 (rows, cols, gt) problems.

    problems: list of (rows: Operand, cols: Operand, gt) with gt = None or (gt_off, gt_ids) CSR tensors.
    All problems of a call share one tensor-core launch.
    """
    lib = N.load()
    if not 1 <= len(problems) <= 2:
        raise ValueError("one or two problems per launch")
    arr = (N.TopkProblem * len(problems))()
    keep = []
    results = []
    fmt = problems[0][0].fmt
    D = problems[0][0].D
    for i, (rows, cols, gt) in enumerate(problems):
        if rows.layout != N.LAYOUT_HI or cols.layout != N.LAYOUT_HI:
            raise N.LeccrError("sim_topk takes single-pass (HI) operands")
        if rows.fmt != fmt or cols.fmt != fmt or rows.D != D or cols.D != D:
            raise N.LeccrError("all operands of a launch must share format and dimension")
        dev = rows.t16.device
        val = torch.empty((rows.n, k), dtype=torch.float32, device=dev)
        idx = torch.empty((rows.n, k), dtype=torch.int32, device=dev)
        p = arr[i]
        p.rows16, p.cols16 = N.ptr(rows.t16), N.ptr(cols.t16)
        p.ld_rows16, p.ld_cols16 = rows.t16.stride(0), cols.t16.stride(0)
        p.n_rows, p.n_cols = rows.n, cols.n
        p.topk_val, p.topk_idx = N.ptr(val), N.ptr(idx)
        rank = counts = gts = None
        if gt is not None:
            gt_off, gt_ids = gt
            if rows.x.dtype != cols.x.dtype:
                raise N.LeccrError("exact re-scoring needs both originals in one dtype")
            rank = torch.empty(rows.n, dtype=torch.int32, device=dev)
            counts = torch.zeros(3, dtype=torch.int32, device=dev)
            gts = torch.empty(max(1, gt_ids.numel()), dtype=torch.float32, device=dev)
            p.gt_off, p.gt_ids = N.ptr(gt_off), N.ptr(gt_ids)
            p.rows_x, p.cols_x = N.ptr(rows.x), N.ptr(cols.x)
            p.ld_rows_x, p.ld_cols_x = rows.x.stride(0), cols.x.stride(0)
            p.x_dtype = rows.x_dtype
            p.rn_hi, p.rn_lo, p.col_stats = N.ptr(rows.rn_hi), N.ptr(rows.rn_lo), N.ptr(cols.stats)
            p.rank, p.recall_counts, p.gt_score = N.ptr(rank), N.ptr(counts), N.ptr(gts)
            keep += [gt_off, gt_ids]
        results.append(TopkResult(val, idx, rank, counts, gts))
    ws_bytes = lib.leccr_sim_topk_workspace(arr, len(problems), tiles_per_chunk)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=problems[0][0].t16.device)
    N.check(lib.leccr_sim_topk(arr, len(problems), D, fmt, k, tiles_per_chunk, N.ptr(ws), ws_bytes,
                               N.stream_ptr()), "leccr_sim_topk")
    return results


def sim_rank(problems):
    """Fused similarity + exact Recall ranks WITHOUT top-k lists (leccr_sim_rank) for 1 or 2 (rows, cols, gt) problems:
    what the reference's itm_eval returns.  gt = (gt_off, gt_ids) CSR tensors (required).  Returns TopkResult objects
    with val / idx = None."""
    lib = N.load()
    if not 1 <= len(problems) <= 2:
        raise ValueError("one or two problems per launch")
    arr = (N.TopkProblem * len(problems))()
    keep, results = [], []
    fmt, D = problems[0][0].fmt, problems[0][0].D
    for i, (rows, cols, gt) in enumerate(problems):
        if gt is None:
            raise N.LeccrError("sim_rank needs the ground truth of every problem")
        if rows.layout != N.LAYOUT_HI or cols.layout != N.LAYOUT_HI:
            raise N.LeccrError("sim_rank takes single-pass (HI) operands")
        if rows.fmt != fmt or cols.fmt != fmt or rows.D != D or cols.D != D:
            raise N.LeccrError("all operands of a launch must share format and dimension")
        if rows.x.dtype != cols.x.dtype:
            raise N.LeccrError("exact re-scoring needs both originals in one dtype")
        if rows.rn_hi is None or cols.stats is None:
            raise N.LeccrError("sim_rank needs operands prepared with statistics (ops.prep(..., want_stats=True))")
        dev = rows.t16.device
        gt_off, gt_ids = gt
        rank = torch.empty(rows.n, dtype=torch.int32, device=dev)
        counts = torch.zeros(3, dtype=torch.int32, device=dev)
        gts = torch.empty(max(1, gt_ids.numel()), dtype=torch.float32, device=dev)
        p = arr[i]
        p.rows16, p.cols16 = N.ptr(rows.t16), N.ptr(cols.t16)
        p.ld_rows16, p.ld_cols16 = rows.t16.stride(0), cols.t16.stride(0)
        p.n_rows, p.n_cols = rows.n, cols.n
        p.gt_off, p.gt_ids = N.ptr(gt_off), N.ptr(gt_ids)
        p.rows_x, p.cols_x = N.ptr(rows.x), N.ptr(cols.x)
        p.ld_rows_x, p.ld_cols_x = rows.x.stride(0), cols.x.stride(0)
        p.x_dtype = rows.x_dtype
        p.rn_hi, p.rn_lo, p.col_stats = N.ptr(rows.rn_hi), N.ptr(rows.rn_lo), N.ptr(cols.stats)
        p.rank, p.recall_counts, p.gt_score = N.ptr(rank), N.ptr(counts), N.ptr(gts)
        keep += [gt_off, gt_ids]
        results.append(TopkResult(None, None, rank, counts, gts))
    ws = torch.empty(lib.leccr_sim_rank_workspace(arr, len(problems)), dtype=torch.uint8, device=problems[0][0].t16.device)
    N.check(lib.leccr_sim_rank(arr, len(problems), D, fmt, N.ptr(ws), ws.numel(), N.stream_ptr()), "leccr_sim_rank")
    return results


def infonce_forward(a: Operand, b: Operand, idx: Optional[torch.Tensor], temp: torch.Tensor,
                    tiles_per_chunk: int = 0):
    """Returns (out[6] = loss, dloss/dtemp, loss_i2t, loss_t2i, dloss_i2t/dtemp, dloss_t2i/dtemp ; lse2 [2, n] ;
    rcnt [2, n])."""
    lib = N.load()
    n = a.n
    if b.n != n or a.D != b.D or a.fmt != b.fmt:
        raise N.LeccrError("image and text operands must have the same shape and format")
    dev = a.t16.device
    out = torch.empty(6, dtype=torch.float32, device=dev)
    lse2 = torch.empty((2, n), dtype=torch.float32, device=dev)
    rcnt = torch.empty((2, n), dtype=torch.float32, device=dev)
    ws_bytes = lib.leccr_infonce_fwd_workspace(n, tiles_per_chunk)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    N.check(lib.leccr_infonce_fwd(N.ptr(a.t16), N.ptr(b.t16), a.t16.stride(0), N.ptr(idx), n, a.D, a.fmt,
                                  N.ptr(temp), N.ptr(out), N.ptr(lse2), N.ptr(rcnt), tiles_per_chunk, N.ptr(ws),
                                  ws_bytes, N.stream_ptr()), "leccr_infonce_fwd")
    return out, lse2, rcnt


def infonce_backward(a: Operand, b: Operand, aT: Optional[torch.Tensor], bT: Optional[torch.Tensor],
                     idx: Optional[torch.Tensor], temp: torch.Tensor, lse2: torch.Tensor, rcnt: torch.Tensor,
                     row_begin: int, row_count: int, grad_out: torch.Tensor):
    """Gradients w.r.t. the local rows [row_begin, row_begin + row_count) of the gathered operands.
    aT / bT (transposed copies, `transpose16`) are only read with LECCR_BWD_MN=0; pass None otherwise: the
    gradient products read the gathered rows as MN-major tensor-core operands."""
    lib = N.load()
    dev = a.t16.device
    dA = torch.empty((row_count, a.D), dtype=torch.float32, device=dev)
    dB = torch.empty((row_count, a.D), dtype=torch.float32, device=dev)
    ws_bytes = lib.leccr_infonce_bwd_workspace(a.n, row_count, a.D)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    N.check(lib.leccr_infonce_bwd(N.ptr(a.t16), N.ptr(b.t16), a.t16.stride(0), N.ptr(aT), N.ptr(bT),
                                  aT.stride(0) if aT is not None else 0, N.ptr(idx), a.n, a.D, a.fmt, N.ptr(temp), N.ptr(lse2), N.ptr(rcnt), row_begin,
                                  row_count, N.ptr(grad_out), N.ptr(dA), N.ptr(dB), N.ptr(ws), ws_bytes,
                                  N.stream_ptr()), "leccr_infonce_bwd")
    return dA, dB


def rank_rows(S: torch.Tensor, gt_off, gt_ids) -> torch.Tensor:
    lib = N.load()
    _require_cuda(S, "S")
    R, C = S.shape
    rank = torch.empty(R, dtype=torch.int32, device=S.device)
    N.check(lib.leccr_rank_rows(N.ptr(S), S.stride(0), R, C, N.ptr(gt_off), N.ptr(gt_ids), N.ptr(rank),
                                N.stream_ptr()), "leccr_rank_rows")
    return rank


def rank_cols(S: torch.Tensor, gt_off, gt_ids) -> torch.Tensor:
    lib = N.load()
    _require_cuda(S, "S")
    R, C = S.shape
    rank = torch.empty(C, dtype=torch.int32, device=S.device)
    scratch = torch.zeros(max(1, gt_ids.numel()), dtype=torch.int32, device=S.device)
    N.check(lib.leccr_rank_cols(N.ptr(S), S.stride(0), R, C, N.ptr(gt_off), N.ptr(gt_ids), N.ptr(scratch),
                                N.ptr(rank), N.stream_ptr()), "leccr_rank_cols")
    return rank


def recall_counts(rank: torch.Tensor) -> torch.Tensor:
    lib = N.load()
    counts = torch.zeros(3, dtype=torch.int32, device=rank.device)
    N.check(lib.leccr_recall_counts(N.ptr(rank), rank.numel(), N.ptr(counts), N.stream_ptr()),
            "leccr_recall_counts")
    return counts


def double_sim_fuse(S: torch.Tensor, Cn: torch.Tensor, alpha: float, mode: int) -> torch.Tensor:
    """In place S <- alpha * f(S) + (1 - alpha) * f(max_n Cn); Cn is [n_cap, *S.shape]."""
    lib = N.load()
    if not S.is_contiguous() or not Cn.is_contiguous():
        raise N.LeccrError("double_sim_fuse needs contiguous matrices")
    n_cap = Cn.shape[0]
    cmax = torch.empty_like(S)
    mm = torch.empty(4, dtype=torch.int32, device=S.device)
    N.check(lib.leccr_double_sim_fuse(N.ptr(S), N.ptr(Cn), n_cap, S.numel(), N.ptr(cmax), N.ptr(mm),
                                      float(alpha), float(1.0 - alpha), mode, N.stream_ptr()),
            "leccr_double_sim_fuse")
    return S


def topk_dense(S: torch.Tensor, k: int, by_columns: bool = False):
    """Top-k of the rows (or columns) of a materialised fp32 matrix: (val, idx int32), score descending."""
    lib = N.load()
    _require_cuda(S, "S")
    if S.dtype != torch.float32 or S.stride(1) != 1:
        raise N.LeccrError("topk_dense needs a row-major fp32 matrix")
    R, C = S.shape
    n = C if by_columns else R
    val = torch.empty((n, k), dtype=torch.float32, device=S.device)
    idx = torch.empty((n, k), dtype=torch.int32, device=S.device)
    N.check(lib.leccr_topk_dense(N.ptr(S), S.stride(0), R, C, int(by_columns), k, N.ptr(val), N.ptr(idx),
                                 N.stream_ptr()), "leccr_topk_dense")
    return val, idx
