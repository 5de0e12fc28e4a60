"""Drop-in for RetrievalModel.dstl_loss (models/model_retrieval_caption.py:94-116) -- SURVEY.md section 8f rank 2.

    dstl_loss(self, image_embeds [B, d], caption_embeds [n, B, d], text_embeds_s [B, d], text_embeds_t [B, d], idx,
              alpha=0.8) -> 0-d loss

Same gathers as the reference (image, captions, text_s, text_t; here ONE collective of the packed rows), then on the
gathered tensors: the three similarity products on the tensor cores with split-precision operands (the loss is a
KL divergence between two nearly uniform rows: 16-bit-operand logits would not hold its 1e-3 tolerance), the
label fusion with the library's double_sim kernel (the reference's positive norm_score differs from it by a
constant per row, which the softmax cancels), one row kernel for both log-sum-exps and the KL terms, and in the
backward the 16-bit strips of dloss/dlogits_tv for the local rows and columns and two tensor-core gradient
products.  Gradients reach image_embeds and text_embeds_t (the reference detaches the labels).
"""
import torch
import torch.distributed as dist

from . import _native as N
from . import ops
from .allgather import gather_into

PRECISION = "f16"


class _DstlGathered(torch.autograd.Function):
    """loss(image_all, caption_all [n, N, d], text_s_all, text_t_all); gradients only for the rows
    [row0, row0 + nloc) of image_all / text_t_all (what AllGather.backward keeps), zeros elsewhere."""

    @staticmethod
    def forward(ctx, image_all, caption_all, text_s_all, text_t_all, alpha, row0, nloc, fmt):
        if not image_all.is_cuda:
            raise N.LeccrError("leccr_b200 has no CPU path: dstl_loss needs CUDA tensors")
        lib = N.load()
        n, NN, D = caption_all.shape
        dev = image_all.device
        img = ops.prep(image_all.detach().float(), fmt, N.LAYOUT_X3_COLS, want_stats=False)
        tt = ops.prep(text_t_all.detach().float(), fmt, N.LAYOUT_X3_ROWS, want_stats=False)
        tsr = ops.prep(text_s_all.detach().float(), fmt, N.LAYOUT_X3_ROWS, want_stats=False)
        tsc = ops.prep(text_s_all.detach().float(), fmt, N.LAYOUT_X3_COLS, want_stats=False)
        cap = ops.prep(caption_all.detach().reshape(n * NN, D).float(), fmt, N.LAYOUT_X3_ROWS, want_stats=False)
        out = torch.empty(1, dtype=torch.float32, device=dev)
        Fm = torch.empty((NN, NN), dtype=torch.float32, device=dev)
        TV = torch.empty((NN, NN), dtype=torch.float32, device=dev)
        lse = torch.empty((2, NN), dtype=torch.float32, device=dev)
        ws = torch.empty(lib.leccr_dstl_fwd_workspace(n, NN), dtype=torch.uint8, device=dev)
        N.check(lib.leccr_dstl_fwd(N.ptr(tt.t16), 3 * D, N.ptr(tsr.t16), 3 * D, N.ptr(tsc.t16), 3 * D, N.ptr(img.t16), 3 * D,
                                   N.ptr(cap.t16), 3 * D, n, NN, 3 * D, fmt, float(alpha), N.ptr(out), N.ptr(Fm), N.ptr(TV),
                                   N.ptr(lse), N.ptr(ws), ws.numel(), N.stream_ptr()), "leccr_dstl_fwd")
        ctx.save_for_backward(Fm, TV, lse, img.t16, tt.t16)
        ctx.meta = (NN, D, fmt, int(row0), int(nloc))
        return out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        Fm, TV, lse, img16, tt16 = ctx.saved_tensors
        NN, D, fmt, row0, nloc = ctx.meta
        lib = N.load()
        dev = Fm.device
        go = grad_out.detach().reshape(()).float().contiguous()
        dimg = torch.zeros((NN, D), dtype=torch.float32, device=dev)
        dtt = torch.zeros((NN, D), dtype=torch.float32, device=dev)
        ws = torch.empty(lib.leccr_dstl_bwd_workspace(NN, nloc, D), dtype=torch.uint8, device=dev)
        # plain 16-bit halves of the split-precision buffers: img16 = [hi | hi | lo], tt16 = [hi | lo | hi]
        N.check(lib.leccr_dstl_bwd(N.ptr(Fm), N.ptr(TV), N.ptr(lse), N.ptr(img16), 3 * D, N.ptr(tt16), 3 * D, NN, D, fmt,
                                   row0, nloc, N.ptr(go), N.ptr(dimg[row0:row0 + nloc]), N.ptr(dtt[row0:row0 + nloc]),
                                   N.ptr(ws), ws.numel(), N.stream_ptr()), "leccr_dstl_bwd")
        return dimg, None, None, dtt, None, None, None, None


def dstl_loss_gathered(image_all, caption_all, text_s_all, text_t_all, alpha=0.8, row_begin=0, row_count=None,
                       precision=None):
    """The loss on already gathered tensors; gradients for rows [row_begin, row_begin + row_count) only."""
    NN = image_all.shape[0]
    if row_count is None:
        row_count = NN - row_begin
    if caption_all.dim() != 3 or caption_all.shape[1] != NN or text_s_all.shape != image_all.shape \
            or text_t_all.shape != image_all.shape:
        raise ValueError("expected image / text_s / text_t [N, d] and captions [n, N, d]")
    if image_all.shape[1] % 8 != 0:
        raise N.LeccrError("embedding dimension must be a multiple of 8 (TMA 16-byte rows)")
    return _DstlGathered.apply(image_all, caption_all, text_s_all, text_t_all, alpha, row_begin, row_count,
                               ops.fmt_of(precision or PRECISION))


class _GatherRows(torch.autograd.Function):
    """All-gather of [B, C] rows with the reference's AllGather semantics (models/xvlm.py:50-67)."""

    @staticmethod
    def forward(ctx, x, rank, world):
        ctx.rank, ctx.b = rank, x.shape[0]
        out = x.new_empty((world * x.shape[0],) + tuple(x.shape[1:]))
        return gather_into(out, x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return g[ctx.b * ctx.rank: ctx.b * (ctx.rank + 1)], None, None


def gather_packed(image_embeds, caption_embeds, text_embeds_s, text_embeds_t, rank, world):
    """The reference's four all-gathers (models/model_retrieval_caption.py:95-98) as ONE collective: every rank
    sends rows [image | text_s | text_t | caption_0 .. caption_{n-1}]; returns (image_all [N, d], caption_all
    [n, N, d], text_s_all, text_t_all) in rank order, with AllGather's backward (local slice) on all of them."""
    n, B, D = caption_embeds.shape
    packed = torch.cat([image_embeds, text_embeds_s, text_embeds_t, caption_embeds.transpose(0, 1).reshape(B, n * D)], 1)
    allp = _GatherRows.apply(packed, rank, world)
    image_all, ts_all, tt_all = allp[:, :D], allp[:, D:2 * D], allp[:, 2 * D:3 * D]
    cap_all = allp[:, 3 * D:].reshape(world * B, n, D).transpose(0, 1)
    return image_all, cap_all, ts_all, tt_all


def dstl_loss(self, image_embeds, caption_embeds, text_embeds_s, text_embeds_t, idx, alpha=0.8):
    """Same contract as models/model_retrieval_caption.py:94-116; bind as a method of RetrievalModel."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    B = image_embeds.shape[0]
    if world > 1:
        image_all, cap_all, ts_all, tt_all = gather_packed(image_embeds, caption_embeds, text_embeds_s, text_embeds_t,
                                                           rank, world)
    else:
        image_all, ts_all, tt_all, cap_all = image_embeds, text_embeds_s, text_embeds_t, caption_embeds
    return dstl_loss_gathered(image_all, cap_all, ts_all, tt_all, alpha, rank * B, B)
