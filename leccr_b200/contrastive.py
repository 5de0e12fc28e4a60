"""Drop-in for XVLMBase.get_contrastive_loss (models/xvlm.py:260-292, dup models/xvlm_video.py:290-322).

Same signature and semantics: `get_contrastive_loss(self, image_feat, text_feat, idx=None)` reads
`self.temp` / `self.embed_dim` and the default process group, all-gathers both sides (and idx) and returns
the symmetric InfoNCE loss as a 0-d tensor wired into autograd: gradients reach the LOCAL rows of
image_feat / text_feat (what AllGather.backward keeps) and `self.temp`.

Mechanism: the local rows are cast to 16-bit tensor-core operands BEFORE the exchange; on one NVLink node
the cast kernel itself performs the exchange through peer pointers (leccr_prep_push; one barrier kernel
instead of three collectives), elsewhere one all-gather of the [image | text] halves replaces two fp32 ones; the N x N logits live only in TMEM, the
forward emits per-row log-sum-exp statistics and d loss / d temp, and the backward recomputes the logits
of the local row strips only.
"""
import torch
import torch.distributed as dist

from . import _native as N
from . import ops
from . import peer
from .allgather import gather_into

PRECISION = "f16"  # tensor-core operand format for fp32 inputs: "f16" (default) or "bf16"


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class _SymmetricInfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_feat, text_feat, temp, idx, rank, world, fmt):
        if not image_feat.is_cuda:
            raise N.LeccrError("leccr_b200 has no CPU path: get_contrastive_loss needs CUDA tensors")
        B, D = image_feat.shape
        n = B * world
        dev = image_feat.device
        dt16 = torch.float16 if fmt == N.FMT_F16 else torch.bfloat16
        # One node, NCCL ranks: the cast kernel stores each rank's rows straight into every rank's gathered
        # operand buffer through peer pointers (leccr_b200.peer).  Otherwise: local cast, then one
        # all-gather of the packed [image | text] 16-bit rows (and one of idx).
        pushed = peer.gather_contrastive(image_feat, text_feat, idx, fmt) if world > 1 else None
        if pushed is not None:
            both, idx_all = pushed
        else:
            local = torch.empty((B, 2 * D), dtype=dt16, device=dev)
            ops.prep_into(image_feat.detach().float(), local[:, :D], fmt)
            ops.prep_into(text_feat.detach().float(), local[:, D:], fmt)
            both = gather_into(torch.empty((n, 2 * D), dtype=dt16, device=dev), local)
            idx_all = None
            if idx is not None:
                idx_all = gather_into(torch.empty(n, dtype=torch.int64, device=dev), idx.detach().view(-1).long())
        a = ops.Operand(both[:, :D], fmt, N.LAYOUT_HI, n, D, None, None, None, both[:, :D])
        b = ops.Operand(both[:, D:], fmt, N.LAYOUT_HI, n, D, None, None, None, both[:, D:])
        temp_dev = temp.detach().reshape(()).float()
        out, lse2, rcnt = ops.infonce_forward(a, b, idx_all, temp_dev)
        ctx.save_for_backward(both, idx_all if idx_all is not None else torch.empty(0, device=dev), temp_dev,
                              lse2, rcnt, out)
        ctx.meta = (rank, B, D, n, fmt, idx is not None)
        return out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        both, idx_all, temp_dev, lse2, rcnt, out = ctx.saved_tensors
        rank, B, D, n, fmt, has_idx = ctx.meta
        a = ops.Operand(both[:, :D], fmt, N.LAYOUT_HI, n, D, None, None, None, both[:, :D])
        b = ops.Operand(both[:, D:], fmt, N.LAYOUT_HI, n, D, None, None, None, both[:, D:])
        aT, bT = ops.transpose16(a), ops.transpose16(b)
        go = grad_out.detach().reshape(()).float().contiguous()
        dA, dB = ops.infonce_backward(a, b, aT, bT, idx_all if has_idx else None, temp_dev, lse2, rcnt,
                                      rank * B, B, go)
        dtemp = (go * out[1]).reshape(())
        return dA, dB, dtemp, None, None, None, None


def contrastive_loss(image_feat, text_feat, temp, idx=None, precision=None):
    """Functional form: temp is a 0-d tensor (parameter); uses the default process group."""
    rank, world = _world()
    fmt = ops.fmt_of(precision or PRECISION)
    return _SymmetricInfoNCE.apply(image_feat, text_feat, temp, idx, rank, world, fmt)


def get_contrastive_loss(self, image_feat, text_feat, idx=None):
    """
    Args:
        image_feat, text_feat: normalized

    Returns: contrastive loss

    (same contract as models/xvlm.py:260-292; bind as a method of XVLMBase)
    """
    assert image_feat.size(-1) == self.embed_dim
    assert text_feat.size(-1) == self.embed_dim
    if idx is not None:
        idx = idx.view(-1, 1)
        assert idx.size(0) == image_feat.size(0)
    return contrastive_loss(image_feat, text_feat, self.temp, idx)
