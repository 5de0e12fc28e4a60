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
import os

import torch
import torch.distributed as dist

from . import _native as N
from . import ops
from . import peer
from .allgather import gather_into

PRECISION = "f16"  # tensor-core operand format for fp32 inputs: "f16" (default) or "bf16"
# world > 1 with peer memory: every rank runs the tensor-core pass for ITS rows only and the per-row statistics
# are exchanged (leccr_itc_forward, strip forward); False = every rank runs the whole N x N forward redundantly
STRIP_FORWARD = os.environ.get("LECCR_ITC_STRIPS", "1") != "0"


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


_ws_cache = {}


def _workspace(kind, nbytes, dev):
    """Scratch reused by every call of one shape (stream-ordered, nothing in it outlives the call)."""
    key = (kind, nbytes, dev)
    t = _ws_cache.get(key)
    if t is None:
        t = torch.empty(max(1, nbytes), dtype=torch.uint8, device=dev)
        _ws_cache[key] = t
    return t


def _a256(x):
    return (x + 255) // 256 * 256


class _SymmetricInfoNCE(torch.autograd.Function):
    """forward = leccr_itc_forward, backward = leccr_itc_backward: two library calls per training step."""

    @staticmethod
    def forward(ctx, image_feat, text_feat, temp, idx, rank, world, fmt, one_dir=False):
        if not image_feat.is_cuda:
            raise N.LeccrError("leccr_b200 has no CPU path: get_contrastive_loss needs CUDA tensors")
        lib = N.load()
        B, D = image_feat.shape
        n = B * world
        dev = image_feat.device
        img, txt = image_feat, text_feat  # autograd.Function.forward runs without grad mode: no detach needed
        if img.dtype != torch.float32 or img.stride(1) != 1:
            img = img.float().contiguous()
        if txt.dtype != torch.float32 or txt.stride(1) != 1:
            txt = txt.float().contiguous()
        ix = None
        if idx is not None:
            ix = idx if idx.dim() == 1 else idx.view(-1)
            if ix.dtype != torch.int64 or not ix.is_contiguous():
                ix = ix.long().contiguous()
        temp_dev = temp if temp.dim() == 0 else temp.reshape(())
        if temp_dev.dtype != torch.float32:
            temp_dev = temp_dev.float()
        # everything the backward needs lives in ONE allocation: [both16 | idx_all | out(4) | lse2 | rcnt]
        o_idx = _a256(n * 2 * D * 2)
        o_out = o_idx + _a256(n * 8)
        o_lse = o_out + 256
        o_rc = o_lse + _a256(2 * n * 4)
        saved = torch.empty(o_rc + _a256(2 * n * 4), dtype=torch.uint8, device=dev)
        base = saved.data_ptr()
        slot = peer.itc_slot(B, D, fmt, dev) if world > 1 else None
        if world > 1 and slot is None:  # no peer memory (e.g. gloo, several nodes): NCCL all-gather of the cast rows
            dt16 = torch.float16 if fmt == N.FMT_F16 else torch.bfloat16
            local = torch.empty((B, 2 * D), dtype=dt16, device=dev)
            ops.prep_into(img, local[:, :D], fmt)
            ops.prep_into(txt, local[:, D:], fmt)
            both = gather_into(saved[:n * 2 * D * 2].view(dt16).view(n, 2 * D), local)
            idx_all = None
            if ix is not None:
                idx_all = gather_into(saved[o_idx:o_idx + n * 8].view(torch.int64), ix)
            ws = _workspace("fwd", lib.leccr_infonce_fwd_workspace(n, 0), dev)
            N.check(lib.leccr_infonce_fwd(base, base + 2 * D, 2 * D, N.ptr(idx_all), n, D, fmt, N.ptr(temp_dev),
                                          base + o_out, base + o_lse, base + o_rc, 0, N.ptr(ws), ws.numel(),
                                          N.stream_ptr()), "leccr_infonce_fwd")
        else:
            ws = _workspace("fwd", lib.leccr_itc_fwd_workspace(n, 0), dev)
            rows_tab = idx_tab = flag_tab = stat_tab = None
            epoch, l_slot, l_stat = 0, None, None
            if slot is not None:
                rows_tab, idx_tab, flag_tab, epoch, l_slot, stat_tab, l_stat = slot
                if not STRIP_FORWARD:
                    stat_tab = l_stat = None
            l_bytes = (o_idx + n * 8) if ix is not None else n * 2 * D * 2
            N.check(lib.leccr_itc_forward(N.ptr(img), img.stride(0), N.ptr(txt), txt.stride(0), N.ptr(ix), B, D, fmt,
                                          rank, world, N.ptr(rows_tab), N.ptr(idx_tab), N.ptr(flag_tab), epoch,
                                          l_slot, l_bytes, N.ptr(stat_tab), l_stat, base, base + o_idx,
                                          N.ptr(temp_dev), base + o_out,
                                          base + o_lse, base + o_rc, N.ptr(ws), ws.numel(), N.stream_ptr()),
                    "leccr_itc_forward")
        ctx.save_for_backward(saved, temp_dev)
        ctx.meta = (rank, B, D, n, fmt, idx is not None, (o_idx, o_out, o_lse, o_rc), bool(one_dir))
        # out = [loss, dloss/dtemp, loss_i2t, loss_t2i, ...]: the one-directional loss is the i2t half.  The result
        # is a 0-d view of the saved block (no copy kernel); modifying it in place trips autograd's version check.
        return saved[o_out + (8 if one_dir else 0): o_out + (12 if one_dir else 4)].view(torch.float32)[0]

    @staticmethod
    def backward(ctx, grad_out):
        saved, temp_dev = ctx.saved_tensors
        rank, B, D, n, fmt, has_idx, (o_idx, o_out, o_lse, o_rc), one_dir = ctx.meta
        lib = N.load()
        dev = saved.device
        base = saved.data_ptr()
        go = grad_out.detach().reshape(())
        if go.dtype != torch.float32:
            go = go.float()
        # separate allocations: autograd can adopt them as .grad without a copy
        dA = torch.empty((B, D), dtype=torch.float32, device=dev)
        dB = torch.empty((B, D), dtype=torch.float32, device=dev)
        dtemp = torch.empty((), dtype=torch.float32, device=dev)
        ws = _workspace("bwd", lib.leccr_itc_bwd_workspace(n, B, D), dev)
        N.check(lib.leccr_itc_backward(base, base + o_idx if has_idx else None, n, D, fmt, N.ptr(temp_dev),
                                       base + o_lse, base + o_rc, base + o_out, rank * B, B, N.ptr(go), N.ptr(dA),
                                       N.ptr(dB), N.ptr(dtemp), int(one_dir), N.ptr(ws), ws.numel(), N.stream_ptr()),
                "leccr_itc_backward")
        return dA, dB, dtemp, None, None, None, None, None


def contrastive_loss(image_feat, text_feat, temp, idx=None, precision=None, one_directional=False):
    """Functional form: temp is a 0-d tensor (parameter); uses the default process group.
    one_directional: only the first half, -mean_i sum_j log_softmax(image text^T / temp, 1)_ij labels_ij."""
    rank, world = _world()
    fmt = ops.fmt_of(precision or PRECISION)
    return _SymmetricInfoNCE.apply(image_feat, text_feat, temp, idx, rank, world, fmt, one_directional)


class _SumGradAcrossRanks(torch.autograd.Function):
    """Identity on a parameter; its gradient is summed over the ranks.  The reference applies cproj / vproj AFTER
    the gather, so every rank's module sees all N rows and holds the full parameter gradient; projecting before
    the gather leaves each rank with its own rows' share, and this sum restores the reference's per-rank value."""

    @staticmethod
    def forward(ctx, p):
        return p.view_as(p)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        return g


def _project(lin, x, world):
    if world == 1:
        return lin(x)
    w = _SumGradAcrossRanks.apply(lin.weight)
    b = _SumGradAcrossRanks.apply(lin.bias) if lin.bias is not None else None
    return torch.nn.functional.linear(x, w, b)


def caption_vision_loss(self, caption, image, idx):
    """Drop-in for RetrievalModel.caption_vision_loss (models/model_retrieval_caption.py:118-143).

    The reference gathers the raw token tensors of all ranks, projects and normalises them, multiplies every
    caption token with every image token of every pair of samples ((N cn) x (N vn) x d) and then averages the
    token pairs of each pair of samples.  The average of dot products is the dot product of the averages, and
    projection / normalisation / averaging act on each sample alone, so they commute with the gather: here every
    rank pools ITS samples first (self.cproj / self.vproj and F.normalize with the reference's default dim=1 stay
    PyTorch) and the [B, d] pooled rows go through the library's gathered multi-positive InfoNCE, first half
    only, temperature 1.  Same value and the same gradients for the local rows and -- through one all-reduce of
    the two projections' parameter gradients -- for self.cproj / self.vproj; the token-level GEMM and the gather
    of token tensors disappear."""
    import torch.nn.functional as F

    _, world = _world()
    caption = caption.transpose(0, 1).contiguous()
    c = F.normalize(_project(self.cproj, caption, world)).mean(dim=1)   # [B, d]; F.normalize's default dim=1 (:123)
    v = F.normalize(_project(self.vproj, image, world)).mean(dim=1)
    one = torch.ones((), dtype=torch.float32, device=c.device)
    return contrastive_loss(c, v, one, idx.view(-1, 1), one_directional=True)


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
