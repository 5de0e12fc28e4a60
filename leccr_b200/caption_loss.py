"""Drop-in for RetrievalModel.get_caption_contrastive_loss (models/model_retrieval_caption.py:145-152, same code in
models/video_model_retrieval_caption.py:171-178) -- SURVEY.md section 8f rank 1.

    get_caption_contrastive_loss(self, caption_embeds [n, B, d], text_feats [B, d]) -> 0-d loss

Local batch only (the reference does not gather here).  sim = caption.reshape(n*B, d) @ text.T runs on the tensor
cores with split-precision operands (so the max over the n caption queries picks the same query as fp32 does),
the max / arg max, both log-sum-exp families and the loss are small CUDA kernels over the n*B*B matrix, and the
backward routes G = dloss/dlogits to the arg-max query and forms d caption = G text and d text = G^T caption as
two more tensor-core products.  Gradients reach caption_embeds, text_feats and self.temp.
"""
import torch

from . import _native as N
from . import ops

PRECISION = "f16"


class _CaptionInfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, caption_embeds, text_feats, temp, fmt):
        if not caption_embeds.is_cuda:
            raise N.LeccrError("leccr_b200 has no CPU path: get_caption_contrastive_loss needs CUDA tensors")
        lib = N.load()
        n, B, D = caption_embeds.shape
        dev = caption_embeds.device
        cap = ops.prep(caption_embeds.detach().reshape(n * B, D).float(), fmt, N.LAYOUT_X3_ROWS)
        txt = ops.prep(text_feats.detach().float(), fmt, N.LAYOUT_X3_COLS)
        temp_dev = temp.detach().reshape(()).float()
        out = torch.empty(2, dtype=torch.float32, device=dev)
        L = torch.empty((B, B), dtype=torch.float32, device=dev)
        amax = torch.empty((B, B), dtype=torch.uint8, device=dev)
        stats = torch.empty((4, B), dtype=torch.float32, device=dev)
        ws = torch.empty(lib.leccr_caploss_fwd_workspace(n, B), dtype=torch.uint8, device=dev)
        N.check(lib.leccr_caploss_fwd(N.ptr(cap.t16), cap.t16.stride(0), N.ptr(txt.t16), txt.t16.stride(0), n, B, 3 * D,
                                      fmt, N.ptr(temp_dev), N.ptr(out), N.ptr(L), N.ptr(amax), N.ptr(stats), N.ptr(ws),
                                      ws.numel(), N.stream_ptr()), "leccr_caploss_fwd")
        # |x| > 65504 cannot be an fp16 operand: the cast then yields inf and the loss comes out inf / NaN (as under
        # fp16 autocast), it is never a plausible wrong number.  No flag is read back here -- that would cost a host
        # sync per training step; un-normalised inputs of that size want precision="bf16" (fp32 range).
        ctx.save_for_backward(cap.t16, txt.t16, temp_dev, out, L, amax, stats, cap.stats, txt.stats)
        ctx.meta = (n, B, D, fmt)
        return out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        cap16, txt16, temp_dev, out, L, amax, stats, cstats, tstats = ctx.saved_tensors
        n, B, D, fmt = ctx.meta
        lib = N.load()
        dev = cap16.device
        go = grad_out.detach().reshape(()).float().contiguous()
        dcap = torch.empty((n, B, D), dtype=torch.float32, device=dev)
        dtxt = torch.empty((B, D), dtype=torch.float32, device=dev)
        dtemp = torch.empty((), dtype=torch.float32, device=dev)
        ws = torch.empty(lib.leccr_caploss_bwd_workspace(n, B, D), dtype=torch.uint8, device=dev)
        # the [hi | ...] first D columns of the split-precision buffers are the plain 16-bit operands (ld = 3D)
        N.check(lib.leccr_caploss_bwd(N.ptr(L), N.ptr(amax), N.ptr(stats), N.ptr(cap16), cap16.stride(0), N.ptr(txt16),
                                      txt16.stride(0), n, B, D, fmt, N.ptr(temp_dev), N.ptr(out), N.ptr(go),
                                      N.ptr(dcap), N.ptr(dtxt), N.ptr(dtemp), N.ptr(ws), ws.numel(), N.stream_ptr()),
                "leccr_caploss_bwd")
        return dcap, dtxt, dtemp, None


def caption_contrastive_loss(caption_embeds, text_feats, temp, precision=None):
    """Functional form: temp is a 0-d tensor (parameter)."""
    if caption_embeds.dim() != 3 or text_feats.dim() != 2 or caption_embeds.shape[1] != text_feats.shape[0] \
            or caption_embeds.shape[2] != text_feats.shape[1]:
        raise ValueError("caption_embeds must be [n, B, d] and text_feats [B, d]")
    if caption_embeds.shape[2] % 8 != 0:
        raise N.LeccrError("embedding dimension must be a multiple of 8 (TMA 16-byte rows)")
    return _CaptionInfoNCE.apply(caption_embeds, text_feats, temp, ops.fmt_of(precision or PRECISION))


def get_caption_contrastive_loss(self, caption_embeds, text_feats):
    """Same contract as models/model_retrieval_caption.py:145-152; bind as a method of RetrievalModel."""
    return caption_contrastive_loss(caption_embeds, text_feats, self.temp)
