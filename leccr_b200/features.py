"""Drop-in for XVLMBase.get_features (models/xvlm.py:241-256; video variant models/xvlm_video.py:260-277) --
SURVEY.md section 8f rank 3, the step right in front of the similarity stage.

Same signature and returns: `F.normalize(proj(embeds[:, 0, :]), dim=-1)` per modality.  The projection stays
torch's `nn.Linear` (out of scope); the L2 normalisation runs in `leccr_normalize_fwd` -- one pass that can
also emit the 16-bit tensor-core operand -- and is wired into autograd with its own backward
(`leccr_normalize_bwd`: dx = (g - y (y . g)) / ||x||), so the features stay trainable exactly as in the reference.
"""
import torch

from . import _native as N


class _NormalizeRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if not x.is_cuda:
            raise N.LeccrError("leccr_b200 has no CPU path: normalize_rows needs a CUDA tensor")
        lib = N.load()
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        if x2.dtype != torch.float32 or x2.stride(1) != 1:
            x2 = x2.float().contiguous()
        n, D = x2.shape
        y = torch.empty((n, D), dtype=torch.float32, device=x.device)
        inv = torch.empty(n, dtype=torch.float32, device=x.device)
        if n > 0:
            N.check(lib.leccr_normalize_fwd(N.ptr(x2), n, D, x2.stride(0), N.ptr(y), D, N.ptr(inv), None, 0, N.FMT_F16,
                                            N.stream_ptr()), "leccr_normalize_fwd")
        ctx.save_for_backward(y, inv)
        ctx.in_dtype = x.dtype
        return y.view(shape)

    @staticmethod
    def backward(ctx, g):
        y, inv = ctx.saved_tensors
        n, D = y.shape
        g2 = g.reshape(n, D)
        if g2.dtype != torch.float32 or g2.stride(1) != 1:
            g2 = g2.float().contiguous()
        dx = torch.empty_like(y)
        if n > 0:
            N.check(N.load().leccr_normalize_bwd(N.ptr(y), D, N.ptr(inv), N.ptr(g2), g2.stride(0), n, D, N.ptr(dx), D,
                                                 N.stream_ptr()), "leccr_normalize_bwd")
        return dx.view(g.shape).to(ctx.in_dtype)


def normalize_rows(x: torch.Tensor) -> torch.Tensor:
    """F.normalize(x, dim=-1) on the B200 path, differentiable."""
    return _NormalizeRows.apply(x)


def get_features(self, image_embeds=None, text_embeds=None, vis_pooling='cls'):
    """models/xvlm.py:241-256 with the normalisation on the B200 path."""
    vision_proj = self.text_proj if self.vision_proj is None else self.vision_proj
    if image_embeds is None:
        return normalize_rows(self.text_proj(text_embeds[:, 0, :]))
    elif text_embeds is None:
        if vis_pooling == 'cls':
            return normalize_rows(vision_proj(image_embeds[:, 0, :]))
        elif vis_pooling == 'mean':
            return normalize_rows(vision_proj(torch.mean(image_embeds, dim=1)))
        raise ValueError("vis_pooling Error!")  # the reference prints this and exits (models/xvlm.py:251-252)
    return normalize_rows(vision_proj(image_embeds[:, 0, :])), normalize_rows(self.text_proj(text_embeds[:, 0, :]))


def get_features_video(self, image_embeds=None, text_embeds=None, vis_pooling='mean', vis_mask=None):
    """models/xvlm_video.py:260-277: masked mean pooling over the frames by default."""
    vision_proj = self.text_proj if self.vision_proj is None else self.vision_proj
    if image_embeds is None:
        return normalize_rows(self.text_proj(text_embeds[:, 0, :]))
    elif text_embeds is None:
        if vis_pooling == 'cls':
            return normalize_rows(vision_proj(image_embeds[:, 0, :]))
        elif vis_pooling == 'mean':
            image_embeds = image_embeds * vis_mask
            image_embeds = torch.sum(image_embeds, dim=1) / torch.sum(vis_mask, dim=1)
            return normalize_rows(vision_proj(image_embeds))
        raise ValueError("vis_pooling Error!")
    return normalize_rows(vision_proj(image_embeds[:, 0, :])), normalize_rows(self.text_proj(text_embeds[:, 0, :]))
