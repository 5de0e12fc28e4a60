"""Drop-in for the reference's AllGather autograd op (models/xvlm.py:50-70, dup models/xvlm_video.py:70-90).

Same call:  allgather(tensor, rank, world_size) -> (world_size * B, ...) in rank order; the gradient is
the rank's own slice of grad_output, with no reduction (models/xvlm.py:62-67).
Differences in mechanism only: one collective into one pre-laid-out buffer (no world_size temporaries, no
torch.cat copy).  The collective is torch.distributed (NCCL on the GPUs; any backend works because this
op moves bytes and computes nothing).
"""
import torch
import torch.distributed as dist


def gather_into(out: torch.Tensor, tensor: torch.Tensor, group=None) -> torch.Tensor:
    """all_gather `tensor` (B, ...) from every rank into `out` (world * B, ...)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        out.copy_(tensor)
        return out
    try:
        dist.all_gather_into_tensor(out, tensor.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):  # backend without the flat variant
        parts = list(out.chunk(world, 0))
        dist.all_gather(parts, tensor.contiguous(), group=group)
    return out


class AllGather(torch.autograd.Function):
    """An autograd function that performs allgather on a tensor (reference docstring, models/xvlm.py:51)."""

    @staticmethod
    def forward(ctx, tensor, rank, world_size):
        ctx.rank = rank
        ctx.batch_size = tensor.shape[0]
        out = tensor.new_empty((world_size * tensor.shape[0],) + tuple(tensor.shape[1:]))
        return gather_into(out, tensor)

    @staticmethod
    def backward(ctx, grad_output):
        return (
            grad_output[ctx.batch_size * ctx.rank: ctx.batch_size * (ctx.rank + 1)],
            None,
            None,
        )


allgather = AllGather.apply
