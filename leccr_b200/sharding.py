"""Multi-GPU decomposition of the evaluation path.

Queries (the rows being ranked) shard with no exchange at all.  A row-partitioned gallery needs one
exchange: every rank produces a per-query top-k over its gallery shard (global column = local + shard
offset), the lists are all-gathered (Q * k * 8 bytes per rank) and merged.
"""
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) of `n` items for `rank` of `world`."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def merge_topk(vals: torch.Tensor, idx: torch.Tensor, k: int):
    """vals / idx: [W, Q, k] partial lists (idx already global) -> merged [Q, k], descending,
    ties by lower column."""
    w, q, kk = vals.shape
    v = vals.permute(1, 0, 2).reshape(q, w * kk)
    i = idx.permute(1, 0, 2).reshape(q, w * kk)
    # stable order: by column first, then a stable sort by score keeps lower columns first among ties
    order = torch.argsort(i, dim=1, stable=True)
    v, i = torch.gather(v, 1, order), torch.gather(i, 1, order)
    order = torch.argsort(v, dim=1, descending=True, stable=True)[:, :k]
    return torch.gather(v, 1, order), torch.gather(i, 1, order)


def allgather_topk(vals: torch.Tensor, idx: torch.Tensor, k: int, group=None):
    """All-gather every rank's [Q, k] partial lists (idx already global) and merge them (identical result on
    every rank).  CUDA tensors are merged by the library's own kernel (leccr_topk_merge_peers over a table of
    pointers into the gathered buffer -- the path taken when peer memory is unavailable, e.g. across nodes);
    `merge_topk` (torch) remains for CPU tensors: the host-logic tests under gloo."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return vals, idx
    gv = torch.empty((world,) + tuple(vals.shape), dtype=vals.dtype, device=vals.device)
    gi = torch.empty((world,) + tuple(idx.shape), dtype=idx.dtype, device=idx.device)
    dist.all_gather_into_tensor(gv, vals.contiguous(), group=group)
    dist.all_gather_into_tensor(gi, idx.contiguous(), group=group)
    if not vals.is_cuda or world > 8 or vals.dtype != torch.float32:
        return merge_topk(gv, gi, k)
    import ctypes

    from . import _native as N

    q, k_in = vals.shape
    gi32 = gi.to(torch.int32)  # global rows of a gallery fit 31 bits (the kernel's column type)
    vt = torch.tensor([gv[r].data_ptr() for r in range(world)], dtype=torch.int64, device=vals.device)
    it = torch.tensor([gi32[r].data_ptr() for r in range(world)], dtype=torch.int64, device=vals.device)
    out_v = torch.empty((q, k), dtype=torch.float32, device=vals.device)
    out_i = torch.empty((q, k), dtype=torch.int32, device=vals.device)
    zeros = (ctypes.c_int64 * world)(*([0] * world))
    N.check(N.load().leccr_topk_merge_peers(N.ptr(vt), N.ptr(it), world, k_in, 0, q, zeros, k, N.ptr(out_v), N.ptr(out_i),
                                            N.stream_ptr()), "leccr_topk_merge_peers")
    return out_v, out_i.to(idx.dtype)
