"""Peer-memory exchange inside one NVLink / NVSwitch node.

The reference exchanges embeddings with `dist.all_gather` + `torch.cat` (models/xvlm.py:53-59) and has no
top-k exchange at all (it ranks on rank 0's CPU).  Here the exchange is done by OUR kernels through peer
pointers: the kernel that casts a rank's rows stores them straight into every rank's gathered operand
buffer (`leccr_prep_push`), and the gallery-partition merge pulls the per-query lists of all ranks while it
merges them (`leccr_topk_merge_peers`).  torch supplies only the plumbing: peer-mapped allocations
(`torch.distributed._symmetric_memory`) and the process group used for their rendezvous.

Every buffer is double-buffered (two slots used alternately) and every exchange ends with ONE cross-rank
barrier (`leccr_peer_barrier`, flag words in the same peer-mapped allocation).  That is enough: a peer can
start writing slot s again only after passing the barrier of the exchange in between, which every rank
enters (stream-ordered) after it has finished reading slot s.
"""
import os

import torch
import torch.distributed as dist

from . import _native as N

_DISABLED = os.environ.get("LECCR_PEER", "1") == "0"
_cache = {}
_warned = False


class host_memory_near_device:
    """Context manager: while it is active the calling thread runs on the CPUs next to GPU `device_index` (NVML's
    ideal affinity), so host memory allocated and first touched inside -- the pinned staging buffers of the host
    paths -- lands on the GPU's NUMA node; the previous affinity is restored on exit (the training loop, its
    autograd thread and the data loaders keep their cores).  torchrun does not place its ranks; with eight ranks
    streaming pinned memory at once, buffers that all sit on one socket share its memory controllers and the
    inter-socket link.  `.bound` is False when NVML or the container's cpuset does not allow it (nothing changes)."""

    def __init__(self, device_index: int):
        self.index = int(device_index)
        self.bound = False
        self._old = None

    def __enter__(self):
        try:
            import os

            import pynvml as nv

            self._old = os.sched_getaffinity(0)
            nv.nvmlInit()
            nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(self.index))
            self.bound = True
        except Exception:
            self.bound = False
        return self

    def __exit__(self, *exc):
        if self.bound and self._old:
            try:
                import os

                os.sched_setaffinity(0, self._old)
            except Exception:
                pass
        return False


def available(device) -> bool:
    """Peer exchange is used for NCCL process groups of CUDA ranks on one node."""
    if _DISABLED or not (dist.is_available() and dist.is_initialized()):
        return False
    if dist.get_world_size() < 2 or dist.get_world_size() > 8 or device.type != "cuda":
        return False
    try:
        return "nccl" in str(dist.get_backend()).lower()
    except Exception:  # pragma: no cover
        return False


class PeerBuffer:
    """A peer-mapped allocation of `nbytes` payload bytes x `slots` slots (2: used alternately) + a block of
    barrier flags, identical on every rank of the default group.  Collective: every rank must construct it
    at the same time."""

    FLAG_BYTES = 256

    def __init__(self, nbytes: int, device, slots: int = 2):
        import torch.distributed._symmetric_memory as symm_mem

        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.slot_bytes = (nbytes + 255) // 256 * 256
        total = self.FLAG_BYTES + slots * self.slot_bytes
        group = dist.group.WORLD
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # deprecated no-op on current torch, required on older ones
            try:
                symm_mem.enable_symm_mem_for_group(group.group_name)
            except Exception:
                pass
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert self.ptrs[self.rank] == self.buf.data_ptr()
        self.flag_table = torch.tensor(self.ptrs, dtype=torch.int64, device=device)
        self._tables = {}
        self.device = device
        self.epoch = 0
        self.calls = 0
        # everybody's flags are zero before anybody's first barrier
        torch.cuda.current_stream().synchronize()
        dist.barrier()

    def next_slot(self) -> int:
        s = self.calls & 1
        self.calls += 1
        return s

    def slot_offset(self, slot: int) -> int:
        return self.FLAG_BYTES + slot * self.slot_bytes

    def table(self, byte_offset: int) -> torch.Tensor:
        """Device array of `world` pointers: every rank's buffer base + byte_offset."""
        t = self._tables.get(byte_offset)
        if t is None:
            t = torch.tensor([p + byte_offset for p in self.ptrs], dtype=torch.int64, device=self.device)
            self._tables[byte_offset] = t
        return t

    def local(self, byte_offset: int, shape, dtype) -> torch.Tensor:
        """View of this rank's own buffer."""
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return self.buf[byte_offset: byte_offset + nbytes].view(dtype).view(shape)

    def barrier(self):
        """Epoch 0 = the kernel advances a device-resident epoch counter: no per-call argument, so sequences
        containing barriers can be captured in CUDA graphs (every rank replays the same number of barriers)."""
        N.check(N.load().leccr_peer_barrier(N.ptr(self.flag_table), self.world, self.rank, 0, N.stream_ptr()),
                "leccr_peer_barrier")


def get_buffer(key, nbytes: int, device, slots: int = 2):
    """Cached PeerBuffer per use (key); None when peer memory cannot be set up (NCCL is used instead)."""
    global _warned
    if key in _cache:
        return _cache[key]
    try:
        pb = PeerBuffer(nbytes, device, slots)
    except Exception as e:  # allocation / rendezvous unsupported on this system
        if not _warned:
            import warnings

            warnings.warn(f"leccr_b200: peer-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL")
            _warned = True
        pb = None
    _cache[key] = pb
    return pb


def itc_slot(B: int, D: int, fmt: int, device):
    """Peer-mapped slot for one contrastive exchange (leccr_itc_forward): returns
    (rows_table, idx_table, flag_table, epoch, local_slot_ptr, stat_table, local_stat_ptr) or None when peer
    exchange is not available.  Slot layout = the private buffer's: [n][2D] 16-bit rows (padded to 256 bytes),
    [n] int64 labels (padded), then the statistics of the strip forward: float lse2[2][n] | float rcnt[2][n] |
    double partial[world][4].  Collective on first use per shape; alternates two slots; epoch 0 = the barrier
    kernels keep their own epoch counter (graph-capturable)."""
    if not available(device):
        return None
    world = dist.get_world_size()
    n = B * world
    rows_bytes = (n * 2 * D * 2 + 255) // 256 * 256
    idx_bytes = (n * 8 + 255) // 256 * 256
    stat_bytes = 16 * n + 32 * world
    pb = get_buffer(("itc", B, D, fmt), rows_bytes + idx_bytes + stat_bytes, device)
    if pb is None:
        return None
    off = pb.slot_offset(pb.next_slot())
    stat_off = off + rows_bytes + idx_bytes
    return (pb.table(off), pb.table(off + rows_bytes), pb.flag_table, 0, pb.buf.data_ptr() + off,
            pb.table(stat_off), pb.buf.data_ptr() + stat_off)


def merge_topk_peers(val, idx, shard_offset: int, k: int, all_queries: bool = True, offsets=None, ranks=None):
    """Row-partitioned gallery exchange: publish this rank's [Q, k_in] local lists, barrier, then pull + merge.
    Returns (val [Q', k], idx int32 [Q', k] global columns, (q_begin, q_end)) where Q' is all queries
    (all_queries) or this rank's balanced slice.  offsets: optional list of every rank's shard_offset.
    ranks: optional sub-group (a 2-D decomposition: the ranks that hold the other gallery parts for the SAME
    query shard); lists are pulled from those ranks only, offsets then lists THEIR shard offsets, and the
    slice is balanced over them.  None when peer exchange is not available."""
    from .sharding import shard_range

    dev = val.device
    if not available(dev):
        return None
    Q, k_in = val.shape
    world, rank = dist.get_world_size(), dist.get_rank()
    pb = get_buffer(("topk", Q, k_in), Q * k_in * 8, dev)
    if pb is None:
        return None
    lib = N.load()
    slot = pb.next_slot()
    off = pb.slot_offset(slot)
    pb.local(off, (Q, k_in), torch.float32).copy_(val)
    pb.local(off + Q * k_in * 4, (Q, k_in), torch.int32).copy_(idx)
    if offsets is not None:  # every rank's first gallery row, known to the caller (no exchange, no host sync)
        host_offs = [int(o) for o in offsets]
    else:
        offs = torch.tensor([shard_offset], dtype=torch.int64, device=dev)
        all_offs = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_offs, offs)  # 8 bytes per rank (the host needs them)
        host_offs = all_offs.cpu().tolist()
    pb.barrier()
    import ctypes

    if ranks is None:
        group = list(range(world))
        vt, it = pb.table(off), pb.table(off + Q * k_in * 4)
    else:
        group = [int(r) for r in ranks]
        key = (off, tuple(group))
        tabs = pb._tables.get(key)
        if tabs is None:
            tabs = (torch.tensor([pb.ptrs[r] + off for r in group], dtype=torch.int64, device=dev),
                    torch.tensor([pb.ptrs[r] + off + Q * k_in * 4 for r in group], dtype=torch.int64, device=dev))
            pb._tables[key] = tabs
        vt, it = tabs
        if len(host_offs) != len(group):
            host_offs = [host_offs[r] for r in group]
    qb, qe = (0, Q) if all_queries else shard_range(Q, group.index(rank), len(group))
    out_v = torch.empty((qe - qb, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((qe - qb, k), dtype=torch.int32, device=dev)
    arr = (ctypes.c_int64 * len(group))(*host_offs)
    N.check(lib.leccr_topk_merge_peers(N.ptr(vt), N.ptr(it), len(group), k_in, qb, qe - qb, arr, k, N.ptr(out_v),
                                       N.ptr(out_i), N.stream_ptr()), "leccr_topk_merge_peers")
    return out_v, out_i, (qb, qe)
