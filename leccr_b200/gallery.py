"""Large-gallery search (BASELINE.json configs[4]): per-query top-k of Q queries
against a G-row gallery of 16-bit embeddings, on the ranks of one NVLink node.

The reference has no such entry: it materialises `image_embeds @ text_embeds.t()` and argsorts every row on
rank 0's CPU (image_Retrieval_caption.py:151,268), which cannot be done for 1M x 100k (400 GB of scores).
This is the additive, sharded form of `fused_eval`'s top-k half.

Layout (north_star): the QUERY set is sharded over S = world / P groups of ranks, the GALLERY is row-partitioned
into P parts inside each group.  Rank r = (shard r // P, part r % P) ranks its query shard against its gallery
part with the fused tensor-core pass (the Q x G scores never reach HBM), finalize writes the local [Qs, k]
lists straight into a peer-mapped buffer, ONE cross-rank barrier, and `leccr_topk_merge_peers` pulls the P
partial lists of this rank's slice of the shard over NVLink while merging them (without peer memory: an NCCL
all-gather of the lists inside the shard's sub-group, merged by the same kernel).  P = 1 is pure query sharding
(no exchange); world = 1 is the single-GPU search.

Two entries:
  search()                       inputs resident in HBM (load_device / the buffers .gal16, .qry16)
  search_host(gallery, queries)  pinned HOST arrays: the gallery part crosses PCIe in windows on a copy stream
                                 while the tensor cores rank the windows that have arrived
                                 (leccr_sim_topk_stream with LECCR_TOPK_LONG), results come back to the host.
With `exchange_gallery=True` (default when P < world) every gallery row crosses PCIe ONCE per node: the ranks
that hold the same gallery part each upload 1/S of every window and push it to the others through peer
pointers (copy engines over NVLink), so the host link carries G / world rows per rank instead of G / P.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _native as N
from . import peer
from .sharding import shard_range


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _pick_windows(row_blocks: int, part_rows: int, sms: int = 148):
    """List of list-slot counts, one per gallery window of the host path (at most 8 slots per query row in total).
    The windows grow geometrically (_window_bounds), so only the first -- small -- one is exposed; all but the first
    use the slot count whose work items fill whole waves of the persistent grid, the first takes what is left.
    More windows are preferred as long as the first stays above 32,768 rows (below that a pass is launch- and
    threshold-warm-up-bound)."""
    def eff(s):
        items = row_blocks * s
        return items / (-(-items // sms) * sms)

    best = None
    for w in (4, 3, 2, 1):
        if w > 1 and part_rows // ((1 << w) - 1) < 32768:
            continue
        if w == 1:
            s_main = max(range(1, 9), key=lambda s: (int(min(eff(s), 0.95) / 0.05 + 1e-9), -s))
            cand = [s_main]
        else:
            top = (8 - 1) // (w - 1)
            s_main = max(range(1, top + 1), key=lambda s: (int(min(eff(s), 0.95) / 0.05 + 1e-9), -s))
            cand = [max(1, min(s_main, 8 - s_main * (w - 1)))] + [s_main] * (w - 1)
        score = (int(min(eff(s_main), 0.95) / 0.05 + 1e-9), w)
        if best is None or score > best[0]:
            best = (score, cand)
    return best[1]


def _window_bounds(part_rows: int, windows: int):
    """Row ranges of the gallery windows of the host path.  Only the FIRST window's upload is exposed (the others
    cross PCIe while the tensor cores rank the window before), and a pass over a window takes several times
    longer than its upload, so the windows grow geometrically: 1 : 2 : 4 ... -- the first one is 1 / (2^W - 1)
    of the part instead of 1 / W.  Boundaries are multiples of 256 rows (whole column tiles)."""
    if windows <= 1:
        return [(0, part_rows)]
    total = (1 << windows) - 1
    bounds, b, acc = [], 0, 0
    for w in range(windows):
        acc += 1 << w
        e = part_rows if w == windows - 1 else min(part_rows, (part_rows * acc // total + 255) // 256 * 256)
        if e > b:
            bounds.append((b, e))
        b = max(b, e)
    return bounds


class SearchHandle:
    """A search issued by `search_host_async`: `.result()` waits for it and returns the pinned host tensors
    (val [n, k], idx int32 [n, k] global gallery rows, (q0, q1)).  The tensors are reused two searches later."""

    def __init__(self, plan, lane):
        self.plan, self.lane = plan, lane

    def result(self):
        L = self.plan.lanes[self.lane]
        L["end"].synchronize()
        return L["host_val"], L["host_idx"], self.plan._result_rows()


class GallerySearchPlan:
    """Static buffers + the launch sequence of one search shape.

        plan = GallerySearchPlan(n_gallery, n_query, dim)            # collective when world > 1
        plan.load_device(gallery_part_16bit, query_shard_16bit)      # rows plan.gallery_rows / plan.query_rows
        val, idx, (q0, q1) = plan.search()                           # this rank's merged query slice, global columns
        val_h, idx_h, (q0, q1) = plan.search_host(gallery_host, queries_host)   # pinned [G, D] / [Q, D] arrays
        h = plan.search_host_async(gallery_host, queries_host); ...; val_h, idx_h, rows = h.result()

    The host path keeps TWO lanes of input buffers: a search issued while the previous one is still being ranked
    uploads (and exchanges) its windows meanwhile, so a stream of searches runs at the pace of the tensor cores.
    """

    def __init__(self, n_gallery, n_query, dim, k=10, dtype=torch.bfloat16, gallery_parts=None, windows=None,
                 exchange_gallery=None):
        if not torch.cuda.is_available():
            raise N.LeccrError("leccr_b200 has no CPU path: a CUDA device (B200) is required")
        if dtype not in (torch.bfloat16, torch.float16):
            raise N.LeccrError("the gallery is stored in a 16-bit format (bf16 / fp16)")
        if dim % 8 != 0:
            raise N.LeccrError("embedding dimension must be a multiple of 8 (TMA 16-byte rows)")
        self.lib = lib = N.load()
        N.check(lib.leccr_check_device(), "leccr_check_device")
        self.dev = dev = torch.device("cuda", torch.cuda.current_device())
        self.rank, self.world = _world()
        P = gallery_parts if gallery_parts is not None else (2 if self.world % 2 == 0 else 1)
        if P < 1 or self.world % P != 0 or P > 8:
            raise N.LeccrError("gallery_parts must divide the number of ranks")
        self.P, self.S = P, self.world // P
        self.shard, self.part = self.rank // P, self.rank % P
        self.G, self.Q, self.D, self.k = n_gallery, n_query, dim, k
        self.fmt = N.FMT_BF16 if dtype == torch.bfloat16 else N.FMT_F16
        self.dtype = dtype
        self.query_rows = shard_range(n_query, self.shard, self.S)
        self.gallery_rows = shard_range(n_gallery, self.part, P)
        qb, qe = self.query_rows
        gb, ge = self.gallery_rows
        self.Qs, self.Gp = qe - qb, ge - gb
        if self.Qs <= 0 or self.Gp <= 0:
            raise N.LeccrError("fewer queries / gallery rows than ranks")
        f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)
        esz = 2
        # ---- input buffers, two lanes.  Gallery part: private, or -- host path with exchange -- peer-mapped, so that
        # the S ranks holding the same part fill it together (each uploads 1/S of every window and pushes it to its
        # mates).  Lane 0 also serves the device-resident path (load_device / search).
        if exchange_gallery is None:
            exchange_gallery = self.S > 1
        self.xchg = self.endb = None
        gal = [None, None]
        if exchange_gallery and self.S > 1 and peer.available(dev):
            gp_max = -(-n_gallery // P)
            self.xchg = peer.get_buffer(("gallery_part", n_gallery, P, dim), gp_max * dim * esz, dev, slots=2)
            if self.xchg is not None:
                self.mates = [self.part + P * s for s in range(self.S)]
                gal = [self.xchg.local(self.xchg.slot_offset(l), (self.Gp, dim), dtype) for l in range(2)]
                if P == 1:  # no merge barrier at the end of a search: a mate must not refill a lane still being ranked
                    self.endb = peer.get_buffer(("gallery_end",), 256, dev, slots=1)
        # ---- local lists: private when there is nothing to merge, else the two slots of a peer-mapped buffer (or,
        # without peer memory, private lists all-gathered with NCCL inside the shard's sub-group)
        self.pb = None
        self.nccl_group = None
        if P > 1:
            if not (dist.is_available() and dist.is_initialized()):
                raise N.LeccrError("a partitioned gallery needs an initialised torch.distributed process group")
            if peer.available(dev):
                qs_max = -(-n_query // self.S)
                self.pb = peer.get_buffer(("gallery_search", n_query, self.S, k), qs_max * k * 8, dev)
            self.group = [self.shard * P + i for i in range(P)]
            if self.pb is not None:
                lists = []
                for slot in range(2):
                    off = self.pb.slot_offset(slot)
                    lists.append((self.pb.local(off, (self.Qs, k), torch.float32),
                                  self.pb.local(off + self.Qs * k * 4, (self.Qs, k), torch.int32), off))
            else:
                groups = [dist.new_group(ranks=[sh * P + i for i in range(P)]) for sh in range(self.S)]
                self.nccl_group = groups[self.shard]
                one = (torch.empty((self.Qs, k), **f32), torch.empty((self.Qs, k), **i32), 0)
                lists = [one, one]
                self.gath_val = torch.empty((P, self.Qs, k), **f32)
                self.gath_idx = torch.empty((P, self.Qs, k), **i32)
            self.part_offsets = (ctypes.c_int64 * P)(*[shard_range(n_gallery, i, P)[0] for i in range(P)])
            mb, me = shard_range(self.Qs, self.part, P)
            self.merge_rows = (mb, me)
            self.out_val = torch.empty((me - mb, k), **f32)
            self.out_idx = torch.empty((me - mb, k), **i32)
            self._tabs = {}
        else:
            one = (torch.empty((self.Qs, k), **f32), torch.empty((self.Qs, k), **i32), 0)
            lists = [one, one]
            self.merge_rows = (0, self.Qs)
            self.out_val, self.out_idx = one[0], one[1]
        # ---- window layout of the host path
        row_blocks = (self.Qs + 127) // 128
        if windows is None:
            subs = _pick_windows(row_blocks, self.Gp)
        else:
            if int(windows) < 1 or int(windows) > 8:
                raise N.LeccrError("1 to 8 gallery windows")
            subs = [max(1, 8 // int(windows))] * int(windows)
        self.bounds = _window_bounds(self.Gp, len(subs))
        subs = subs[len(subs) - len(self.bounds):]     # tiny parts give fewer windows: keep the main slot counts
        self.subs = subs
        W = len(self.bounds)
        sub_total = sum(subs)
        sub_begin = [sum(subs[:w]) for w in range(W)]
        self.ws_stream = torch.empty(lib.leccr_sim_topk_stream_workspace(self.Qs, sub_total), dtype=torch.uint8, device=dev)
        mb, me = self.merge_rows
        # ---- lanes: lane l = input buffers l, list slot l, windowed problems, events, pinned results
        self.lanes = []
        for l in range(2):
            L = {"qry16": torch.empty((self.Qs, dim), dtype=dtype, device=dev),
                 "gal16": gal[l], "lists": lists[l]}
            self.lanes.append(L)
        # the second lane's private gallery copy is allocated on first use of the host path (device-only users and
        # plans without exchange pay for one copy)
        if self.lanes[0]["gal16"] is None:
            self.lanes[0]["gal16"] = torch.empty((self.Gp, dim), dtype=dtype, device=dev)
        for l, L in enumerate(self.lanes):
            L["ev_q"] = torch.cuda.Event()
            L["ev_win"] = [torch.cuda.Event() for _ in self.bounds]
            L["end"] = torch.cuda.Event()
            L["busy"] = False
            L["host_val"] = torch.empty((me - mb, k), dtype=torch.float32).pin_memory()
            L["host_idx"] = torch.empty((me - mb, k), dtype=torch.int32).pin_memory()
            L["stream_calls"] = None
        self._stream_layout = (sub_begin, subs, sub_total)
        self._dev_end = torch.cuda.Event()
        self._dev_used = False
        self.calls = 0
        # ---- one-shot problems (device-resident inputs live in lane 0; the list slots alternate)
        self.probs = []
        for (lv, li, _off) in lists:
            pr = (N.TopkProblem * 1)()
            self._fill(pr[0], self.lanes[0], self.lanes[0]["gal16"].data_ptr(), self.Gp, lv, li)
            self.probs.append(pr)
        self.ws = torch.empty(lib.leccr_sim_topk_workspace(self.probs[0], 1, 0), dtype=torch.uint8, device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.launches_per_search = 2 + (2 if P > 1 else 0)  # tensor-core pass, finalize (, barrier, merge)

    # convenience views of lane 0 (the device-resident path)
    @property
    def qry16(self):
        return self.lanes[0]["qry16"]

    @property
    def gal16(self):
        return self.lanes[0]["gal16"]

    # -------------------------------------------------------------------------------------------------
    def _fill(self, p, L, cols_ptr, n_cols, lv, li):
        p.rows16, p.cols16 = L["qry16"].data_ptr(), cols_ptr
        p.ld_rows16 = p.ld_cols16 = self.D
        p.n_rows, p.n_cols = self.Qs, n_cols
        p.topk_val, p.topk_idx = lv.data_ptr(), li.data_ptr()

    def _lane_calls(self, l):
        """Windowed problems of lane l (built on first use)."""
        L = self.lanes[l]
        if L["stream_calls"] is None:
            if L["gal16"] is None:
                L["gal16"] = torch.empty((self.Gp, self.D), dtype=self.dtype, device=self.dev)
            sub_begin, subs, sub_total = self._stream_layout
            lv, li, _off = L["lists"]
            W = len(self.bounds)
            calls = []
            for w, (b, e) in enumerate(self.bounds):
                pr = (N.TopkProblem * 1)()
                so = (N.TopkStream * 1)()
                self._fill(pr[0], L, L["gal16"].data_ptr() + b * self.D * 2, e - b, lv, li)
                o = so[0]
                o.phases = N.TOPK_LONG | N.TOPK_GEMM | (N.TOPK_INIT if w == 0 else 0) | (N.TOPK_FINALIZE if w == W - 1 else 0)
                o.sub_begin, o.sub_count, o.sub_total = sub_begin[w], subs[w], sub_total
                o.col_begin = b
                o.n_cols_total = self.Gp
                o.workspace, o.workspace_bytes = self.ws_stream.data_ptr(), self.ws_stream.numel()
                calls.append((pr, so))
            L["stream_calls"] = calls
            one = (N.TopkProblem * 1)()
            self._fill(one[0], L, L["gal16"].data_ptr(), self.Gp, lv, li)
            L["one_shot"] = one
        return L["stream_calls"]

    def load_device(self, gallery_part, query_shard):
        """Device-resident inputs: rows `gallery_rows` of the gallery and `query_rows` of the queries."""
        self.lanes[0]["gal16"].copy_(gallery_part)
        self.lanes[0]["qry16"].copy_(query_shard)

    # -------------------------------------------------------------------------------------------------
    def _merge(self, slot, host_path=False):
        """Exchange step of a partitioned gallery: barrier, then pull + merge this rank's slice of the shard."""
        if self.P == 1:
            if host_path and self.endb is not None:
                self.endb.barrier()
            return
        pb = self.pb
        tabs = self._tabs.get(slot)
        lv, li, off = self.lanes[slot]["lists"]
        if pb is None:
            dist.all_gather_into_tensor(self.gath_val, lv, group=self.nccl_group)
            dist.all_gather_into_tensor(self.gath_idx, li, group=self.nccl_group)
            if tabs is None:
                tabs = (torch.tensor([self.gath_val[i].data_ptr() for i in range(self.P)], dtype=torch.int64, device=self.dev),
                        torch.tensor([self.gath_idx[i].data_ptr() for i in range(self.P)], dtype=torch.int64, device=self.dev))
                self._tabs[slot] = tabs
        else:
            pb.barrier()
            if tabs is None:
                tabs = (torch.tensor([pb.ptrs[r] + off for r in self.group], dtype=torch.int64, device=self.dev),
                        torch.tensor([pb.ptrs[r] + off + self.Qs * self.k * 4 for r in self.group], dtype=torch.int64,
                                     device=self.dev))
                self._tabs[slot] = tabs
        mb, me = self.merge_rows
        N.check(self.lib.leccr_topk_merge_peers(N.ptr(tabs[0]), N.ptr(tabs[1]), self.P, self.k, mb, me - mb,
                                                self.part_offsets, self.k, N.ptr(self.out_val), N.ptr(self.out_idx),
                                                N.stream_ptr()), "leccr_topk_merge_peers")

    def _result_rows(self):
        qb = self.query_rows[0]
        return qb + self.merge_rows[0], qb + self.merge_rows[1]

    @torch.no_grad()
    def search(self):
        """Inputs resident in HBM.  Returns (val [n, k], idx int32 [n, k] global gallery rows, (q0, q1)): the
        merged lists of queries [q0, q1) -- this rank's slice; the slices of all ranks tile the query set."""
        slot = self.calls % 2
        self.calls += 1
        N.check(self.lib.leccr_sim_topk(self.probs[slot], 1, self.D, self.fmt, self.k, 0, self.ws.data_ptr(),
                                        self.ws.numel(), N.stream_ptr()), "leccr_sim_topk")
        self._merge(slot)
        self._dev_end.record(torch.cuda.current_stream())   # a later host search must not refill lane 0 under this one
        self._dev_used = True
        if self.P == 1:
            lv, li, _ = self.lanes[slot]["lists"]
            return lv, li, self._result_rows()
        return self.out_val, self.out_idx, self._result_rows()

    def _issue_host(self, gallery_host, queries_host):
        cur = torch.cuda.current_stream()
        cs = self.copy_stream
        l = self.calls % 2
        self.calls += 1
        L = self.lanes[l]
        calls = self._lane_calls(l)
        if L["busy"]:
            L["end"].synchronize()   # the search issued two calls ago: its pinned results are about to be reused
        qb, qe = self.query_rows
        gb, _ge = self.gallery_rows
        # the lane's buffers were last read by the search issued two calls ago (same stream order on every rank; its
        # end lies behind the merge barrier every rank enters after its passes, so no mate is still reading them)
        if L["busy"]:
            cs.wait_event(L["end"])
        else:
            cs.wait_stream(cur)
        if self._dev_used:
            cs.wait_event(self._dev_end)
        with torch.cuda.stream(cs):
            L["qry16"].copy_(queries_host[qb:qe], non_blocking=True)
            L["ev_q"].record(cs)
            for w, (b, e) in enumerate(self.bounds):
                if self.xchg is None:
                    L["gal16"][b:e].copy_(gallery_host[gb + b: gb + e], non_blocking=True)
                else:
                    # my 1/S of the window over PCIe, then pushed to the mates' copies by the copy engines (NVLink)
                    sb, se = shard_range(e - b, self.shard, self.S)
                    sb, se = b + sb, b + se
                    L["gal16"][sb:se].copy_(gallery_host[gb + sb: gb + se], non_blocking=True)
                    nbytes = (se - sb) * self.D * 2
                    src = L["gal16"].data_ptr() + sb * self.D * 2
                    base = self.xchg.slot_offset(l)
                    for r in self.mates:
                        if r == self.rank:
                            continue
                        dst = self.xchg.ptrs[r] + base + sb * self.D * 2
                        _memcpy_async(dst, src, nbytes, cs.cuda_stream)
                    self.xchg.barrier()  # on the copy stream: everybody's pushes of this window have landed
                L["ev_win"][w].record(cs)
        st = cur.cuda_stream
        cur.wait_event(L["ev_q"])
        other = self.lanes[1 - l]
        if other["busy"] and not other["end"].query():
            # the previous search is still being ranked: by the time this one's pass can start its windows will
            # have arrived, so it runs as ONE pass over the whole part (no per-window launches, warm thresholds)
            cur.wait_event(L["ev_win"][-1])
            N.check(self.lib.leccr_sim_topk(L["one_shot"], 1, self.D, self.fmt, self.k, 0, self.ws.data_ptr(),
                                            self.ws.numel(), st), "leccr_sim_topk")
        else:
            for w, (pr, so) in enumerate(calls):
                cur.wait_event(L["ev_win"][w])
                N.check(self.lib.leccr_sim_topk_stream(pr, so, 1, self.D, self.fmt, self.k, st), "leccr_sim_topk_stream")
        self._merge(l, host_path=True)
        if self.P == 1:
            lv, li, _ = L["lists"]
            L["host_val"].copy_(lv, non_blocking=True)
            L["host_idx"].copy_(li, non_blocking=True)
        else:
            L["host_val"].copy_(self.out_val, non_blocking=True)
            L["host_idx"].copy_(self.out_idx, non_blocking=True)
        L["end"].record(cur)
        L["busy"] = True
        return SearchHandle(self, l)

    def _check_host(self, gallery_host, queries_host):
        g = torch.as_tensor(gallery_host)
        q = torch.as_tensor(queries_host)
        if g.shape != (self.G, self.D) or q.shape != (self.Q, self.D) or g.dtype != self.dtype or q.dtype != self.dtype:
            raise N.LeccrError("search_host needs the [G, D] gallery and [Q, D] queries in the planned 16-bit dtype")
        if g.is_cuda or q.is_cuda:
            raise N.LeccrError("search_host takes host arrays; use load_device + search for device-resident inputs")
        return g, q

    @torch.no_grad()
    def search_host_async(self, gallery_host, queries_host) -> SearchHandle:
        """Issue a search of HOST arrays and return at once; `.result()` of the handle waits for it.  Two searches
        may be in flight: the second one's windows cross PCIe (and NVLink) while the first is being ranked.  The
        host arrays must stay unchanged until the handle's result has been taken."""
        g, q = self._check_host(gallery_host, queries_host)
        return self._issue_host(g, q)

    @torch.no_grad()
    def search_host(self, gallery_host, queries_host):
        """HOST inputs: the whole [G, D] gallery and [Q, D] query arrays in (ideally pinned) host memory, 16-bit.
        Each rank uploads only what it needs; returns pinned host tensors (val, idx, (q0, q1)) of its slice."""
        return self.search_host_async(gallery_host, queries_host).result()

    @property
    def h2d_bytes(self):
        """Bytes this rank moves host -> device per search_host call."""
        g_rows = self.Gp if self.xchg is None else sum(
            shard_range(e - b, self.shard, self.S)[1] - shard_range(e - b, self.shard, self.S)[0] for b, e in self.bounds)
        return (g_rows + self.Qs) * self.D * 2

    @property
    def d2h_bytes(self):
        return self.lanes[0]["host_val"].numel() * 8


def _memcpy_async(dst: int, src: int, nbytes: int, stream: int):
    """Copy between device pointers, one of them peer-mapped: a copy-engine transfer over NVLink, no SMs."""
    N.check(N.load().leccr_memcpy_peer_async(dst, src, nbytes, stream), "leccr_memcpy_peer_async")
