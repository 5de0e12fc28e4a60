"""bench.py -- headline measurement of the hot path (BASELINE.json: queries/sec sim+top-k & contrastive
fwd+bwd us at 1/2/4/8 B200, % roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[4], the large-gallery search north_star sets the target on:
1,000,000 gallery rows x 100,000 queries, D = 256, bf16 storage, top-10 per query (synthetic, seeded, SURVEY.md
section 8d: queries = normalize(gallery[gt] + sigma * randn)).  It fits one GPU (512 MB + 51 MB), so the SAME
problem is solved at every N: strong scaling.  A "step" is one whole search = 100,000 queries.
  N = 1 : one fused tensor-core pass over the whole problem (the 400 GB score matrix never exists) + finalize.
  N > 1 : the north_star layout, leccr_b200.GallerySearchPlan: query shards x 2 gallery parts; every rank ranks
          its query shard against its gallery part, the partial top-10 lists are merged over NVLink peer memory
          (leccr_peer_barrier + leccr_topk_merge_peers).  The data path's exchange is that merge.

  value : inputs resident in HBM: tensor-core pass -> finalize (-> barrier -> merge).  K steps between CUDA
          events, barrier + synchronize on both sides, max over ranks; value = 100,000 * K / time.
  e2e   : the same search through the public API with PINNED HOST inputs (GallerySearchPlan.search_host): the
          gallery crosses PCIe in windows on a copy stream while the tensor cores rank what has arrived; at
          N > 1 every gallery row crosses PCIe once per node (each rank uploads 1/N of the gallery and pushes
          it to the ranks sharing its part over NVLink); the merged top-10 lists are copied back to the host.
  roofline : the tensor-core launches (sim_gemm_kernel<EpiTopK>) of the timed steps themselves, bracketed by CUDA
          events on their stream (leccr_profile_*, no synchronisation inside the loop), algorithmic FLOPs
          2 * Qs * Gp * D of this rank's launch, vs MEASURED_PEAKS.json.
  multi_gpu_check : parity guard inside the run (the GPU test box has one GPU): merged lists of sampled queries vs
          a single pass over the whole gallery and vs fp32 matmul + top-k; get_contrastive_loss on the real
          exchange vs the fp64 oracle on the concatenated batch.
  extras : contrastive fwd+bwd (models/xvlm.py:260-292 drop-in, real exchange) at every N; at N = 1 also
          cfg2 (MSCOCO-5K eval, last round's headline), cfg1 and cfg4 latencies.
  cpu_baseline : oracle port of the reference's CPU path (fp32 matmul + per-row np.argsort) on a bounded sample
          of the same workload (rank 0, N = 1 only).
--impl reference times that CPU path alone and prints the same line shape.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GALLERY, N_QUERY, DIM, TOPK = 1_000_000, 100_000, 256, 10
WORKLOAD = "cfg5_large_gallery_1000000_x_100000_queries_d256_bf16_top10"
METRIC = "queries/sec sim+top-k"
CONFIG = {"workload": WORKLOAD, "gallery_rows": N_GALLERY, "queries": N_QUERY, "embed_dim": DIM, "k": TOPK,
          "storage": "bf16", "l2": "inputs (512 MB gallery per pass) exceed the 126 MB L2; no flush needed"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), \
            "measured (MEASURED_PEAKS.json bf16_tflops, burst; bf16_tflops_sustained beside it)"
    except Exception:
        return 1590.0, 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"


def traffic_from_profile(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled (NVML, every 5 ms, own thread) while the timed region runs."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = False
        self._thread = None

    def _run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self._stop:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(h))
                for n, bit in names.items():
                    if mask & int(bit):
                        self.reasons.add(n)
                time.sleep(0.005)
        except Exception as e:  # report, never fail the measurement
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def start(self):
        import threading

        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop = True
        if self._thread is not None:
            self._thread.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------- CPU path (oracle port)
def cpu_search_sample(gallery32, queries32, gt, n_q, threads):
    """The reference's CPU way of ranking n_q queries against the whole gallery: fp32 score matrix with torch on
    all host threads (image_Retrieval_caption.py:151-152), then per query a full np.argsort and the top-k /
    ground-truth position on one thread (:288-295) -- oracle.gallery_eval.  Returns (seconds, queries)."""
    import numpy as np
    import torch

    from oracle import oracle

    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    ev, _val, top = oracle.gallery_eval(gallery32, queries32[:n_q], gt[:n_q], k=TOPK)
    dt = time.perf_counter() - t0
    assert top.shape == (n_q, TOPK) and np.isfinite(ev["img_r1"])
    return dt, n_q


def cpu_data(seed=1237):
    """Host fp32 copy of the workload for the CPU arm (generated on the CPU: the reference arm has no GPU code)."""
    from leccr_b200 import synth

    gal, qry, gt = synth.cfg5_gallery(N_GALLERY, 4096, seed=seed, device="cpu")  # the arm ranks a sample of queries
    return gal.float(), qry.float(), gt.tolist()


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    gal, qry, gt = cpu_data()
    n_q = 64
    for _ in range(min(args.warmup, 1)):
        cpu_search_sample(gal, qry, gt, 4, threads)
    total_t, total_q = 0.0, 0
    for s in range(args.steps):
        lo = (s * n_q) % (len(gt) - n_q)
        dt, q = cpu_search_sample(gal, qry[lo:lo + n_q], gt[lo:lo + n_q], n_q, threads)
        total_t += dt
        total_q += q
    qps = total_q / total_t
    sample = (f"per step: {n_q} of the 100,000 queries against the whole 1,000,000-row gallery: fp32 score rows on "
              f"{threads} threads + a full np.argsort per query on 1 thread (the reference's itm_eval is "
              f"single-threaded); oracle port of the reference CPU path; queries/s extrapolates linearly")
    emit({
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": CONFIG,
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ----------------------------------------------------------------------------- B200 path
def run_ours(args, rank, world, local_rank):
    os.environ.setdefault("LECCR_PEER_TIMEOUT_S", "120")  # a lost rank must end the run, not hang the box
    import torch
    import torch.distributed as dist

    import leccr_b200
    from leccr_b200 import _native as N
    from leccr_b200 import ops, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = N.load()
    N.check(lib.leccr_check_device(), "leccr_check_device")
    steps, warmup = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def all_ok(flag):
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    # ---- data: every rank generates the same seeded workload on its GPU; the host copies are pinned, allocated
    # while this thread sits on the CPUs of its GPU's NUMA node (restored afterwards)
    from leccr_b200 import peer as _peer

    gal, qry, gt = synth.cfg5_gallery(N_GALLERY, N_QUERY, device=dev)
    with _peer.host_memory_near_device(local_rank) as near:
        gal_h = torch.empty(gal.shape, dtype=gal.dtype).pin_memory()
        qry_h = torch.empty(qry.shape, dtype=qry.dtype).pin_memory()
        gal_h.copy_(gal)
        qry_h.copy_(qry)
        plan = leccr_b200.GallerySearchPlan(N_GALLERY, N_QUERY, DIM, k=TOPK, dtype=torch.bfloat16)
    numa_bound = near.bound
    gb, ge = plan.gallery_rows
    qb, qe = plan.query_rows
    plan.load_device(gal[gb:ge], qry[qb:qe])

    def device_step():
        return plan.search()

    def e2e_step():
        return plan.search_host(gal_h, qry_h)

    for _ in range(warmup):
        device_step()
    barrier()
    for _ in range(warmup):
        e2e_step()
    barrier()

    # ---- parity guard (before timing; rank-local work + one all_reduce of the verdicts)
    check = parity_guard(torch, ops, plan, gal, qry, gt, gal_h, qry_h, dev, world)
    check["all_ranks_ok"] = all_ok(check["ok"])
    plan.load_device(gal[gb:ge], qry[qb:qe])  # the host path refilled the buffers with the same data; be explicit

    # ---- timed regions
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # roofline: the tensor-core launches of the timed steps themselves, bracketed by CUDA events on their own
    # stream (recorded without synchronising, resolved after the loop)
    import ctypes

    lib.leccr_profile_enable(1)
    barrier()
    e0.record()
    for _ in range(steps):
        device_step()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    tot, cnt = ctypes.c_double(), ctypes.c_int()
    lib.leccr_profile_read(ctypes.byref(tot), ctypes.byref(cnt))
    lib.leccr_profile_enable(0)
    gemm_ms = tot.value / max(1, cnt.value)
    barrier()

    t0 = time.perf_counter()
    e0.record()
    prev = None
    for _ in range(steps):   # a stream of searches: the next one is issued before the previous result is read (two lanes)
        h = plan.search_host_async(gal_h, qry_h)
        if prev is not None:
            prev.result()
        prev = h
    prev.result()
    e1.record()
    barrier()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop()

    ms_dev, ms_e2e, gemm_ms_max, wall_e2e = max_over_ranks([ms_dev, ms_e2e, gemm_ms, wall_e2e])
    h2d = max_over_ranks([float(plan.h2d_bytes)])[0]

    # ---- contrastive leg (every N, real exchange) and, at N = 1, the other configs
    peak, peak_sus, peak_src = peaks()
    extras = {"contrastive_fwd_bwd": contrastive_leg(torch, dist, leccr_b200, synth, dev, rank, world, peak, max_over_ranks,
                                                     all_ok)}
    check["contrastive"] = extras["contrastive_fwd_bwd"].pop("check")
    check["ok"] = bool(check["all_ranks_ok"] and check["contrastive"]["all_ranks_ok"])
    if world == 1:
        del gal, qry
        torch.cuda.empty_cache()
        extras.update(extras_single_gpu(torch, leccr_b200, ops, synth, lib, dev, peak))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flops_rank = 2.0 * plan.Qs * plan.Gp * DIM          # this rank's launch (2 * N * M * D)
    flops_all = 2.0 * N_QUERY * N_GALLERY * DIM
    achieved = flops_rank / (gemm_ms_max * 1e-3) / 1e12
    step_ms = ms_dev / steps
    line = {
        "metric": METRIC, "value": N_QUERY * steps / (ms_dev * 1e-3), "unit": "queries/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": CONFIG,
        "layout": {"query_shards": plan.S, "gallery_parts": plan.P, "queries_per_rank": plan.Qs,
                   "gallery_rows_per_rank": plan.Gp,
                   "exchange": "none" if plan.P == 1 else (
                       "top-10 lists merged over NVLink peer memory (leccr_peer_barrier + leccr_topk_merge_peers)"
                       if plan.pb is not None else
                       "top-10 lists all-gathered with NCCL inside each shard's sub-group, merged by leccr_topk_merge_peers "
                       "(peer memory unavailable)"),
                   "host_path_windows": len(plan.bounds), "host_path_gallery_exchange": plan.xchg is not None},
        "e2e": {"value": N_QUERY * steps / (ms_e2e * 1e-3), "unit": "queries/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(plan.d2h_bytes),
                "ms_per_step": ms_e2e / steps, "wall_ms_per_step": wall_e2e / steps,
                "api": "leccr_b200.GallerySearchPlan.search_host_async(pinned host bf16 gallery, queries).result(): every "
                       "search uploads its inputs and returns its lists to the host; the next search is issued before the "
                       "previous result is read (two lanes of input buffers)",
                "bytes_are": "per rank (max over ranks)", "host_thread_bound_to_gpu_numa_node": numa_bound},
        "gpu_launches": plan.launches_per_search * steps,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "sim_gemm_kernel<EpiTopK<16,31,2>, kARes> (filter epilogue)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "frac_of_sustained_peak": achieved / peak_sus,
                     "traffic": traffic_from_profile("cfg5_dram_bytes_per_launch"), "peak_source": peak_src,
                     "kernel_ms": gemm_ms_max, "flops_per_launch": flops_rank,
                     "whole_step": {"tflops_all_gpus": flops_all / (step_ms * 1e-3) / 1e12,
                                    "frac_of_n_x_peak": flops_all / (step_ms * 1e-3) / 1e12 / (peak * world),
                                    "frac_of_n_x_sustained_peak": flops_all / (step_ms * 1e-3) / 1e12 / (peak_sus * world),
                                    "includes": "tensor-core pass, finalize, barrier, merge"}},
        "multi_gpu_check": check,
        "recall_check": check.get("recall"),
        "extras": extras,
    }
    if world == 1:
        threads = os.cpu_count() or 1
        g32, q32, gtl = cpu_data()
        cpu_search_sample(g32, q32, gtl, 4, threads)
        n_q = 512
        dt, q = cpu_search_sample(g32, q32, gtl, n_q, threads)
        line["cpu_baseline"] = {
            "value": q / dt, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{n_q} of the 100,000 queries against the whole 1,000,000-row gallery: fp32 score rows on {threads} "
                      f"threads + a full np.argsort per query on 1 thread (oracle port of image_Retrieval_caption.py:"
                      f"151-152,288-295), {dt:.2f} s"}
    emit(line)
    if world > 1:
        os.dup2(2, 1)  # teardown chatter stays off stdout
        dist.destroy_process_group()


def parity_guard(torch, ops, plan, gal, qry, gt, gal_h, qry_h, dev, world):
    """This rank's merged lists of sampled queries against (a) ONE pass of the same kernels over the whole
    gallery, (b) fp32 matmul + top-k of the same bf16 inputs, and the host path against the device path."""
    val, idx, (q0, q1) = plan.search()
    val, idx = val.clone(), idx.clone()
    n_s = min(256, q1 - q0)
    rows = torch.linspace(q0, q1 - 1, n_s, device=dev).long()
    # (a) single pass over the whole gallery
    Qs = ops.prep(qry[rows].contiguous(), want_stats=False)
    Gf = ops.prep(gal, want_stats=False)
    (single,) = ops.sim_topk([(Qs, Gf, None)], k=TOPK)
    same_idx = bool(torch.equal(single.idx, idx[rows - q0]))
    same_val = bool(torch.equal(single.val, val[rows - q0]))
    # (b) fp32 reference of the same inputs
    ref = qry[rows].float() @ gal.float().t()
    rv, ri = ref.topk(TOPK, dim=1)
    got_i = idx[rows - q0].long()
    true_at_got = torch.gather(ref, 1, got_i).sort(dim=1, descending=True).values
    tol_ok = bool((rv - true_at_got).abs().max() < 1e-3)   # identical up to ties inside the bf16 tolerance
    frac_same = float((got_i == ri).all(dim=1).float().mean())
    g = gt[rows]
    rec_ours = [100.0 * float((got_i[:, :c] == g[:, None]).any(dim=1).float().mean()) for c in (1, 5, 10)]
    rec_ref = [100.0 * float((ri[:, :c] == g[:, None]).any(dim=1).float().mean()) for c in (1, 5, 10)]
    # Recall@1/5/10 of ALL of this rank's queries from its merged lists (reported, rank 0's slice)
    gall = gt[q0:q1]
    rec_slice = [100.0 * float((idx.long()[:, :c] == gall[:, None]).any(dim=1).float().mean()) for c in (1, 5, 10)]
    # host path == device path
    hv, hi, (h0, h1) = plan.search_host(gal_h, qry_h)
    host_same = (h0, h1) == (q0, q1) and bool(torch.equal(hi.to(dev), idx)) and bool(torch.equal(hv.to(dev), val))
    ok = same_idx and same_val and tol_ok and host_same and rec_ours == rec_ref
    del ref
    return {"ok": ok, "sampled_queries_per_rank": n_s, "merged_equals_single_pass_idx": same_idx,
            "merged_equals_single_pass_val": same_val, "within_tolerance_of_fp32_topk": tol_ok,
            "rows_identical_to_fp32_topk": frac_same, "host_path_equals_device_path": host_same,
            "recall": {"sampled_ours_r1_r5_r10": rec_ours, "sampled_fp32_reference_r1_r5_r10": rec_ref,
                       "rank0_slice_r1_r5_r10": rec_slice},
            "ranks": world}


def contrastive_leg(torch, dist, leccr_b200, synth, dev, rank, world, peak, max_over_ranks, all_ok):
    """get_contrastive_loss (models/xvlm.py:260-292 drop-in) forward + backward through autograd with the real
    exchange, checked against the fp64 oracle on the concatenated batch.  Two shapes:
    B = 512 per rank, and the global batch fixed at 4096."""
    import types

    from oracle import oracle

    out = {}
    chk = {"ok": True}
    shapes = [("b512_per_rank", 512, 512 * world)]
    if 4096 % world == 0 and 4096 // world != 512:
        shapes.append(("global_4096", 4096 // world, 4096))
    for name, B, n in shapes:
        cb = synth.cfg3_itc(n, DIM, seed=7)
        me = types.SimpleNamespace(embed_dim=DIM, temp=torch.nn.Parameter(torch.tensor(cb.temp, device=dev)))
        a = cb.image[rank * B:(rank + 1) * B].to(dev).requires_grad_(True)
        b = cb.text[rank * B:(rank + 1) * B].to(dev).requires_grad_(True)
        idx = cb.idx[rank * B:(rank + 1) * B].to(dev)

        def step():
            a.grad = b.grad = me.temp.grad = None
            loss = leccr_b200.get_contrastive_loss(me, a, b, idx)
            loss.backward()
            return loss

        loss = step()
        rl, ra, rb, rt = oracle.contrastive_loss_and_grads(cb.image, cb.text, cb.temp, cb.idx, rank=rank, batch_size=B,
                                                           dtype=torch.float64)
        e = [abs(loss.item() - rl.item()) / abs(rl.item()), ((a.grad.cpu().double() - ra).norm() / ra.norm()).item(),
             ((b.grad.cpu().double() - rb).norm() / rb.norm()).item(), abs(me.temp.grad.item() - rt.item()) / abs(rt.item())]
        good = e[0] < 1e-3 and e[1] < 2e-3 and e[2] < 2e-3 and e[3] < 2e-3
        e = max_over_ranks(e)
        chk[name] = {"loss_rel": e[0], "dA_rel": e[1], "dB_rel": e[2], "dtemp_rel": e[3],
                     "tolerance": "loss 1e-3, grads 2e-3 relative (north_star)", "vs": "fp64 oracle, concatenated batch"}
        chk["ok"] = chk["ok"] and good
        for _ in range(5):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 30
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        wall_us = (time.perf_counter() - t0) / reps * 1e6
        us, wall_us = max_over_ranks([e0.elapsed_time(e1) / reps * 1e3, wall_us])
        flops = 2.0 * n * n * DIM + 2.0 * 2.0 * B * n * DIM   # fwd 2 N^2 D + local grads 2 * (2 B N D); no recompute
        # the same step captured once in a CUDA graph and replayed (what a training loop under torch.cuda.graphs
        # pays): device-bound time, the eager figure above is bound by the host issuing ~12 launches through autograd
        graph_us = None
        cap_ok, g = True, None
        from leccr_b200 import peer as _peer_mod

        # with the NCCL fallback (no peer memory) the step contains NCCL collectives: do not capture those here
        peer_path = world == 1 or any(k[0] == "itc" and v is not None for k, v in _peer_mod._cache.items())
        try:
            if not peer_path:
                raise RuntimeError("peer-memory exchange inactive: the step is not captured")
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                a2 = a.detach().clone().requires_grad_(True)
                b2 = b.detach().clone().requires_grad_(True)
                me2 = types.SimpleNamespace(embed_dim=DIM, temp=torch.nn.Parameter(torch.tensor(cb.temp, device=dev)))

                def step2():
                    leccr_b200.get_contrastive_loss(me2, a2, b2, idx).backward()

                for _ in range(3):
                    a2.grad = b2.grad = me2.temp.grad = None
                    step2()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            a2.grad = b2.grad = me2.temp.grad = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step2()
        except Exception as ex:  # capture is optional evidence, never fatal
            cap_ok = False
            sys.stderr.write(f"contrastive graph capture failed: {type(ex).__name__}: {str(ex)[:200]}\n")
        if all_ok(cap_ok):  # every rank replays the same number of barriers, or nobody replays
            for _ in range(3):
                g.replay()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            graph_us = max_over_ranks([e0.elapsed_time(e1) / reps * 1e3])[0]
            ga = max_over_ranks([((a2.grad - a.grad).norm() / a.grad.norm()).item()])[0]
            chk[name]["graph_replay_dA_rel_vs_eager"] = ga
            chk["ok"] = chk["ok"] and ga < 1e-5
        out[name] = {"per_rank_batch": B, "global_batch": n, "us_per_step": us, "wall_us_per_step": wall_us,
                     "graph_replay_us_per_step": graph_us,
                     "graph_replay_frac_of_peak": None if graph_us is None else flops / graph_us / 1e6 / peak,
                     "algorithmic_gflop_per_rank": flops / 1e9, "tflops_per_rank": flops / us / 1e6,
                     "frac_of_peak": flops / us / 1e6 / peak,
                     "includes": "cast + exchange (peer-memory push + barrier), forward, backward of the local rows, "
                                 "through torch.autograd; idx labels"}
    if world == 1:
        # the reference's own way on the host cores, beside it: fp32 matmul + cross entropy + autograd on
        # the concatenated 4096-row batch, rank 0's 512-row share of the gradients (oracle port of models/xvlm.py:260-292)
        cb = synth.cfg3_itc(4096, DIM, seed=7)
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        oracle.contrastive_loss_and_grads(cb.image, cb.text, cb.temp, cb.idx, rank=0, batch_size=512)
        best = float("inf")
        for _ in range(5):
            t0 = time.perf_counter()
            oracle.contrastive_loss_and_grads(cb.image, cb.text, cb.temp, cb.idx, rank=0, batch_size=512)
            best = min(best, time.perf_counter() - t0)
        out["cpu_baseline"] = {"us_per_step": best * 1e6, "cores": threads, "kind": "port",
                               "sample": "N = 4096 concatenated rows, D = 256, fp32, idx labels, forward + autograd "
                                         "backward, best of 5 after one warm-up"}
    chk["all_ranks_ok"] = all_ok(chk["ok"])
    from leccr_b200 import peer

    chk["peer_memory_paths"] = {str(k[0]): (v is not None) for k, v in peer._cache.items()}
    out["check"] = chk
    return out


def extras_single_gpu(torch, leccr_b200, ops, synth, lib, dev, peak):
    """N = 1 only: the other BASELINE.json configurations.
    cfg2: MSCOCO-5K-shaped evaluation, 5,000 x 25,000, both directions, top-10 + exact Recall (device-resident
          value, pinned-host e2e, tensor-core launch alone); cfg1 and cfg4: latencies through the public API."""
    import ctypes

    out = {}
    # ---- cfg2
    n_img, n_txt = 5000, 25000
    rs = synth.cfg2_mscoco5k()
    img_h, txt_h = rs.image.contiguous().pin_memory(), rs.text.contiguous().pin_memory()
    img_d, txt_d = img_h.to(dev), txt_h.to(dev)
    gt = leccr_b200.prepare_gt(rs.txt2img, rs.img2txt, n_img, n_txt, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    plan = leccr_b200.FusedEvalPlan(n_img, n_txt, DIM, k=TOPK, gt=gt)
    plan.img.copy_(img_d)
    plan.txt.copy_(txt_d)
    splan = leccr_b200.StreamedEvalPlan(n_img, n_txt, DIM, k=TOPK, gt=gt)

    def timed(fn, steps):
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for s in range(steps):
            flush.zero_()
            ev0[s].record()
            fn()
            ev1[s].record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in zip(ev0, ev1)) / steps

    for _ in range(3):
        plan.launch()
        splan.run(img_h, txt_h)
    ev = splan.run(img_h, txt_h)
    ev_eager = leccr_b200.fused_eval(img_h, txt_h, k=TOPK, gt=gt, return_topk=False)
    assert ev == ev_eager and plan.run(img_h, txt_h) == ev_eager, "cfg2: streamed / graph / eager paths disagree"
    ms_dev = timed(plan.launch, 20)
    ms_e2e = timed(lambda: splan.run(img_h, txt_h), 20)
    lib.leccr_profile_enable(1)
    for _ in range(10):
        flush.zero_()
        I, T = ops.prep(img_d), ops.prep(txt_d)
        ops.sim_topk([(I, T, gt[0]), (T, I, gt[1])], k=TOPK)
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_int()
    lib.leccr_profile_read(ctypes.byref(tot), ctypes.byref(cnt))
    lib.leccr_profile_enable(0)
    gemm_ms = tot.value / max(1, cnt.value)
    flops = 2.0 * 2.0 * n_img * n_txt * DIM
    q = n_img + n_txt
    out["cfg2"] = {"workload": "mscoco5k_eval_5000img_x_25000txt_d256_i2t+t2i_top10_recall", "queries_per_step": q,
                   "value_queries_per_s": q / (ms_dev * 1e-3), "ms_per_step": ms_dev,
                   "e2e_queries_per_s": q / (ms_e2e * 1e-3), "e2e_ms_per_step": ms_e2e,
                   "e2e_h2d_bytes_per_step": q * DIM * 4, "e2e_d2h_bytes_per_step": 32,
                   "l2": "flushed between steps (256 MiB write)",
                   "roofline": {"bound": "tensor", "kernel": "sim_gemm_kernel<EpiTopK> (both directions, one launch)",
                                "achieved": flops / (gemm_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                                "frac": flops / (gemm_ms * 1e-3) / 1e12 / peak, "kernel_ms": gemm_ms,
                                "flops_per_launch": flops, "traffic": traffic_from_profile("dram_bytes_per_launch")},
                   "recall_check": {k: ev[k] for k in ("txt_r1", "txt_r5", "txt_r10", "img_r1", "img_r5", "img_r10")}}
    # the same evaluation when only the Recall dict is wanted (all the reference's itm_eval returns): counting
    # epilogue, no candidate lists (leccr_sim_rank)
    rplan = leccr_b200.FusedEvalPlan(n_img, n_txt, DIM, k=TOPK, gt=gt, lists=False)
    rplan.img.copy_(img_d)
    rplan.txt.copy_(txt_d)
    assert rplan.run(img_h, txt_h) == ev_eager, "cfg2: Recall-only path disagrees"
    ms_rank = timed(rplan.launch, 20)
    lib.leccr_profile_enable(1)
    for _ in range(10):
        flush.zero_()
        I, T = ops.prep(img_d), ops.prep(txt_d)
        ops.sim_rank([(I, T, gt[0]), (T, I, gt[1])])
    torch.cuda.synchronize()
    lib.leccr_profile_read(ctypes.byref(tot), ctypes.byref(cnt))
    lib.leccr_profile_enable(0)
    rank_ms = tot.value / max(1, cnt.value)
    out["cfg2_recall_only"] = {"workload": "mscoco5k_eval_5000img_x_25000txt_d256_i2t+t2i_recall (no top-k lists: what itm_eval returns)",
                               "value_queries_per_s": q / (ms_rank * 1e-3), "ms_per_step": ms_rank,
                               "roofline": {"bound": "tensor", "kernel": "sim_gemm_kernel<EpiRank> (both directions, one launch)",
                                            "achieved": flops / (rank_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                                            "frac": flops / (rank_ms * 1e-3) / 1e12 / peak, "kernel_ms": rank_ms,
                                            "flops_per_launch": flops,
                                            "traffic": traffic_from_profile("cfg2_rank_dram_bytes_per_launch")},
                               "recall_check": "equal to cfg2.recall_check (asserted)"}
    del plan, splan, rplan, flush
    # ---- cfg1 / cfg4: microseconds per evaluation through the public API, inputs resident in HBM
    def us_per_call(fn, reps=30):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e6

    r1 = synth.cfg1_multi30k()
    i1, t1 = r1.image.to(dev), r1.text.to(dev)
    g1 = leccr_b200.prepare_gt(r1.txt2img, r1.img2txt, 1000, 5000, dev)
    p1 = leccr_b200.FusedEvalPlan(1000, 5000, DIM, k=TOPK, gt=g1)
    ev1 = p1.run(i1, t1)
    out["cfg1"] = {"workload": "multi30k_eval_1000img_x_5000txt_d256_i2t+t2i_top10_recall",
                   "us_per_eval": us_per_call(lambda: p1.run(i1, t1)),
                   "api": "FusedEvalPlan.run(device fp32): cast, fused pass, finalize, Recall counts to the host",
                   "recall_check": {k: ev1[k] for k in ("txt_r1", "txt_r5", "txt_r10", "img_r1", "img_r5", "img_r10")}}
    p1r = leccr_b200.FusedEvalPlan(1000, 5000, DIM, k=TOPK, gt=g1, lists=False)
    assert p1r.run(i1, t1) == ev1, "cfg1: Recall-only path disagrees"
    out["cfg1"]["us_per_eval_recall_only"] = us_per_call(lambda: p1r.run(i1, t1))
    r4 = synth.cfg4_msrvtt()
    i4, t4, c4 = r4.image.to(dev), r4.text.to(dev), r4.caption.to(dev)
    g4 = leccr_b200.prepare_gt(r4.txt2img, r4.img2txt, 1000, 1000, dev)

    def cfg4_eval():
        return leccr_b200.fused_eval(i4, t4, k=TOPK, gt=g4, caption_embeds=c4, alpha=0.9, fusion="norm", return_topk=False)

    ev4 = cfg4_eval()
    out["cfg4"] = {"workload": "msrvtt_double_sim_1000vid_x_1000txt_n2_alpha0.9_norm_fusion_top10_recall",
                   "us_per_eval": us_per_call(cfg4_eval),
                   "api": "fused_eval(device fp32, caption_embeds, fusion='norm'): Recall counts to the host",
                   "recall_check": {k: ev4[k] for k in ("txt_r1", "txt_r5", "txt_r10", "img_r1", "img_r5", "img_r10")}}
    torch.cuda.empty_cache()
    return out


_REAL_STDOUT = None


def emit(line: dict):
    """Print the result line on the real stdout."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    # stdout carries the one JSON line only: libraries that print to the C-level stdout (NCCL's version banner)
    # are sent to stderr until the line is ready
    sys.stdout.flush()
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
