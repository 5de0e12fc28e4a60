"""bench.py -- headline measurement of the hot path (BASELINE.json: queries/sec, similarity + top-k).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[1], the MSCOCO-5K-shaped evaluation: 5,000 images x
25,000 captions, D = 256 fp32 embeddings (synthetic, seeded), both ranking directions, top-10 and exact
Recall@1/5/10.  A "step" is one whole evaluation of one such set = 30,000 queries.
N > 1 (torchrun, one rank per GPU): every rank evaluates its own independent set (the reference evaluates
one language / split after another, image_Retrieval_caption.py:452-459) -- weak scaling, no data-path
collective; value = all ranks' queries / max-over-ranks device time.

  value : device-resident inputs (fp32 embeddings already in HBM): cast -> fused tensor-core pass ->
          finalize, timed per step with CUDA events on the launch stream, L2 flushed between steps.
  e2e   : the same through the public API leccr_b200.StreamedEvalPlan.run with PINNED HOST inputs: H2D of the
          embeddings (in windows, overlapped with the tensor-core passes) and D2H of the Recall counts inside
          the timed region.
  roofline : the tensor-core launch (sim_gemm_kernel<EpiTopK>) timed alone with CUDA events on its
          stream (leccr_profile_*), algorithmic FLOPs 2*N*M*D per direction, vs MEASURED_PEAKS.json.
  cpu_baseline : the oracle port of the reference's CPU path (torch matmul + per-row np.argsort) on this
          box's host cores, on a bounded sample of the same workload (rank 0, N = 1 only).
--impl reference times that CPU path alone and prints the same line shape.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_IMG, PER_IMG, DIM, TOPK = 5000, 5, 256, 10
N_TXT = N_IMG * PER_IMG
QUERIES_PER_STEP = N_IMG + N_TXT
WORKLOAD = "mscoco5k_eval_5000img_x_25000txt_d256_i2t+t2i_top10_recall"
METRIC = "queries/sec sim+top-k"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    except Exception:
        return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"


def traffic_from_profile():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
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
def cpu_path_sample(rs, frac, threads):
    """The reference's CPU evaluation (image_Retrieval_caption.py:151-163 + :261-295) on a bounded sample:
    the full score matrix (torch fp32 matmul, all host threads) and np.argsort ranking of `frac` of the
    rows of each direction (single host thread, as in the reference).  Returns (seconds, queries)."""
    import numpy as np
    import torch

    from oracle import oracle

    torch.set_num_threads(threads)
    n_i = max(1, int(N_IMG * frac))
    n_t = max(1, int(N_TXT * frac))
    t0 = time.perf_counter()
    i2t, t2i = oracle.score_matrices(rs.image, rs.text)
    sub_i2t = i2t[:n_i]
    sub_t2i = t2i[:n_t]
    ev = oracle.itm_eval(sub_i2t, sub_t2i, {t: rs.txt2img[t] for t in range(n_t)},
                         {i: rs.img2txt[i] for i in range(n_i)})
    dt = time.perf_counter() - t0
    assert np.isfinite(ev["r_mean"])
    return dt, n_i + n_t


def run_reference(args, rank):
    if rank != 0:
        return
    from leccr_b200 import synth

    threads = os.cpu_count() or 1
    rs = synth.cfg2_mscoco5k()
    frac = 0.2
    for _ in range(min(args.warmup, 1)):
        cpu_path_sample(rs, 0.02, threads)
    total_t, total_q = 0.0, 0
    for _ in range(args.steps):
        dt, q = cpu_path_sample(rs, frac, threads)
        total_t += dt
        total_q += q
    qps = total_q / total_t
    sample = (f"per step: full {N_IMG}x{N_TXT} fp32 score matrix on {threads} threads + np.argsort ranking of "
              f"{int(frac * 100)}% of the rows of each direction on 1 thread (the reference's itm_eval is "
              f"single-threaded); oracle port of the reference CPU path")
    emit({
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_fraction_ranked": frac},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ----------------------------------------------------------------------------- B200 path
def run_ours(args, rank, world, local_rank):
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

    rs = synth.retrieval_set(N_IMG, PER_IMG, DIM, seed=1235 + rank)  # rank 0 == synth.cfg2_mscoco5k()
    img_h = rs.image.contiguous().pin_memory()
    txt_h = rs.text.contiguous().pin_memory()
    img_d, txt_d = img_h.to(dev), txt_h.to(dev)
    gt = leccr_b200.prepare_gt(rs.txt2img, rs.img2txt, N_IMG, N_TXT, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # The step is the product's repeated-evaluation path: leccr_b200.FusedEvalPlan (static buffers + the
    # cast / tensor-core / finalize launches captured in one CUDA graph).
    plan = leccr_b200.FusedEvalPlan(N_IMG, N_TXT, DIM, k=TOPK, gt=gt)
    plan.img.copy_(img_d)
    plan.txt.copy_(txt_d)

    def device_step():       # inputs already resident in HBM
        plan.launch()

    # e2e: the product's host-input path, leccr_b200.StreamedEvalPlan: the text set crosses PCIe in windows
    # on a copy stream while the tensor cores rank what has arrived; one CUDA graph per pair of pinned buffers.
    splan = leccr_b200.StreamedEvalPlan(N_IMG, N_TXT, DIM, k=TOPK, gt=gt)

    def e2e_step():          # pinned host inputs -> windowed H2D overlapped with the passes -> D2H of the counts
        return splan.run(img_h, txt_h)

    def eager_step():        # the same launches issued one by one (roofline leg: per-launch events)
        I, T = ops.prep(img_d), ops.prep(txt_d)
        return ops.sim_topk([(I, T, gt[0]), (T, I, gt[1])], k=TOPK)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """Sum of per-step device times (CUDA events on the current stream), L2 flushed between steps."""
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for s in range(steps):
            flush.zero_()
            ev0[s].record()
            fn()
            ev1[s].record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))

    for _ in range(max(args.warmup, 3)):
        device_step()
        e2e_step()
    # correctness guard: the step must reproduce the reference's Recall (rank 0's set is cfg2)
    ev = e2e_step()
    ev_eager = leccr_b200.fused_eval(img_h, txt_h, k=TOPK, gt=gt, return_topk=False)
    assert ev == ev_eager, "streamed plan and eager path disagree"
    assert plan.run(img_h, txt_h) == ev_eager, "graph replay and eager path disagree"
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev = timed(device_step, args.steps)
    barrier()
    ms_e2e = timed(e2e_step, args.steps)
    barrier()
    clocks = sampler.stop()

    # roofline leg: the tensor-core launch alone, CUDA events on its stream
    lib.leccr_profile_enable(1)
    for _ in range(min(args.steps, 20)):
        flush.zero_()
        eager_step()
    torch.cuda.synchronize()
    import ctypes

    tot = ctypes.c_double()
    cnt = ctypes.c_int()
    lib.leccr_profile_read(ctypes.byref(tot), ctypes.byref(cnt))
    lib.leccr_profile_enable(0)
    gemm_ms = tot.value / max(1, cnt.value)

    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    flops = 2.0 * 2.0 * N_IMG * N_TXT * DIM  # both directions, 2*N*M*D each
    achieved = flops / (gemm_ms * 1e-3) / 1e12
    total_q = QUERIES_PER_STEP * world * args.steps
    line = {
        "metric": METRIC, "value": total_q / (ms_dev * 1e-3), "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu": "one independent evaluation set per rank", "k": TOPK,
                   "embed_dim": DIM, "l2": "flushed between steps (256 MiB write)", "operands": "fp32 -> fp16 tensor-core operands, fp32 accumulate, exact fp32 re-check for Recall"},
        "e2e": {"value": total_q / (ms_e2e * 1e-3), "unit": "queries/s",
                "h2d_bytes_per_step": QUERIES_PER_STEP * DIM * 4, "d2h_bytes_per_step": 32,
                "ms_per_step": ms_e2e / args.steps, "api": "leccr_b200.StreamedEvalPlan.run(pinned host fp32)",
                "gpu_launches_per_step": 15},
        "gpu_launches": 4 * args.steps,  # cast, tensor-core pass, finalize, rank_post (each for both directions)
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "sim_gemm_kernel<EpiTopK<16>>", "achieved": achieved,
                     "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic_from_profile(),
                     "peak_source": peak_src, "kernel_ms": gemm_ms, "flops_per_launch": flops},
        "recall_check": {k: ev[k] for k in ("txt_r1", "txt_r5", "txt_r10", "img_r1", "img_r5", "img_r10")},
    }
    if world == 1:
        line["extras"] = extras_single_gpu(torch, ops, synth, lib, dev, peak)
        threads = os.cpu_count() or 1
        frac = 0.2
        cpu_path_sample(rs, 0.02, threads)
        dt, q = cpu_path_sample(rs, frac, threads)
        line["cpu_baseline"] = {
            "value": q / dt, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"1 step: full {N_IMG}x{N_TXT} fp32 score matrix on {threads} threads + np.argsort ranking "
                      f"of {int(frac * 100)}% of the rows of each direction on 1 thread (oracle port of "
                      f"image_Retrieval_caption.py:151-163,261-295), {dt:.2f} s"}
    emit(line)
    if world > 1:
        os.dup2(2, 1)  # teardown chatter stays off stdout
        dist.destroy_process_group()


def extras_single_gpu(torch, ops, synth, lib, dev, peak):
    """Secondary measurements reported beside the headline (not part of the driver's contract):
    cfg3 (contrastive fwd+bwd at global batch 4096, this GPU playing rank 0 of 8) and the per-GPU share
    of cfg5 under 8-way query sharding (12,500 bf16 queries x 1,000,000 bf16 gallery rows, top-10)."""
    import ctypes

    import torch.nn.functional as F

    from leccr_b200 import _native as N

    out = {}
    # ---- cfg3: operands of all 4096 rows are "gathered" already; local rows [0, 512)
    cb = synth.cfg3_itc()
    a32, b32, idx = cb.image.to(dev), cb.text.to(dev), cb.idx.to(dev)
    temp = torch.tensor(cb.temp, device=dev)
    go = torch.tensor(1.0, device=dev)

    def itc_step():
        A, B = ops.prep(a32, want_stats=False), ops.prep(b32, want_stats=False)
        o, lse2, rcnt = ops.infonce_forward(A, B, idx, temp)
        aT, bT = ops.transpose16(A), ops.transpose16(B)
        return ops.infonce_backward(A, B, aT, bT, idx, temp, lse2, rcnt, 0, 512, go)

    for _ in range(5):
        itc_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        itc_step()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    flops = 2.0 * 2 * 4096 * 4096 * 256 + 2.0 * 2 * 2 * 512 * 4096 * 256  # fwd both orientations + strips + grads
    # the reference's CPU path for the same call (oracle port of models/xvlm.py:260-292 + autograd), all host threads
    import time as _time

    from oracle import oracle as _oracle

    torch.set_num_threads(os.cpu_count() or 1)
    _oracle.contrastive_loss_and_grads(cb.image, cb.text, cb.temp, cb.idx, rank=0, batch_size=512)
    cpu_ts = []
    for _ in range(3):
        t0 = _time.perf_counter()
        _oracle.contrastive_loss_and_grads(cb.image, cb.text, cb.temp, cb.idx, rank=0, batch_size=512)
        cpu_ts.append(_time.perf_counter() - t0)
    out["contrastive_cpu_baseline"] = {"us_per_step": min(cpu_ts) * 1e6, "cores": os.cpu_count() or 1, "kind": "port",
                                       "sample": "full cfg3 call, N = 4096, fp32, fwd + autograd bwd, best of 3"}
    out["contrastive_fwd_bwd"] = {"config": "cfg3: global batch 4096 (8 x 512), D=256, idx labels, rank 0's rows",
                                  "us_per_step": us, "tflops": flops / us / 1e6, "frac_of_peak": flops / us / 1e6 / peak,
                                  "includes": "fp32->fp16 cast of 2 x 4096 rows, forward, transposes, backward of 512 local rows"}
    # ---- cfg5 per-GPU share (query sharding over 8 GPUs)
    g = torch.Generator(device=dev).manual_seed(1237)
    gal = F.normalize(torch.randn(1_000_000, 256, device=dev, generator=g), dim=-1).to(torch.bfloat16)
    qry = F.normalize(torch.randn(12_500, 256, device=dev, generator=g), dim=-1).to(torch.bfloat16)
    Q, G = ops.prep(qry), ops.prep(gal)
    for _ in range(2):
        ops.sim_topk([(Q, G, None)], k=10)
    torch.cuda.synchronize()
    lib.leccr_profile_enable(1)
    for _ in range(3):
        ops.sim_topk([(Q, G, None)], k=10)
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_int()
    lib.leccr_profile_read(ctypes.byref(tot), ctypes.byref(cnt))
    lib.leccr_profile_enable(0)
    ms = tot.value / max(1, cnt.value)
    fl = 2.0 * 12_500 * 1_000_000 * 256
    out["cfg5_per_gpu_share"] = {"config": "12,500 bf16 queries x 1,000,000 bf16 gallery rows, D=256, top-10 (1/8 of cfg5's queries)",
                                 "kernel_ms": ms, "tflops": fl / ms / 1e9, "frac_of_peak": fl / ms / 1e9 / peak,
                                 "queries_per_s_per_gpu": 12_500 / (ms * 1e-3)}
    del gal, qry, Q, G
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
    ap.add_argument("--steps", type=int, default=50)
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
