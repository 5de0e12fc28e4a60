"""Bring-up checks for the CUDA kernels against torch on the same GPU (development tool).

`python tools/gpu_check.py` runs every group in its own subprocess (a trapped kernel poisons the CUDA
context) with a timeout and writes gpurun_out/check_<group>.log.  `python tools/gpu_check.py <group>`
runs one group in-process.  The committed parity tests live in tests/ and use the CPU oracle.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["gemm", "gemm_x3", "topk", "loss", "rank", "perf"]


def g_ncu():
    """Short workload for an ncu launch list: one of each hot launch."""
    import torch
    from leccr_b200 import ops, _native as N

    n, m = 5000, 25000
    img, txt = synth(n, m, 256, 3)
    img, txt = img.cuda(), txt.cuda()
    I, T = ops.prep(img, N.FMT_F16), ops.prep(txt, N.FMT_F16)
    per = m // n
    gi = ops.csr_from_lists([list(range(i * per, (i + 1) * per)) for i in range(n)], img.device)
    gt = ops.csr_from_lists([[t // per] for t in range(m)], img.device)
    for _ in range(2):
        ops.sim_topk([(I, T, gi), (T, I, gt)], k=10)
        ops.sim_topk([(I, T, None)], k=10)
        out = ops.sim_matrix(I, T)
    a, b = synth(4096, 4096, 256, 7, noise=0.5)
    a, b = a.cuda(), b.cuda()
    temp = torch.tensor(0.07, device="cuda")
    idx = torch.randint(0, 2048, (4096,), device="cuda")
    go = torch.tensor(1.0, device="cuda")
    for _ in range(2):
        A, B = ops.prep(a, N.FMT_F16), ops.prep(b, N.FMT_F16)
        o, lse2, rcnt = ops.infonce_forward(A, B, idx, temp)
        aT, bT = ops.transpose16(A), ops.transpose16(B)
        ops.infonce_backward(A, B, aT, bT, idx, temp, lse2, rcnt, 0, 512, go)
    torch.cuda.synchronize()
    return True


def synth(n, m, d, seed, noise=1.5):
    import torch

    g = torch.Generator().manual_seed(seed)
    a = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1)
    per = max(1, m // n)
    base = a[torch.arange(m) // per % n]
    b = torch.nn.functional.normalize(base + noise * 4.0 / d ** 0.5 * torch.randn(m, d, generator=g), dim=-1)
    return a, b


def g_gemm():
    import torch
    from leccr_b200 import ops, _native as N

    ok = True
    for (n, m, d, fmt) in [(128, 256, 64, N.FMT_F16), (128, 256, 256, N.FMT_F16), (256, 512, 256, N.FMT_BF16),
                           (1000, 5000, 256, N.FMT_F16), (130, 300, 256, N.FMT_F16), (77, 1000, 128, N.FMT_BF16),
                           (4096, 4096, 256, N.FMT_F16)]:
        a, b = synth(n, m, d, 1)
        a, b = a.cuda(), b.cuda()
        A, B = ops.prep(a, fmt), ops.prep(b, fmt)
        S = ops.sim_matrix(A, B)
        torch.cuda.synchronize()
        ref16 = A.t16.float() @ B.t16.float().t()
        ref32 = (a.double() @ b.double().t()).float()
        e16 = (S - ref16).abs().max().item()
        e32 = (S - ref32).abs().max().item()
        bound = (A.rn_lo.max() * B.stats[0] + A.rn_hi.max() * B.stats[1] + A.rn_lo.max() * B.stats[1]).item()
        good = e16 < 2e-5 and e32 <= bound + 1e-4
        ok &= good
        print(f"gemm n={n} m={m} d={d} fmt={fmt}: err_vs_16bit_ref={e16:.3e} err_vs_fp64={e32:.3e} "
              f"bound={bound:.3e} stats={B.stats.tolist()} {'PASS' if good else 'FAIL'}")
    return ok


def g_gemm_x3():
    import torch
    from leccr_b200 import ops, _native as N

    ok = True
    for (n, m, d, fmt) in [(1000, 5000, 256, N.FMT_F16), (1000, 1000, 256, N.FMT_BF16), (333, 777, 64, N.FMT_F16)]:
        a, b = synth(n, m, d, 2)
        a, b = a.cuda(), b.cuda()
        A = ops.prep(a, fmt, N.LAYOUT_X3_ROWS)
        B = ops.prep(b, fmt, N.LAYOUT_X3_COLS)
        S = ops.sim_matrix(A, B)
        torch.cuda.synchronize()
        ref = (a.double() @ b.double().t())
        err = (S.double() - ref).abs().max().item()
        err32 = ((a @ b.t()).double() - ref).abs().max().item()
        tol = 2e-6 if fmt == N.FMT_F16 else 5e-5
        good = err < tol
        ok &= good
        print(f"gemm_x3 n={n} m={m} d={d} fmt={fmt}: err_vs_fp64={err:.3e} (torch fp32 matmul err {err32:.3e}) "
              f"{'PASS' if good else 'FAIL'}")
    return ok


def g_topk():
    import torch
    from leccr_b200 import ops, _native as N

    ok = True
    for (n, m, d, tpc) in [(1000, 5000, 256, 0), (1000, 5000, 256, 3), (300, 700, 256, 1), (5000, 25000, 256, 0)]:
        img, txt = synth(n, m, d, 3)
        img, txt = img.cuda(), txt.cuda()
        per = m // n
        img2txt = [list(range(i * per, (i + 1) * per)) for i in range(n)]
        txt2img = [[t // per] for t in range(n * per)] + [[0] for _ in range(m - n * per)]
        I, T = ops.prep(img, N.FMT_F16), ops.prep(txt, N.FMT_F16)
        gi = ops.csr_from_lists(img2txt, img.device)
        gt = ops.csr_from_lists(txt2img, img.device)
        r_i2t, r_t2i = ops.sim_topk([(I, T, gi), (T, I, gt)], k=10, tiles_per_chunk=tpc)
        torch.cuda.synchronize()
        S = (img.double() @ txt.double().t())
        for name, res, Sx, gts in (("i2t", r_i2t, S, img2txt), ("t2i", r_t2i, S.t(), txt2img)):
            tv, ti = torch.topk(Sx, 10, dim=1)
            same = (ti == res.idx.long()).all(dim=1)
            # rows whose index lists differ must only differ by near-ties
            gap = (torch.gather(Sx, 1, res.idx.long()) - tv).abs().max().item()
            verr = (res.val.double() - torch.gather(Sx, 1, res.idx.long())).abs().max().item()
            gtt = torch.tensor([g + [g[0]] * (per - len(g)) if len(g) < per else g for g in gts] if name == "i2t"
                               else gts, device=Sx.device)
            gs = torch.gather(Sx, 1, gtt)                     # [rows, ngt]
            ranks = (Sx.unsqueeze(2) > gs.unsqueeze(1)).sum(1).min(dim=1).values if Sx.numel() < 3e7 else None
            if ranks is None:
                ranks = torch.stack([(Sx > gs[:, j:j + 1]).sum(1) for j in range(gs.shape[1])], 1).min(1).values
            exp_counts = [(ranks < c).sum().item() for c in (1, 5, 10)]
            got_counts = res.recall_counts.tolist()
            small = ranks < 10
            rank_ok = (res.rank.long()[small] == ranks[small]).all().item() and (res.rank.long()[~small] >= 10).all().item()
            good = gap < 1e-3 and verr < 1e-3 and exp_counts == got_counts and rank_ok
            ok &= good
            print(f"topk {name} n={n} m={m} tpc={tpc}: identical_rows={same.float().mean().item():.4f} "
                  f"max_score_gap={gap:.2e} val_err={verr:.2e} recall {got_counts} vs {exp_counts} "
                  f"rank_ok={rank_ok} {'PASS' if good else 'FAIL'}")
    return ok


def ref_loss(a, b, idx, temp):
    import torch
    import torch.nn.functional as F

    logits = a @ b.t() / temp
    n = a.shape[0]
    if idx is None:
        labels = torch.arange(n, device=a.device)
        return (F.cross_entropy(logits, labels) + F.cross_entropy(logits.t(), labels)) / 2
    idx = idx.view(-1, 1)
    pos = torch.eq(idx, idx.t()).to(logits.dtype)
    labels = pos / pos.sum(1, keepdim=True)
    l1 = -torch.sum(F.log_softmax(logits, dim=1) * labels, dim=1).mean()
    l2 = -torch.sum(F.log_softmax(logits.t(), dim=1) * labels, dim=1).mean()
    return (l1 + l2) / 2


def g_loss():
    import torch
    from leccr_b200 import ops, _native as N

    ok = True
    for (n, d, with_idx, rb, rc, fmt) in [(512, 256, False, 0, 512, N.FMT_F16), (4096, 256, True, 512, 512, N.FMT_F16),
                                          (4096, 256, False, 3584, 512, N.FMT_F16), (1000, 256, True, 250, 250, N.FMT_F16),
                                          (300, 64, True, 100, 50, N.FMT_BF16)]:
        g = torch.Generator().manual_seed(7)
        a = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1)
        b = torch.nn.functional.normalize(a + 0.5 * 4.0 / d ** 0.5 * torch.randn(n, d, generator=g), dim=-1)
        idx = torch.randint(0, max(1, n // 2), (n,), generator=g) if with_idx else None
        a, b = a.cuda(), b.cuda()
        idx = idx.cuda() if idx is not None else None
        temp = torch.tensor(0.07, device="cuda")
        ad, bd, td = a.double().requires_grad_(), b.double().requires_grad_(), temp.double().requires_grad_()
        loss_ref = ref_loss(ad, bd, idx, td)
        loss_ref.backward()
        A, B = ops.prep(a, fmt), ops.prep(b, fmt)
        out, lse2, rcnt = ops.infonce_forward(A, B, idx, temp)
        aT, bT = ops.transpose16(A), ops.transpose16(B)
        go = torch.tensor(1.0, device="cuda")
        dA, dB = ops.infonce_backward(A, B, aT, bT, idx, temp, lse2, rcnt, rb, rc, go)
        torch.cuda.synchronize()
        rel = abs(out[0].item() - loss_ref.item()) / abs(loss_ref.item())
        dt_rel = abs(out[1].item() - td.grad.item()) / max(1e-12, abs(td.grad.item()))
        ga, gb = ad.grad[rb:rb + rc], bd.grad[rb:rb + rc]
        ea = ((dA.double() - ga).norm() / ga.norm()).item()
        eb = ((dB.double() - gb).norm() / gb.norm()).item()
        tol_g = 2e-3 if fmt == N.FMT_F16 else 2e-2
        good = rel < 1e-3 and dt_rel < 2e-3 and ea < tol_g and eb < tol_g
        ok &= good
        print(f"loss n={n} d={d} idx={with_idx} rows=[{rb},{rb + rc}) fmt={fmt}: loss={out[0].item():.6f} "
              f"ref={loss_ref.item():.6f} rel={rel:.2e} dtemp_rel={dt_rel:.2e} dA_rel={ea:.2e} dB_rel={eb:.2e} "
              f"{'PASS' if good else 'FAIL'}")
    return ok


def g_rank():
    import torch
    from leccr_b200 import ops, _native as N

    ok = True
    n, m = 1000, 5000
    img, txt = synth(n, m, 256, 5)
    S = (img @ txt.t()).cuda().contiguous()
    per = m // n
    gi = ops.csr_from_lists([list(range(i * per, (i + 1) * per)) for i in range(n)], S.device)
    gt = ops.csr_from_lists([[t // per] for t in range(m)], S.device)
    rr = ops.rank_rows(S, *gi)
    rc = ops.rank_cols(S, *gt)
    torch.cuda.synchronize()
    gtt = torch.arange(m, device=S.device).view(n, per)
    gs = torch.gather(S, 1, gtt)
    exp_r = torch.stack([(S > gs[:, j:j + 1]).sum(1) for j in range(per)], 1).min(1).values
    gcol = torch.arange(m, device=S.device) // per
    exp_c = (S > S[gcol, torch.arange(m, device=S.device)].unsqueeze(0)).sum(0)
    g1 = (rr.long() == exp_r).all().item()
    g2 = (rc.long() == exp_c).all().item()
    print(f"rank rows {'PASS' if g1 else 'FAIL'} cols {'PASS' if g2 else 'FAIL'} counts={ops.recall_counts(rr).tolist()}")
    ok &= g1 and g2
    # double_sim fusion
    nv = 1000
    v, t = synth(nv, nv, 256, 6)
    cap = v.unsqueeze(0) + 0.1 * torch.randn(2, nv, 256)
    S = (v @ t.t()).cuda().contiguous()
    Cn = torch.stack([c @ t.t() for c in cap]).cuda().contiguous()

    def norm_score(x):
        s = -x
        s = s - torch.min(s)
        s = s / torch.max(s)
        return -s

    C = Cn.max(dim=0)[0]
    exp = 0.9 * norm_score(S) + (1. - 0.9) * norm_score(C)
    got = ops.double_sim_fuse(S.clone(), Cn, 0.9, N.FUSE_NORM)
    exp_raw = 0.8 * S + (1 - 0.8) * C
    got_raw = ops.double_sim_fuse(S.clone(), Cn, 0.8, N.FUSE_RAW)
    torch.cuda.synchronize()
    e1 = (got - exp).abs().max().item()
    e2 = (got_raw - exp_raw).abs().max().item()
    g3 = e1 < 1e-6 and e2 < 1e-6
    print(f"double_sim fuse norm err={e1:.2e} raw err={e2:.2e} bit_equal={(got == exp).float().mean().item():.4f} "
          f"{'PASS' if g3 else 'FAIL'}")
    return ok and g3


def g_perf():
    import torch
    from leccr_b200 import ops, _native as N

    def timeit(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3  # us

    for (n, m) in [(5000, 25000), (16384, 65536), (20000, 250000)]:
        img, txt = synth(n, min(m, 100000), 256, 3)
        if m > txt.shape[0]:
            txt = txt.repeat((m + txt.shape[0] - 1) // txt.shape[0], 1)[:m]
        img, txt = img.cuda(), txt.cuda()
        I, T = ops.prep(img, N.FMT_F16), ops.prep(txt, N.FMT_F16)
        for tpc in (0, 13, 25, 49, 98):
            try:
                us = timeit(lambda: ops.sim_topk([(I, T, None), (T, I, None)], k=10, tiles_per_chunk=tpc))
            except Exception as e:
                print(f"perf topk both n={n} m={m} tpc={tpc}: {e}")
                continue
            fl = 2 * 2.0 * n * m * 256
            print(f"perf topk both n={n} m={m} tpc={tpc}: {us:.1f} us  {fl / us / 1e6:.1f} TFLOP/s")
        us = timeit(lambda: ops.sim_topk([(I, T, None)], k=10))
        print(f"perf topk i2t n={n} m={m}: {us:.1f} us  {2.0 * n * m * 256 / us / 1e6:.1f} TFLOP/s")
        us = timeit(lambda: ops.sim_topk([(T, I, None)], k=10))
        print(f"perf topk t2i n={n} m={m}: {us:.1f} us  {2.0 * n * m * 256 / us / 1e6:.1f} TFLOP/s")
        if n * m <= 5000 * 25000:
            out = torch.empty((n, m), device="cuda")
            us = timeit(lambda: ops.sim_matrix(I, T, out=out))
            print(f"perf store n={n} m={m}: {us:.1f} us  {2.0 * n * m * 256 / us / 1e6:.1f} TFLOP/s")
            us = timeit(lambda: torch.matmul(I.t16, T.t16.t()))
            print(f"perf torch fp16 matmul n={n} m={m}: {us:.1f} us  {2.0 * n * m * 256 / us / 1e6:.1f} TFLOP/s")
    n = 4096
    a, b = synth(n, n, 256, 7, noise=0.5)
    a, b = a.cuda(), b.cuda()
    temp = torch.tensor(0.07, device="cuda")
    idx = torch.randint(0, n // 2, (n,), device="cuda")
    go = torch.tensor(1.0, device="cuda")

    def step():
        A, B = ops.prep(a, N.FMT_F16), ops.prep(b, N.FMT_F16)
        out, lse2, rcnt = ops.infonce_forward(A, B, idx, temp)
        aT, bT = ops.transpose16(A), ops.transpose16(B)
        return ops.infonce_backward(A, B, aT, bT, idx, temp, lse2, rcnt, 0, 512, go)

    print(f"perf infonce fwd+bwd n=4096 local 512: {timeit(step):.1f} us")
    A, B = ops.prep(a, N.FMT_F16), ops.prep(b, N.FMT_F16)
    print(f"perf infonce fwd only: {timeit(lambda: ops.infonce_forward(A, B, idx, temp)):.1f} us")
    return True


def main():
    if len(sys.argv) > 1:
        fn = globals()["g_" + sys.argv[1]]
        ok = fn()
        print("GROUP", sys.argv[1], "PASS" if ok else "FAIL")
        sys.exit(0 if ok else 1)
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    summary = []
    for g in GROUPS:
        t0 = time.time()
        log = os.path.join(out_dir, f"check_{g}.log")
        with open(log, "w") as f:
            try:
                rc = subprocess.run([sys.executable, os.path.abspath(__file__), g], stdout=f, stderr=subprocess.STDOUT,
                                    timeout=300).returncode
            except subprocess.TimeoutExpired:
                rc = "timeout"
        summary.append(f"{g}: rc={rc} ({time.time() - t0:.0f}s)")
        print(summary[-1], flush=True)
        with open(log) as f:
            print("".join(f.readlines()[-40:]), flush=True)
    with open(os.path.join(out_dir, "check_summary.txt"), "w") as f:
        f.write("\n".join(summary) + "\n")


if __name__ == "__main__":
    main()
