"""Development measurement: e2e of the streamed plan vs the graph plan on cfg2 (pinned host inputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leccr_b200
from leccr_b200 import synth
rs = synth.cfg2_mscoco5k()
img_h, txt_h = rs.image.contiguous().pin_memory(), rs.text.contiguous().pin_memory()
gt = leccr_b200.prepare_gt(rs.txt2img, rs.img2txt, 5000, 25000)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, steps=30):
    for _ in range(3): fn()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for s in range(steps):
        flush.zero_(); ev0[s].record(); fn(); ev1[s].record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in zip(ev0, ev1)) / steps
dst = torch.empty_like(txt_h, device="cuda"); dsti = torch.empty_like(img_h, device="cuda")
def h2d(): dsti.copy_(img_h, non_blocking=True); dst.copy_(txt_h, non_blocking=True)
print(f"pure H2D of 30.72 MB: {timed(h2d):.3f} ms")
g = leccr_b200.FusedEvalPlan(5000, 25000, 256, gt=gt)
print(f"graph plan e2e: {timed(lambda: g.run(img_h, txt_h)):.3f} ms")
for W, a, b in [(3, 2, 3), ((0.36, 0.36, 0.28), 2, 3), ((0.4, 0.35, 0.25), 2, 3), ((0.38, 0.34, 0.28), 2, 2), (4, 2, 3), ((0.3, 0.3, 0.25, 0.15), 2, 3), (3, 2, 2)]:
    p = leccr_b200.StreamedEvalPlan(5000, 25000, 256, gt=gt, windows=W, img_subs=a, txt_subs=b)
    ev = p.run(img_h, txt_h)
    print(f"eager {timed(lambda: p.run(img_h, txt_h, graph=False)):.3f} ms; graph: streamed windows={[e-b for b,e in p.bounds]} img_subs={a} txt_subs={b}: {timed(lambda: p.run(img_h, txt_h)):.3f} ms  r1 {ev['txt_r1']:.2f}/{ev['img_r1']:.3f}", flush=True)
