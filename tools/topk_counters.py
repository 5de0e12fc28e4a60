"""Development measurement: event counts of the top-k epilogue (chunks, hit chunks, hit groups, shrink rounds, appends)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from leccr_b200 import ops
cnt = torch.zeros(5, dtype=torch.int64, device="cuda")
os.environ["LECCR_TOPK_COUNTERS"] = hex(cnt.data_ptr())
def unit(n, d, dtype): return torch.nn.functional.normalize(torch.randn(n, d, device="cuda"), dim=-1).to(dtype)
for (n, m, dt, both) in [(5000, 25000, torch.float32, True), (12500, 1000000, torch.bfloat16, False), (100000, 125000, torch.bfloat16, False)]:
    q, g = unit(n, 256, dt), unit(m, 256, dt)
    Q, G = ops.prep(q), ops.prep(g)
    probs = [(Q, G, None), (G, Q, None)] if both else [(Q, G, None)]
    cnt.zero_(); ops.sim_topk(probs, k=10); torch.cuda.synchronize()
    c = cnt.tolist()
    rows = n + (m if both else 0)
    print(f"n={n} m={m} both={both}: warp-chunks {c[0]}, hit chunks {c[1]} ({100*c[1]/max(1,c[0]):.1f}%), hit groups {c[2]} ({c[2]/max(1,c[1]):.2f}/hit chunk), "
          f"shrink rounds {c[3]} ({c[3]/(rows/32):.1f}/warp-rowset), appends {c[4]} ({c[4]/rows:.1f}/row)")
    del q, g, Q, G
