"""Development measurement: top-k launch time (tensor-core launch alone) on several shapes, optionally
with the epilogue progressively disabled (LECCR_TOPK_DEBUG=1 filter only, 2 mainloop only)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from leccr_b200 import ops, _native as N

lib = N.load()

def gemm_us(fn, reps=3):
    fn(); torch.cuda.synchronize()
    lib.leccr_profile_enable(1)
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_int()
    lib.leccr_profile_read(ctypes.byref(tot), ctypes.byref(cnt)); lib.leccr_profile_enable(0)
    return tot.value / cnt.value * 1e3

def unit(n, d, dtype):
    return torch.nn.functional.normalize(torch.randn(n, d, device="cuda"), dim=-1).to(dtype)

modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["0"]
print("LECCR_TOPK_WGS", os.environ.get("LECCR_TOPK_WGS"))
SHAPES = [(5000, 25000, torch.float32, True), (20000, 250000, torch.float32, False),
          (12500, 1000000, torch.bfloat16, False), (100000, 125000, torch.bfloat16, False)]
if len(sys.argv) > 2:  # optional: comma-separated shape indices
    SHAPES = [SHAPES[int(i)] for i in sys.argv[2].split(",")]
for (n, m, dt, both) in SHAPES:
    q, g = unit(n, 256, dt), unit(m, 256, dt)
    Q, G = ops.prep(q), ops.prep(g)
    for mode in modes:
        os.environ["LECCR_TOPK_DEBUG"] = mode
        probs = [(Q, G, None), (G, Q, None)] if both else [(Q, G, None)]
        us = gemm_us(lambda: ops.sim_topk(probs, k=10))
        fl = 2.0 * n * m * 256 * len(probs)
        print(f"n={n} m={m} {str(dt)[6:]} both={both} mode={mode}: gemm {us:.1f} us  {fl/us/1e6:.1f} TFLOP/s", flush=True)
    del q, g, Q, G
    torch.cuda.empty_cache()
