"""Development measurement: cfg2 top-k launch time against tiles_per_chunk (column chunk length)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from leccr_b200 import ops, _native as N
lib = N.load()
def gemm_us(fn, reps=5):
    fn(); torch.cuda.synchronize()
    lib.leccr_profile_enable(1)
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_int()
    lib.leccr_profile_read(ctypes.byref(tot), ctypes.byref(cnt)); lib.leccr_profile_enable(0)
    return tot.value / cnt.value * 1e3
def unit(n, d): return torch.nn.functional.normalize(torch.randn(n, d, device="cuda"), dim=-1)
q, g = unit(5000, 256), unit(25000, 256)
Q, G = ops.prep(q), ops.prep(g)
for tpc in (0, 7, 10, 13, 14, 17, 20, 25, 33):
    try:
        us = gemm_us(lambda: ops.sim_topk([(Q, G, None), (G, Q, None)], k=10, tiles_per_chunk=tpc))
        print(f"both tpc={tpc}: {us:.1f} us")
    except Exception as e:
        print(f"both tpc={tpc}: {e}")
for tpc in (0, 10, 20):
    print(f"t2i only tpc={tpc}: {gemm_us(lambda: ops.sim_topk([(G, Q, None)], k=10, tiles_per_chunk=tpc)):.1f} us")
for tpc in (0, 14, 20, 25, 33):
    print(f"i2t only tpc={tpc}: {gemm_us(lambda: ops.sim_topk([(Q, G, None)], k=10, tiles_per_chunk=tpc)):.1f} us")
