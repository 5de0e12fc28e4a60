import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from leccr_b200 import _native as N
lib = N.load()
def timeit(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for n in (5000, 25000, 100000):
    x = torch.nn.functional.normalize(torch.randn(n, 256, device="cuda"), dim=-1)
    dst = torch.empty(n, 256, dtype=torch.float16, device="cuda")
    rn = torch.empty(n, device="cuda"); rl = torch.empty(n, device="cuda"); st = torch.zeros(4, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    f_all = lambda: lib.leccr_prep(x.data_ptr(), n, 256, 256, 0, 0, 0, dst.data_ptr(), 256, rn.data_ptr(), rl.data_ptr(), st.data_ptr(), s)
    f_nostat = lambda: lib.leccr_prep(x.data_ptr(), n, 256, 256, 0, 0, 0, dst.data_ptr(), 256, rn.data_ptr(), rl.data_ptr(), None, s)
    f_none = lambda: lib.leccr_prep(x.data_ptr(), n, 256, 256, 0, 0, 0, dst.data_ptr(), 256, None, None, None, s)
    print(n, "prep all %.1f us, no stats %.1f us, no norms %.1f us, torch .half() %.1f us" % (timeit(f_all), timeit(f_nostat), timeit(f_none), timeit(lambda: x.half())))
