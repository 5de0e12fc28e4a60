"""Development measurement: host vs device cost of the contrastive step through the public autograd API."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leccr_b200
from leccr_b200 import synth

for B in (512, 4096):
    cb = synth.cfg3_itc(B, 256, seed=7)
    me = types.SimpleNamespace(embed_dim=256, temp=torch.nn.Parameter(torch.tensor(cb.temp, device="cuda")))
    a = cb.image.cuda().requires_grad_(True)
    b = cb.text.cuda().requires_grad_(True)
    idx = cb.idx.cuda()

    def step():
        me.temp.grad = None; a.grad = None; b.grad = None
        l = leccr_b200.get_contrastive_loss(me, a, b, idx)
        l.backward()

    for _ in range(10): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): step()
    e1.record()
    t_host = (time.perf_counter() - t0) / 100 * 1e6
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 100 * 1e6
    print(f"B=N={B}: host issue {t_host:.1f} us/step, wall {t_all:.1f} us/step, device events {e0.elapsed_time(e1)*10:.1f} us/step")
    # forward only, host side
    t0 = time.perf_counter()
    for _ in range(100):
        l = leccr_b200.get_contrastive_loss(me, a, b, idx)
    t_f = (time.perf_counter() - t0) / 100 * 1e6
    torch.cuda.synchronize()
    print(f"   forward host issue {t_f:.1f} us")
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(10): step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
