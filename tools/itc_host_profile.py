"""Development measurement: where the host time of one get_contrastive_loss step (forward + backward through
torch.autograd) goes on one GPU: whole step, forward call alone, backward alone, and the bare C-ABI calls."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leccr_b200
from leccr_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cb = synth.cfg3_itc(B, 256, seed=7)
me = types.SimpleNamespace(embed_dim=256, temp=torch.nn.Parameter(torch.tensor(cb.temp, device="cuda")))
a = cb.image.cuda().requires_grad_(True)
b = cb.text.cuda().requires_grad_(True)
idx = cb.idx.cuda()


def step():
    a.grad = b.grad = me.temp.grad = None
    loss = leccr_b200.get_contrastive_loss(me, a, b, idx)
    loss.backward()


def timeit(fn, reps=200, sync_each=False):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
        if sync_each:
            torch.cuda.synchronize()
    t_issue = (time.perf_counter() - t0) / reps * 1e6
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / reps * 1e6
    return t_issue, t_all


print("step (issue us, issue+drain us):", timeit(step))
print("step, synchronised each (latency):", timeit(step, sync_each=True))
fw = lambda: leccr_b200.get_contrastive_loss(me, a, b, idx)
print("forward only:", timeit(fw))
with torch.no_grad():
    print("forward only, no_grad:", timeit(fw))
loss = fw()
print("backward only (retain_graph):", timeit(lambda: loss.backward(retain_graph=True)))
print("torch.empty((512, 256), cuda):", timeit(lambda: torch.empty((512, 256), device="cuda")))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):  # leaves and their AccumulateGrad nodes created on the capture stream
    a2, b2 = cb.image.cuda().requires_grad_(True), cb.text.cuda().requires_grad_(True)
    me2 = types.SimpleNamespace(embed_dim=256, temp=torch.nn.Parameter(torch.tensor(cb.temp, device="cuda")))

    def step2():
        loss = leccr_b200.get_contrastive_loss(me2, a2, b2, idx)
        loss.backward()

    for _ in range(3):
        a2.grad = b2.grad = me2.temp.grad = None
        step2()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
try:
    g = torch.cuda.CUDAGraph()
    a2.grad = b2.grad = me2.temp.grad = None
    with torch.cuda.graph(g):
        step2()
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print("device time per step (CUDA graph replay of the whole step):", e0.elapsed_time(e1) / 50 * 1e3, "us")
except Exception as ex:
    print("graph capture failed:", type(ex).__name__, str(ex)[:300])
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(200):
    step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
