"""Development measurement: device time of the parts of the 8 x 512 contrastive step that can be run on one GPU
(rank 0's shapes: N = 4096 gathered rows, 512 local rows), each timed back to back (device-bound)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from leccr_b200 import ops, synth

cb = synth.cfg3_itc()
a32, b32, idx = cb.image.cuda(), cb.text.cuda(), cb.idx.cuda()
temp = torch.tensor(cb.temp, device="cuda")
go = torch.tensor(1.0, device="cuda")
A, B = ops.prep(a32, want_stats=False), ops.prep(b32, want_stats=False)
o, lse2, rcnt = ops.infonce_forward(A, B, idx, temp)
aT, bT = ops.transpose16(A), ops.transpose16(B)


def t(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print("cast 512 + 512 rows (prep x2):", t(lambda: (ops.prep(a32[:512], want_stats=False), ops.prep(b32[:512], want_stats=False))))
print("full forward N = 4096 (both orientations, EpiLse + finalize):", t(lambda: ops.infonce_forward(A, B, idx, temp)))
print("transposes of both gathered operands:", t(lambda: (ops.transpose16(A), ops.transpose16(B))))
print("backward of 512 local rows (EpiGrad strips + EpiStore split-K + reduce):",
      t(lambda: ops.infonce_backward(A, B, aT, bT, idx, temp, lse2, rcnt, 0, 512, go)))
big = torch.empty(4096 * 512 * 2 + 4096 * 8, dtype=torch.uint8, device="cuda")
dst = torch.empty_like(big)
print("D2D copy of the gathered slot (4.2 MB):", t(lambda: dst.copy_(big)))
