"""Development aid: InfoNCE forward at N = 4096 only (for ncu captures of sim_gemm_kernel<EpiLse>)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from leccr_b200 import ops, synth
cb = synth.cfg3_itc()
A, B = ops.prep(cb.image.cuda(), want_stats=False), ops.prep(cb.text.cuda(), want_stats=False)
idx, temp = cb.idx.cuda(), torch.tensor(cb.temp, device="cuda")
for _ in range(5):
    ops.infonce_forward(A, B, idx, temp)
torch.cuda.synchronize()
