"""Development aid for `ncu --set full -k regex:sim_gemm_kernel`: a few whole-problem cfg5 searches (1M gallery x
100k queries, bf16, top-10) through GallerySearchPlan.search() on one GPU, nothing else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leccr_b200
from leccr_b200 import synth

G, Q = int(os.environ.get("G", 1_000_000)), int(os.environ.get("Q", 100_000))
gal, qry, _ = synth.cfg5_gallery(G, Q, device="cuda")
plan = leccr_b200.GallerySearchPlan(G, Q, 256, k=10)
plan.load_device(gal, qry)
for _ in range(int(os.environ.get("REPS", 4))):
    plan.search()
torch.cuda.synchronize()
print("done")
