"""Development aid for ncu: cfg2 (5,000 x 25,000, both directions) through the list path (leccr_sim_topk) or, with
MODE=rank, the Recall-only path (leccr_sim_rank); a few calls, nothing else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leccr_b200
from leccr_b200 import ops, synth

rs = synth.cfg2_mscoco5k()
img, txt = rs.image.cuda(), rs.text.cuda()
gt = leccr_b200.prepare_gt(rs.txt2img, rs.img2txt, 5000, 25000, torch.device("cuda"))
I, T = ops.prep(img), ops.prep(txt)
for _ in range(int(os.environ.get("REPS", 4))):
    if os.environ.get("MODE", "rank") == "rank":
        ops.sim_rank([(I, T, gt[0]), (T, I, gt[1])])
    else:
        ops.sim_topk([(I, T, gt[0]), (T, I, gt[1])], k=10)
torch.cuda.synchronize()
print("done")
