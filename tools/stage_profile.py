"""Development measurement: per-stage device times of one fused evaluation step (LECCR_STAGE_PROFILE=1)."""
import ctypes, os, sys
os.environ["LECCR_STAGE_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leccr_b200
from leccr_b200 import ops, synth, _native as N
lib = N.load()
rs = synth.cfg2_mscoco5k()
img, txt = rs.image.cuda(), rs.text.cuda()
gt = leccr_b200.prepare_gt(rs.txt2img, rs.img2txt, 5000, 25000)
def step():
    I, T = ops.prep(img), ops.prep(txt)
    return ops.sim_topk([(I, T, gt[0]), (T, I, gt[1])], k=10)
for _ in range(3): step()
torch.cuda.synchronize()
tot, cnt = ctypes.c_double(), ctypes.c_int()
lib.leccr_profile_read(ctypes.byref(tot), ctypes.byref(cnt))   # flush marks of warm-up
print("---- one warm step", file=sys.stderr)
step(); torch.cuda.synchronize()
lib.leccr_profile_read(ctypes.byref(tot), ctypes.byref(cnt))
