"""Development measurement: the Recall-only launch (sim_gemm_kernel<EpiRank>) on cfg2 with the epilogue progressively
disabled (LECCR_RANK_DEBUG: 1 empty band, 2 mainloop only), the row block streamed (LECCR_RANK_ARES=0) and other
chunk lengths (LECCR_RANK_TPC); stage times of the whole call."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leccr_b200
from leccr_b200 import ops, synth, _native as N

lib = N.load()
rs = synth.cfg2_mscoco5k()
img, txt = rs.image.cuda(), rs.text.cuda()
gt = leccr_b200.prepare_gt(rs.txt2img, rs.img2txt, 5000, 25000, torch.device("cuda"))
I, T = ops.prep(img), ops.prep(txt)


def run():
    return ops.sim_rank([(I, T, gt[0]), (T, I, gt[1])])


def gemm_us(reps=5):
    run(); torch.cuda.synchronize()
    lib.leccr_profile_enable(1)
    for _ in range(reps):
        run()
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_int()
    lib.leccr_profile_read(ctypes.byref(tot), ctypes.byref(cnt)); lib.leccr_profile_enable(0)
    return tot.value / cnt.value * 1e3


def call_us(reps=10):
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for env in ({}, {"LECCR_RANK_DEBUG": "1"}, {"LECCR_RANK_DEBUG": "2"}, {"LECCR_RANK_TPC": "4"}, {"LECCR_RANK_TPC": "13"},
            {"LECCR_RANK_TPC": "20"}):
    for k in ("LECCR_RANK_DEBUG", "LECCR_RANK_TPC"):
        os.environ.pop(k, None)
    os.environ.update(env)
    print(env, f"gemm {gemm_us():.1f} us, whole call {call_us():.1f} us", flush=True)
