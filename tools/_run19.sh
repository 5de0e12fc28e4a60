python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python - <<'PY'
import sys, os, ctypes
sys.path.insert(0, os.getcwd())
import torch
from leccr_b200 import ops, synth, _native as N
lib = N.load()
cb = synth.cfg3_itc()
a32, b32, idx = cb.image.cuda(), cb.text.cuda(), cb.idx.cuda()
temp = torch.tensor(cb.temp, device="cuda"); go = torch.tensor(1.0, device="cuda")
A, B = ops.prep(a32, want_stats=False), ops.prep(b32, want_stats=False)
def fwd(): return ops.infonce_forward(A, B, idx, temp)
o, lse2, rcnt = fwd()
aT, bT = ops.transpose16(A), ops.transpose16(B)
def bwd(): return ops.infonce_backward(A, B, aT, bT, idx, temp, lse2, rcnt, 0, 512, go)
for name, fn in (("fwd", fwd), ("bwd", bwd)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    lib.leccr_profile_enable(1)
    for _ in range(10): fn()
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_int()
    lib.leccr_profile_read(ctypes.byref(tot), ctypes.byref(cnt)); lib.leccr_profile_enable(0)
    print(f"infonce {name}: tensor-core launches {cnt.value//10} per call, {tot.value/10*1e3:.1f} us per call")
PY
