"""Multi-GPU check (torchrun, NCCL): the contrastive drop-in on W ranks against the oracle on the
concatenated batch, and the AllGather drop-in.  Development / gpu-box tool."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
os.environ.setdefault("LECCR_PEER_TIMEOUT_S", "60")
import leccr_b200
from leccr_b200 import synth
from oracle import oracle

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = 512
cb = synth.cfg3_itc(B * world, 256, seed=7)
ok = True
for with_idx in (False, True):
    me = types.SimpleNamespace(embed_dim=256, temp=torch.nn.Parameter(torch.tensor(cb.temp, device="cuda")))
    a = cb.image[rank * B:(rank + 1) * B].cuda().requires_grad_(True)
    b = cb.text[rank * B:(rank + 1) * B].cuda().requires_grad_(True)
    idx = cb.idx[rank * B:(rank + 1) * B].cuda() if with_idx else None
    loss = leccr_b200.get_contrastive_loss(me, a, b, idx)
    loss.backward()
    rl, ra, rb, rt = oracle.contrastive_loss_and_grads(cb.image, cb.text, cb.temp, cb.idx if with_idx else None,
                                                       rank=rank, batch_size=B, dtype=torch.float64)
    e = [abs(loss.item() - rl.item()) / abs(rl.item()), ((a.grad.cpu().double() - ra).norm() / ra.norm()).item(),
         ((b.grad.cpu().double() - rb).norm() / rb.norm()).item(), abs(me.temp.grad.item() - rt.item()) / abs(rt.item())]
    good = e[0] < 1e-3 and e[1] < 2e-3 and e[2] < 2e-3 and e[3] < 2e-3
    ok &= good
    print(f"rank {rank}/{world} idx={with_idx}: loss {loss.item():.6f} rel {e[0]:.1e} dA {e[1]:.1e} dB {e[2]:.1e} dtemp {e[3]:.1e} {'PASS' if good else 'FAIL'}", flush=True)
x = torch.full((3, 4), float(rank), device="cuda", requires_grad=True)
g = leccr_b200.allgather(x, rank, world)
(g * torch.arange(g.numel(), device="cuda").view_as(g)).sum().backward()
exp = torch.arange(world, device="cuda").repeat_interleave(3).view(-1, 1).expand(-1, 4).float()
ok &= bool(torch.equal(g.detach(), exp)) and bool(torch.equal(x.grad, torch.arange(world * 12, device="cuda").view(-1, 4)[rank * 3:(rank + 1) * 3].float()))
# ---- sharded evaluation against the reference's golden dict for cfg1 (queries split over the ranks)
import numpy as np
gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "baseline_configs.npz"))
rs = synth.cfg1_multi30k()
ev, _ = leccr_b200.fused_eval_sharded(rs.image, rs.text, rs.txt2img, rs.img2txt)
ev_ok = all(float(ev[k_]) == float(gold[f"cfg1_ev_{k_}"]) for k_ in leccr_b200.evaluation.EVAL_KEYS)
ok &= ev_ok
# ---- row-partitioned gallery + NCCL merge against a single-GPU pass over the whole gallery
gal, qry, _gt = synth.cfg5_gallery(40000, 512, device="cuda")
from leccr_b200 import sharding, ops as _ops
b0, e0 = sharding.shard_range(40000, rank, world)
mv, mi = leccr_b200.topk_gallery_sharded(qry, gal[b0:e0], b0, k=10)
full, = _ops.sim_topk([(_ops.prep(qry), _ops.prep(gal), None)], k=10)
merge_ok = bool(torch.equal(mi, full.idx.long())) and bool(torch.equal(mv, full.val))
ok &= merge_ok
from leccr_b200 import peer as _peer
paths = {str(k_[0]): (v_ is not None) for k_, v_ in _peer._cache.items()}
print(f"rank {rank}: sharded eval == golden {ev_ok}; gallery-partition merge == single pass {merge_ok}; peer-memory paths used: {paths}", flush=True)
if os.environ.get("LECCR_PEER", "1") != "0":
    ok &= paths.get("itc", False) and paths.get("topk", False)  # on a B200 NVLink box the peer kernels must be what ran
# ---- dstl_loss and the caption contrastive loss through their drop-ins on W ranks, vs the oracle on the gathered batch
gd = torch.Generator().manual_seed(5)
nrm = torch.nn.functional.normalize
Nd, nq = B * world, 2
d_img = nrm(torch.randn(Nd, 256, generator=gd), dim=-1)
d_ts = nrm(d_img + 0.4 * torch.randn(Nd, 256, generator=gd), dim=-1)
d_tt = nrm(d_img + 0.4 * torch.randn(Nd, 256, generator=gd), dim=-1)
d_cap = d_ts[None] + 0.1 * torch.randn(nq, Nd, 256, generator=gd)
sl = slice(rank * B, (rank + 1) * B)
im_l = d_img[sl].cuda().requires_grad_(True)
tt_l = d_tt[sl].cuda().requires_grad_(True)
dl = leccr_b200.dstl_loss(types.SimpleNamespace(), im_l, d_cap[:, sl].cuda(), d_ts[sl].cuda(), tt_l, None, alpha=0.8)
dl.backward()
w_l, w_dim, w_dtt = oracle.dstl_loss_and_grads(d_img, d_cap, d_ts, d_tt, 0.8, rank=rank, batch_size=B, dtype=torch.float64)
e = [abs(dl.item() - w_l.item()) / abs(w_l.item()), ((im_l.grad.cpu().double() - w_dim).norm() / w_dim.norm()).item(),
     ((tt_l.grad.cpu().double() - w_dtt).norm() / w_dtt.norm()).item()]
dstl_ok = e[0] < 1e-3 and e[1] < 3e-3 and e[2] < 3e-3
ok &= dstl_ok
print(f"rank {rank}: dstl_loss {dl.item():.6f} rel {e[0]:.1e} dimage {e[1]:.1e} dtext_t {e[2]:.1e} {'PASS' if dstl_ok else 'FAIL'}", flush=True)
# ---- caption_vision_loss: pooled drop-in on W ranks vs the oracle's token-level restatement on the gathered batch
gc = torch.Generator().manual_seed(9)
Nc, cn, vn, dc = 64 * world, 2, 5, 64
c_img = torch.randn(Nc, vn, dc, generator=gc)
c_cap = c_img.mean(1)[None] * 0.5 + torch.randn(cn, Nc, dc, generator=gc)
c_idx = torch.randint(0, Nc // 2, (Nc,), generator=gc)
Wc, bc = torch.randn(dc, dc, generator=gc) / dc ** 0.5, 0.1 * torch.randn(dc, generator=gc)
Wv, bv = torch.randn(dc, dc, generator=gc) / dc ** 0.5, 0.1 * torch.randn(dc, generator=gc)
me_c = types.SimpleNamespace(cproj=torch.nn.Linear(dc, dc).cuda(), vproj=torch.nn.Linear(dc, dc).cuda())
with torch.no_grad():
    me_c.cproj.weight.copy_(Wc); me_c.cproj.bias.copy_(bc); me_c.vproj.weight.copy_(Wv); me_c.vproj.bias.copy_(bv)
slc = slice(rank * 64, (rank + 1) * 64)
ci = c_img[slc].cuda().requires_grad_(True)
cc = c_cap[:, slc].cuda().requires_grad_(True)
cl = leccr_b200.caption_vision_loss(me_c, cc, ci, c_idx[slc].cuda())
cl.backward()
o_l, o_dim, o_dcp, o_dwc, o_dwv = oracle.caption_vision_loss_and_grads(c_cap, c_img, c_idx, Wc, bc, Wv, bv, rank, 64, dtype=torch.float64)
rel = lambda a, b: ((a.cpu().double() - b).norm() / b.norm()).item()
ec = [abs(cl.item() - o_l.item()) / abs(o_l.item()), rel(ci.grad, o_dim), rel(cc.grad, o_dcp), rel(me_c.cproj.weight.grad, o_dwc),
      rel(me_c.vproj.weight.grad, o_dwv)]
cv_ok = ec[0] < 1e-3 and max(ec[1:]) < 5e-3
ok &= cv_ok
print(f"rank {rank}: caption_vision_loss {cl.item():.6f} rel {ec[0]:.1e} dimage {ec[1]:.1e} dcaption {ec[2]:.1e} dWc {ec[3]:.1e} dWv {ec[4]:.1e} {'PASS' if cv_ok else 'FAIL'}", flush=True)
# ---- the C-ABI NCCL entries (hosts without torch.distributed): communicator from a broadcast unique id,
# all-gather of this rank's top-k lists, merge over a table of pointers into the gathered buffer
import ctypes
from leccr_b200 import _native as _N
_lib = _N.load()
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    buf = ctypes.create_string_buffer(128)
    _N.check(_lib.leccr_comm_unique_id(buf), "leccr_comm_unique_id")
    uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
uid_d = uid.cuda()
dist.broadcast(uid_d, 0)
idbuf = ctypes.create_string_buffer(bytes(uid_d.cpu().tolist()), 128)
comm = ctypes.c_void_p()
_N.check(_lib.leccr_comm_init(idbuf, rank, world, ctypes.byref(comm)), "leccr_comm_init")
local_res, = _ops.sim_topk([(_ops.prep(qry), _ops.prep(gal[b0:e0]), None)], k=10)
Qn = qry.shape[0]
send = torch.cat([local_res.val.reshape(-1).view(torch.int32), local_res.idx.reshape(-1)]).contiguous()   # [Q*k vals | Q*k idx]
recv = torch.empty(world * send.numel(), dtype=torch.int32, device="cuda")
_N.check(_lib.leccr_allgather(comm, _N.ptr(send), _N.ptr(recv), send.numel() * 4, _N.stream_ptr()), "leccr_allgather")
per = send.numel() * 4
vt = torch.tensor([recv.data_ptr() + r * per for r in range(world)], dtype=torch.int64, device="cuda")
it = torch.tensor([recv.data_ptr() + r * per + Qn * 10 * 4 for r in range(world)], dtype=torch.int64, device="cuda")
offs = (ctypes.c_int64 * world)(*[sharding.shard_range(40000, r, world)[0] for r in range(world)])
ov = torch.empty((Qn, 10), dtype=torch.float32, device="cuda")
oi = torch.empty((Qn, 10), dtype=torch.int32, device="cuda")
_N.check(_lib.leccr_topk_merge_peers(_N.ptr(vt), _N.ptr(it), world, 10, 0, Qn, offs, 10, _N.ptr(ov), _N.ptr(oi),
                                     _N.stream_ptr()), "leccr_topk_merge_peers")
nccl_ok = bool(torch.equal(oi.long(), full.idx.long())) and bool(torch.equal(ov, full.val))
_N.check(_lib.leccr_comm_destroy(comm), "leccr_comm_destroy")
ok &= nccl_ok
print(f"rank {rank}: C-ABI NCCL all-gather + merge == single pass {nccl_ok}", flush=True)
# ---- GallerySearchPlan: query shards x gallery parts with the peer merge (P = 2), and pure query sharding with the
# gallery pushed between ranks on the host path (P = 1); both against a single pass over the whole gallery
G5, Q5 = 160_000, 1024
gal5, qry5, _gt5 = synth.cfg5_gallery(G5, Q5, device="cuda", seed=21)
full5, = _ops.sim_topk([(_ops.prep(qry5, want_stats=False), _ops.prep(gal5, want_stats=False), None)], k=10)
gal5_h, qry5_h = gal5.cpu().pin_memory(), qry5.cpu().pin_memory()
plan_ok = True
for parts in ((2, 1) if world % 2 == 0 else (1,)):
    plan = leccr_b200.GallerySearchPlan(G5, Q5, 256, k=10, gallery_parts=parts)
    gb5, ge5 = plan.gallery_rows
    qb5, qe5 = plan.query_rows
    plan.load_device(gal5[gb5:ge5], qry5[qb5:qe5])
    for rep in range(3):  # both slots of the double-buffered lists
        v5, i5, (r0, r1) = plan.search()
        plan_ok &= bool(torch.equal(i5, full5.idx[r0:r1])) and bool(torch.equal(v5, full5.val[r0:r1]))
    for rep in range(3):
        hv5, hi5, (h0, h1) = plan.search_host(gal5_h, qry5_h)
        plan_ok &= (h0, h1) == (r0, r1) and bool(torch.equal(hi5.cuda(), full5.idx[r0:r1])) and bool(torch.equal(hv5.cuda(), full5.val[r0:r1]))
    handles = [plan.search_host_async(gal5_h, qry5_h) for _ in range(2)]   # two searches in flight (both lanes)
    handles.append(plan.search_host_async(gal5_h, qry5_h))                 # a third: waits for the first lane
    for hd in handles[1:]:
        hv5, hi5, _ = hd.result()
        plan_ok &= bool(torch.equal(hi5.cuda(), full5.idx[r0:r1])) and bool(torch.equal(hv5.cuda(), full5.val[r0:r1]))
    print(f"rank {rank}: GallerySearchPlan parts={parts} shards={plan.S} windows={len(plan.bounds)} exchange={plan.xchg is not None} "
          f"h2d {plan.h2d_bytes} B: {'PASS' if plan_ok else 'FAIL'}", flush=True)
ok &= plan_ok
# timing of the training step (fwd + bwd) on this rank, max over ranks
def step():
    me.temp.grad = None
    l = leccr_b200.get_contrastive_loss(me, a, b, idx)
    l.backward()
for _ in range(5): step()
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): step()
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 20 * 1e3], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"contrastive fwd+bwd, B={B}/rank, N={B*world}: {t.item():.1f} us per step (max over ranks)  ALL {'PASS' if ok else 'FAIL'}")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
