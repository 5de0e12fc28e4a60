"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leccr_b200
from leccr_b200 import synth
rs = synth.retrieval_set(301, 3, d=64, seed=4, n_caption_queries=2)
ev, topk = leccr_b200.fused_eval(rs.image, rs.text, rs.txt2img, rs.img2txt)
S = leccr_b200.score_matrix(rs.image, rs.text)
ev2 = leccr_b200.itm_eval(S.cpu().numpy(), S.cpu().numpy().T, rs.txt2img, rs.img2txt)
rv = synth.retrieval_set(130, 1, d=64, seed=5, n_caption_queries=2)
D = leccr_b200.double_sim_matrix(rv.image, rv.text, rv.caption, 0.9)
cb = synth.cfg3_itc(333, d=64, seed=3)
me = types.SimpleNamespace(embed_dim=64, temp=torch.nn.Parameter(torch.tensor(0.07, device="cuda")))
a = cb.image.cuda().requires_grad_(True); b = cb.text.cuda().requires_grad_(True)
for idx in (None, cb.idx.cuda()):
    loss = leccr_b200.get_contrastive_loss(me, a, b, idx); loss.backward()
torch.cuda.synchronize()
print("sanitize case done", ev["r_mean"], ev2["r_mean"], float(D.mean()), loss.item())
