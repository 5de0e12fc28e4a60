"""TEST INFRASTRUCTURE -- round-2 additions to tests/golden, produced by the REFERENCE's own functions.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden_r2

  gt20_small.npz     itm_eval (image_Retrieval_caption.py:261-317) on a set with 20 ground-truth texts per image
                     (the MSR-VTT layout of run_video.sh: 20 captions per video) -- pins rank = min over EVERY
                     ground-truth entry (:274-278), not the first 16.
  gallery_small.npz  the cfg5 shape in small: bf16-stored gallery x queries, one ground-truth gallery row per
                     query; scores by the reference's matmul (:151), Recall by its itm_eval, top-10 by argsort.
  get_features.npz   XVLMBase.get_features (models/xvlm.py:241-256): F.normalize(proj(x[:, 0, :])) for both
                     modalities, with autograd gradients w.r.t. the projection output (pins the fused
                     normalise + its backward).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leccr_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import _dict_arrays, ref_image_eval  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    ref = ref_loader.load()
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---- 20 ground-truth texts per image
    rs = synth.retrieval_set(30, 20, d=64, seed=31)
    i2t, t2i = ref_image_eval(ref, rs)
    ev = ref["itm_eval"](i2t, t2i, rs.txt2img, rs.img2txt)
    np.savez_compressed(os.path.join(OUT, "gt20_small.npz"), image=rs.image.numpy(), text=rs.text.numpy(),
                        i2t=i2t, **_dict_arrays("ev_", ev))
    print("gt20", {k: round(float(v), 3) for k, v in ev.items()})

    # ---- cfg5 in small: the gallery plays the reference's "image" set, the queries its "text" set
    gal, qry, gt = synth.cfg5_gallery(3000, 256, d=64, seed=1237)
    rs5 = types.SimpleNamespace(image=gal.float(), text=qry.float())
    g2q, q2g = ref_image_eval(ref, rs5)                      # (3000, 256) and its transpose view (256, 3000)
    txt2img = {q: int(gt[q]) for q in range(qry.shape[0])}
    img2txt = {g: [] for g in range(gal.shape[0])}            # gallery rows carry no ground truth of their own
    ev5 = ref["itm_eval"](g2q, q2g, txt2img, img2txt)
    order = np.argsort(-q2g, axis=1, kind="stable")[:, :10]
    np.savez_compressed(os.path.join(OUT, "gallery_small.npz"), gallery_bf16_bits=gal.view(torch.int16).numpy(),
                        query_bf16_bits=qry.view(torch.int16).numpy(), gt=gt.numpy(), top10=order,
                        top10_val=np.take_along_axis(q2g, order, 1),
                        **_dict_arrays("ev_", {k: v for k, v in ev5.items() if k.startswith("img_")}))
    print("gallery", {k: round(float(v), 3) for k, v in ev5.items() if k.startswith("img_")})

    # ---- get_features (models/xvlm.py:241-256) driven unbound on a SimpleNamespace carrying the two projections
    import importlib

    xvlm = importlib.import_module("models.xvlm")
    g = torch.Generator().manual_seed(41)
    B, T, W, D = 24, 5, 96, 64
    vproj, tproj = torch.nn.Linear(W, D), torch.nn.Linear(W, D)
    with torch.no_grad():
        for lin in (vproj, tproj):
            lin.weight.copy_(torch.randn(D, W, generator=g) / W ** 0.5)
            lin.bias.copy_(0.1 * torch.randn(D, generator=g))
    me = types.SimpleNamespace(vision_proj=vproj, text_proj=tproj)
    img_tok = torch.randn(B, T, W, generator=g, requires_grad=True)
    txt_tok = torch.randn(B, T, W, generator=g, requires_grad=True)
    fi, ft = xvlm.XVLMBase.get_features(me, img_tok, txt_tok)
    up_i, up_t = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
    ((fi * up_i).sum() + (ft * up_t).sum()).backward()
    np.savez_compressed(os.path.join(OUT, "get_features.npz"), img_tok=img_tok.detach().numpy(),
                        txt_tok=txt_tok.detach().numpy(), vW=vproj.weight.detach().numpy(), vb=vproj.bias.detach().numpy(),
                        tW=tproj.weight.detach().numpy(), tb=tproj.bias.detach().numpy(), feat_i=fi.detach().numpy(),
                        feat_t=ft.detach().numpy(), up_i=up_i.numpy(), up_t=up_t.numpy(),
                        d_img_tok=img_tok.grad.numpy(), d_txt_tok=txt_tok.grad.numpy(),
                        d_vW=vproj.weight.grad.numpy(), d_tW=tproj.weight.grad.numpy())
    print("get_features", fi.shape, ft.shape)


if __name__ == "__main__":
    main()
