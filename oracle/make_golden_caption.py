"""TEST INFRASTRUCTURE -- generates tests/golden/caption_loss.npz by running the REFERENCE's own
RetrievalModel.get_caption_contrastive_loss (models/model_retrieval_caption.py:145-152; the video model's copy,
models/video_model_retrieval_caption.py:171-178, is the same code) on seeded inputs, with autograd gradients.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden_caption
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def case(fn, n, bsz, d, seed, temp=0.07, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    text = torch.nn.functional.normalize(torch.randn(bsz, d, generator=g), dim=-1)
    # caption queries: NOT normalised (output of caption_proj1), correlated with their text
    cap = scale * (text[None] + 0.6 * torch.randn(n, bsz, d, generator=g))
    cap_r = cap.clone().requires_grad_(True)
    text_r = text.clone().requires_grad_(True)
    me = types.SimpleNamespace(temp=torch.nn.Parameter(torch.tensor(temp)))
    loss = fn(me, cap_r, text_r)
    loss.backward()
    return {"caption": cap.numpy(), "text": text.numpy(), "temp": np.float32(temp), "loss": loss.detach().numpy(),
            "dcaption": cap_r.grad.numpy(), "dtext": text_r.grad.numpy(), "dtemp": me.temp.grad.numpy()}


def main():
    ref_loader.load()  # installs the stubs and the reference on sys.path
    mod = importlib.import_module("models.model_retrieval_caption")
    fn = mod.RetrievalModel.get_caption_contrastive_loss
    out = {}
    for name, (n, bsz, d, seed, scale) in {"a": (2, 48, 64, 31, 1.0), "b": (4, 70, 64, 32, 1.0),
                                           "c": (1, 33, 64, 33, 2.5)}.items():
        for k, v in case(fn, n, bsz, d, seed, scale=scale).items():
            out[f"{name}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "caption_loss.npz"), **out)
    print("wrote caption_loss.npz", {k: v.shape for k, v in out.items() if k.endswith("caption")})


if __name__ == "__main__":
    main()
