"""TEST INFRASTRUCTURE -- generates tests/golden/dstl_loss.npz by running the REFERENCE's own
RetrievalModel.dstl_loss (models/model_retrieval_caption.py:94-116, with its norm_score :87-90 and the reference's
AllGather) under gloo with 1 and 2 ranks, with autograd gradients of the local inputs.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden_dstl
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


def inputs(n, N, d, seed):
    g = torch.Generator().manual_seed(seed)
    nrm = torch.nn.functional.normalize
    image = nrm(torch.randn(N, d, generator=g), dim=-1)
    text_s = nrm(image + 0.8 * torch.randn(N, d, generator=g), dim=-1)
    text_t = nrm(image + 0.8 * torch.randn(N, d, generator=g), dim=-1)
    cap = text_s[None] + 0.6 * torch.randn(n, N, d, generator=g)   # un-normalised caption queries
    return image, cap, text_s, text_t


def _worker(rank, world, image, cap, text_s, text_t, alpha, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = ref_loader.load()
    mod = importlib.import_module("models.model_retrieval_caption")
    cls = mod.RetrievalModel
    me = types.SimpleNamespace(allgather=ref["AllGather"].apply)
    me.norm_score = lambda score: cls.norm_score(me, score)
    B = image.shape[0] // world
    sl = slice(rank * B, (rank + 1) * B)
    im = image[sl].clone().requires_grad_(True)
    cp = cap[:, sl].clone().requires_grad_(True)
    ts = text_s[sl].clone().requires_grad_(True)
    tt = text_t[sl].clone().requires_grad_(True)
    loss = cls.dstl_loss(me, im, cp, ts, tt, None, alpha=alpha)
    loss.backward()
    z = lambda t, like: np.zeros_like(like.detach().numpy()) if t.grad is None else t.grad.numpy()
    q.put((rank, loss.item(), z(im, im), z(tt, tt), z(ts, ts), z(cp, cp)))
    dist.barrier()
    dist.destroy_process_group()


def run(image, cap, text_s, text_t, alpha, world, port):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, image, cap, text_s, text_t, alpha, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get() for _ in range(world))
    for p in procs:
        p.join()
    return res


def main():
    out = {}
    for name, (n, N, d, seed, world, alpha) in {"w1": (2, 48, 64, 51, 1, 0.8), "w2": (3, 64, 64, 52, 2, 0.8)}.items():
        image, cap, text_s, text_t = inputs(n, N, d, seed)
        res = run(image, cap, text_s, text_t, alpha, world, 29541 + world)
        out.update({f"{name}_image": image.numpy(), f"{name}_caption": cap.numpy(), f"{name}_text_s": text_s.numpy(),
                    f"{name}_text_t": text_t.numpy(), f"{name}_alpha": np.float32(alpha), f"{name}_world": np.int32(world)})
        for rank, loss, dim, dtt, dts, dcp in res:
            out[f"{name}_r{rank}_loss"] = np.float64(loss)
            out[f"{name}_r{rank}_dimage"] = dim
            out[f"{name}_r{rank}_dtext_t"] = dtt
            assert not dts.any() and not dcp.any(), "labels are detached: no gradient to text_s / captions"
    np.savez_compressed(os.path.join(OUT, "dstl_loss.npz"), **out)
    print("wrote dstl_loss.npz", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if "loss" in k})


if __name__ == "__main__":
    main()
