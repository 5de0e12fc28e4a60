"""TEST INFRASTRUCTURE -- generates tests/golden/caption_vision_loss.npz by running the REFERENCE's own
RetrievalModel.caption_vision_loss (models/model_retrieval_caption.py:118-143, with the reference's AllGather and
two seeded nn.Linear projections) under gloo with 1 and 2 ranks, with autograd gradients of the local inputs.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden_cv
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


def inputs(N, cn, vn, d, seed):
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(N, vn, d, generator=g)
    caption = image.mean(1)[None] * 0.5 + torch.randn(cn, N, d, generator=g)     # [cn, N, d] as the model holds it
    idx = torch.randint(0, max(2, N // 2), (N,), generator=g)
    Wc, bc = torch.randn(d, d, generator=g) / d ** 0.5, 0.1 * torch.randn(d, generator=g)
    Wv, bv = torch.randn(d, d, generator=g) / d ** 0.5, 0.1 * torch.randn(d, generator=g)
    return image, caption, idx, Wc, bc, Wv, bv


def _worker(rank, world, image, caption, idx, Wc, bc, Wv, bv, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = ref_loader.load()
    cls = importlib.import_module("models.model_retrieval_caption").RetrievalModel
    d = image.shape[2]
    me = types.SimpleNamespace(allgather=ref["AllGather"].apply, cproj=torch.nn.Linear(d, d), vproj=torch.nn.Linear(d, d))
    with torch.no_grad():
        me.cproj.weight.copy_(Wc); me.cproj.bias.copy_(bc); me.vproj.weight.copy_(Wv); me.vproj.bias.copy_(bv)
    B = image.shape[0] // world
    sl = slice(rank * B, (rank + 1) * B)
    im = image[sl].clone().requires_grad_(True)
    cp = caption[:, sl].clone().requires_grad_(True)
    loss = cls.caption_vision_loss(me, cp, im, idx[sl])
    loss.backward()
    q.put((rank, loss.item(), im.grad.numpy(), cp.grad.numpy(), me.cproj.weight.grad.numpy(), me.vproj.weight.grad.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def run(args, world, port):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world) + args + (port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get() for _ in range(world))
    for p in procs:
        p.join()
    return res


def main():
    out = {}
    for name, (N, cn, vn, d, seed, world) in {"w1": (24, 2, 5, 64, 61, 1), "w2": (32, 3, 4, 64, 62, 2)}.items():
        args = inputs(N, cn, vn, d, seed)
        res = run(args, world, 29551 + world)
        for k, v in zip(("image", "caption", "idx", "Wc", "bc", "Wv", "bv"), args):
            out[f"{name}_{k}"] = v.numpy()
        out[f"{name}_world"] = np.int32(world)
        for rank, loss, dim, dcp, dwc, dwv in res:
            out[f"{name}_r{rank}_loss"] = np.float64(loss)
            out[f"{name}_r{rank}_dimage"] = dim
            out[f"{name}_r{rank}_dcaption"] = dcp
            out[f"{name}_r{rank}_dWc"] = dwc
            out[f"{name}_r{rank}_dWv"] = dwv
    np.savez_compressed(os.path.join(OUT, "caption_vision_loss.npz"), **out)
    print("wrote caption_vision_loss.npz", {k: float(v) for k, v in out.items() if k.endswith("loss")})


if __name__ == "__main__":
    main()
