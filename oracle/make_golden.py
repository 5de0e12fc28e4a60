"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the REFERENCE's own functions.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
The reference has no tests or golden vectors; these fixtures are what pins the oracle (and through it
the CUDA path) to the reference's behaviour.  Small cases store inputs + full outputs; the
BASELINE.json-sized cases (cfg1..cfg4) regenerate their inputs from leccr_b200.synth seeds and store the
reference's Recall dict / loss plus sampled entries, so the fixtures stay small.
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

OUT = os.path.join(ROOT, "tests", "golden")


# ----------------------------------------------------------------------------- fakes that feed embeddings
class _Tok:
    """Stands in for the HF tokenizer: 'texts' are decimal strings, input_ids carry the index."""

    def __call__(self, text, padding=None, truncation=None, max_length=None, return_tensors=None):
        ids = torch.tensor([[int(t)] for t in text], dtype=torch.long)
        ns = types.SimpleNamespace(input_ids=ids, attention_mask=torch.ones_like(ids))
        ns.to = lambda device: ns
        return ns


class _ImageModel:
    def __init__(self, image, text):
        self.image, self.text = image, text

    def eval(self):
        pass

    def get_text_embeds(self, ids, mask):
        return ids

    def get_vision_embeds(self, image):
        return image, None

    def get_caption_embeds(self, ids, mask):
        return ids

    def interaction_with_caption(self, image_embeds, caption_embeds, key_padding_mask):
        return image_embeds.view(1, -1, 1), None, None  # the caller transposes (0, 1)

    def get_features(self, image_embeds=None, text_embeds=None):
        if text_embeds is not None:
            return self.text[text_embeds[:, 0]]
        return self.image[image_embeds[:, 0, 0].long()]


class _VideoModel(_ImageModel):
    def __init__(self, image, text, caption):
        super().__init__(image, text)
        self.caption = caption  # [n, N, D]

    def get_vision_embeds(self, video, mask):
        return video, mask

    def interaction_with_caption(self, image_embeds, caption_embeds, key_padding_mask, video_mask=None):
        return image_embeds.view(1, -1, 1), image_embeds, None

    def get_features(self, image_embeds=None, text_embeds=None, vis_mask=None):
        return super().get_features(image_embeds=image_embeds, text_embeds=text_embeds)

    def caption_proj1(self, idx):
        return self.caption[:, idx.view(-1).long()]


def _loader(n_items, bs, video=False):
    ds = types.SimpleNamespace(text=None)

    class L:
        dataset = ds

        def __iter__(self):
            for s in range(0, n_items, bs):
                ids = torch.arange(s, min(n_items, s + bs), dtype=torch.float32)
                caps = [str(int(i)) for i in ids]
                if video:
                    yield ids, torch.ones(len(ids), 1), caps, ids
                else:
                    yield ids, caps, ids

    return L()


CONFIG = {"batch_size_test_text": 64, "max_tokens": 8, "caption_encoder_name": "mbert"}


def ref_image_eval(ref, rs):
    mod = ref["image_module"]
    mod.args = types.SimpleNamespace(distributed=False)
    loader = _loader(rs.image.shape[0], 32)
    loader.dataset.text = [str(t) for t in range(rs.text.shape[0])]
    return ref["image_evaluation_coarse"](_ImageModel(rs.image, rs.text), loader, _Tok(), "cpu", CONFIG)


def ref_video_eval(ref, rs, alpha=0.9):
    mod = ref["video_module"]
    mod.args = types.SimpleNamespace(distributed=False)
    loader = _loader(rs.image.shape[0], 32, video=True)
    loader.dataset.text = [str(t) for t in range(rs.text.shape[0])]
    return ref["video_evaluation_coarse"](_VideoModel(rs.image, rs.text, rs.caption), loader, _Tok(), "cpu",
                                          CONFIG, alpha=alpha)


def _dict_arrays(prefix, d):
    return {f"{prefix}{k}": np.float64(v) for k, v in d.items()}


# ----------------------------------------------------------------------------- reference contrastive loss
def _loss_worker(rank, world, a, b, idx, temp, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = ref_loader.load()
    bs = a.shape[0] // world
    al = a[rank * bs:(rank + 1) * bs].clone().requires_grad_(True)
    bl = b[rank * bs:(rank + 1) * bs].clone().requires_grad_(True)
    t = torch.tensor(temp, requires_grad=True)
    me = types.SimpleNamespace(embed_dim=a.shape[1], temp=t)
    il = None if idx is None else idx[rank * bs:(rank + 1) * bs]
    loss = ref["get_contrastive_loss"](me, al, bl, il)
    loss.backward()
    q.put((rank, loss.item(), al.grad.numpy(), bl.grad.numpy(), t.grad.item()))
    dist.barrier()
    dist.destroy_process_group()


def ref_loss(a, b, idx, temp, world=1, port=29533):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_loss_worker, args=(r, world, a, b, idx, temp, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get() for _ in range(world))
    for p in procs:
        p.join()
    return res


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load()
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---- small image retrieval case: full inputs and outputs
    rs = synth.retrieval_set(40, 5, d=64, seed=11)
    i2t, t2i = ref_image_eval(ref, rs)
    ev = ref["itm_eval"](i2t, t2i, rs.txt2img, rs.img2txt)
    np.savez_compressed(os.path.join(OUT, "image_small.npz"), image=rs.image.numpy(), text=rs.text.numpy(),
                        i2t=i2t, t2i_is_view=np.bool_(not t2i.flags["C_CONTIGUOUS"]), **_dict_arrays("ev_", ev))

    # ---- small video double_sim case
    rv = synth.retrieval_set(48, 1, d=64, seed=12, n_caption_queries=2)
    vi2t, vt2i = ref_video_eval(ref, rv, alpha=0.9)
    vev = ref["video_itm_eval"](vi2t, vt2i, rv.txt2img, rv.img2txt)
    ns = ref["norm_score"](torch.from_numpy(i2t)).numpy()
    np.savez_compressed(os.path.join(OUT, "video_small.npz"), image=rv.image.numpy(), text=rv.text.numpy(),
                        caption=rv.caption.numpy(), i2t=vi2t, t2i=np.ascontiguousarray(vt2i), norm_of_image_i2t=ns,
                        **_dict_arrays("ev_", vev))

    # ---- small contrastive cases: world 1 (idx None / idx) and world 2 (AllGather slices)
    cb = synth.cfg3_itc(96, d=64, seed=13)
    # copies: torch.multiprocessing moves the tensors' storage to shared memory, which would leave
    # numpy views of the old storage dangling
    out = {"image": cb.image.numpy().copy(), "text": cb.text.numpy().copy(), "idx": cb.idx.numpy().copy(),
           "temp": np.float32(cb.temp)}
    for name, idx in (("noidx", None), ("idx", cb.idx)):
        (_, loss, ga, gb, gt), = ref_loss(cb.image, cb.text, idx, cb.temp, world=1)
        out.update({f"{name}_loss": np.float64(loss), f"{name}_dA": ga, f"{name}_dB": gb, f"{name}_dtemp": np.float64(gt)})
    res = ref_loss(cb.image, cb.text, cb.idx, cb.temp, world=2, port=29534)
    for r, loss, ga, gb, gt in res:
        out.update({f"w2_r{r}_loss": np.float64(loss), f"w2_r{r}_dA": ga, f"w2_r{r}_dB": gb,
                    f"w2_r{r}_dtemp": np.float64(gt)})
    np.savez_compressed(os.path.join(OUT, "contrastive_small.npz"), **out)

    # ---- BASELINE-sized cases: inputs come from synth seeds, fixtures hold the reference's answers
    rng = np.random.default_rng(0)
    big = {}
    for tag, rs_big in (("cfg1", synth.cfg1_multi30k()), ("cfg2", synth.cfg2_mscoco5k())):
        bi2t, bt2i = ref_image_eval(ref, rs_big)
        bev = ref["itm_eval"](bi2t, bt2i, rs_big.txt2img, rs_big.img2txt)
        rows = rng.integers(0, bi2t.shape[0], 512)
        cols = rng.integers(0, bi2t.shape[1], 512)
        big.update(_dict_arrays(f"{tag}_ev_", bev))
        big.update({f"{tag}_rows": rows, f"{tag}_cols": cols, f"{tag}_vals": bi2t[rows, cols],
                    f"{tag}_top10_rows": rows[:64], f"{tag}_top10": np.argsort(-bi2t[rows[:64]], axis=1)[:, :10]})
        print(tag, {k: round(float(v), 3) for k, v in bev.items()})
    r4 = synth.cfg4_msrvtt()
    v4i, v4t = ref_video_eval(ref, r4, alpha=0.9)
    ev4 = ref["video_itm_eval"](v4i, v4t, r4.txt2img, r4.img2txt)
    rows = rng.integers(0, 1000, 512)
    cols = rng.integers(0, 1000, 512)
    big.update(_dict_arrays("cfg4_ev_", ev4))
    big.update({"cfg4_rows": rows, "cfg4_cols": cols, "cfg4_vals": v4i[rows, cols], "cfg4_vals_t2i": v4t[cols, rows]})
    print("cfg4", {k: round(float(v), 3) for k, v in ev4.items()})
    c3 = synth.cfg3_itc()
    for name, idx in (("noidx", None), ("idx", c3.idx)):
        (_, loss, ga, gb, gt), = ref_loss(c3.image, c3.text, idx, c3.temp, world=1, port=29535)
        big.update({f"cfg3_{name}_loss": np.float64(loss), f"cfg3_{name}_dtemp": np.float64(gt),
                    f"cfg3_{name}_dA_rows": ga[:8], f"cfg3_{name}_dB_rows": gb[:8],
                    f"cfg3_{name}_dA_norm": np.float64(np.linalg.norm(ga)),
                    f"cfg3_{name}_dB_norm": np.float64(np.linalg.norm(gb))})
        print("cfg3", name, loss, gt)
    np.savez_compressed(os.path.join(OUT, "baseline_configs.npz"), **big)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
