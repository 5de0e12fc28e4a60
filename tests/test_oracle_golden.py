"""The oracle (oracle/oracle.py) against the reference's own outputs (tests/golden, made by
oracle/make_golden.py from /root/reference).  CPU only."""
import numpy as np
import torch

from leccr_b200 import synth
from oracle import oracle

EV_KEYS = ('txt_r1', 'txt_r5', 'txt_r10', 'txt_r_mean', 'txt_sum_r', 'img_r1', 'img_r5', 'img_r10',
           'img_r_mean', 'r_mean', 'img_sumr', 'sumr_avg', 'sumr_sum')


def ev_of(g, prefix):
    return {k: float(g[f"{prefix}{k}"]) for k in EV_KEYS}


def assert_ev_equal(got, want):
    assert set(got) == set(EV_KEYS)
    for k in EV_KEYS:
        assert float(got[k]) == want[k], (k, got[k], want[k])


def test_image_small(golden):
    g = golden("image_small.npz")
    img, txt = torch.from_numpy(g["image"]), torch.from_numpy(g["text"])
    i2t, t2i = oracle.score_matrices(img, txt)
    np.testing.assert_allclose(i2t, g["i2t"], rtol=0, atol=1e-6)
    assert bool(g["t2i_is_view"]) and not t2i.flags["C_CONTIGUOUS"] and np.shares_memory(i2t, t2i)
    n, m = i2t.shape
    txt2img = {t: t // 5 for t in range(m)}
    img2txt = {i: list(range(5 * i, 5 * i + 5)) for i in range(n)}
    want = ev_of(g, "ev_")
    assert_ev_equal(oracle.itm_eval(g["i2t"], g["i2t"].T, txt2img, img2txt), want)
    assert_ev_equal(oracle.itm_eval_by_count(g["i2t"], g["i2t"].T, txt2img, img2txt), want)


def test_video_small(golden):
    g = golden("video_small.npz")
    img, txt, cap = (torch.from_numpy(g[k]) for k in ("image", "text", "caption"))
    i2t, t2i = oracle.double_sim_matrices(img, txt, cap, alpha=0.9, fusion="norm")
    np.testing.assert_allclose(i2t, g["i2t"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(t2i, g["t2i"], rtol=0, atol=1e-6)
    assert i2t.max() <= 0.0 and i2t.min() >= -1.0 - 1e-6
    n = i2t.shape[0]
    want = ev_of(g, "ev_")
    assert_ev_equal(oracle.itm_eval(g["i2t"], g["t2i"], {t: t for t in range(n)}, {i: [i] for i in range(n)}), want)
    gi = golden("image_small.npz")
    np.testing.assert_array_equal(oracle.norm_score(torch.from_numpy(gi["i2t"])).numpy(), g["norm_of_image_i2t"])


def test_contrastive_small(golden):
    g = golden("contrastive_small.npz")
    a, b, idx, temp = torch.from_numpy(g["image"]), torch.from_numpy(g["text"]), torch.from_numpy(g["idx"]), float(g["temp"])
    for name, ix in (("noidx", None), ("idx", idx)):
        loss, da, db, dt = oracle.contrastive_loss_and_grads(a, b, temp, ix)
        assert abs(loss.item() - float(g[f"{name}_loss"])) < 1e-6
        np.testing.assert_allclose(da.numpy(), g[f"{name}_dA"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(db.numpy(), g[f"{name}_dB"], rtol=1e-5, atol=1e-7)
        assert abs(dt.item() - float(g[f"{name}_dtemp"])) <= 1e-5 * abs(float(g[f"{name}_dtemp"]))
    # two ranks: every rank sees the same loss / dtemp and its own slice of the gradients (AllGather.backward)
    bs = a.shape[0] // 2
    for r in range(2):
        loss, da, db, dt = oracle.contrastive_loss_and_grads(a, b, temp, idx, rank=r, batch_size=bs)
        assert abs(loss.item() - float(g[f"w2_r{r}_loss"])) < 1e-6
        np.testing.assert_allclose(da.numpy(), g[f"w2_r{r}_dA"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(db.numpy(), g[f"w2_r{r}_dB"], rtol=1e-5, atol=1e-7)
        assert abs(dt.item() - float(g[f"w2_r{r}_dtemp"])) <= 1e-5 * abs(float(g[f"w2_r{r}_dtemp"]))


def test_cfg1_multi30k(golden):
    g = golden("baseline_configs.npz")
    rs = synth.cfg1_multi30k()
    i2t, t2i = oracle.score_matrices(rs.image, rs.text)
    np.testing.assert_allclose(i2t[g["cfg1_rows"], g["cfg1_cols"]], g["cfg1_vals"], rtol=0, atol=1e-6)
    want = ev_of(g, "cfg1_ev_")
    assert_ev_equal(oracle.itm_eval(i2t, t2i, rs.txt2img, rs.img2txt), want)
    assert_ev_equal(oracle.itm_eval_by_count(i2t, t2i, rs.txt2img, rs.img2txt), want)
    _, order = oracle.topk(i2t[g["cfg1_top10_rows"]], 10)
    np.testing.assert_array_equal(order, g["cfg1_top10"])


def test_cfg2_mscoco5k(golden):
    g = golden("baseline_configs.npz")
    rs = synth.cfg2_mscoco5k()
    i2t, t2i = oracle.score_matrices(rs.image, rs.text)
    np.testing.assert_allclose(i2t[g["cfg2_rows"], g["cfg2_cols"]], g["cfg2_vals"], rtol=0, atol=1e-6)
    assert_ev_equal(oracle.itm_eval_by_count(i2t, t2i, rs.txt2img, rs.img2txt), ev_of(g, "cfg2_ev_"))


def test_cfg4_msrvtt_double_sim(golden):
    g = golden("baseline_configs.npz")
    rs = synth.cfg4_msrvtt()
    i2t, t2i = oracle.double_sim_matrices(rs.image, rs.text, rs.caption, alpha=0.9)
    np.testing.assert_allclose(i2t[g["cfg4_rows"], g["cfg4_cols"]], g["cfg4_vals"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(t2i[g["cfg4_cols"], g["cfg4_rows"]], g["cfg4_vals_t2i"], rtol=0, atol=1e-6)
    assert_ev_equal(oracle.itm_eval(i2t, t2i, rs.txt2img, rs.img2txt), ev_of(g, "cfg4_ev_"))


def test_cfg3_itc(golden):
    g = golden("baseline_configs.npz")
    cb = synth.cfg3_itc()
    for name, ix in (("noidx", None), ("idx", cb.idx)):
        loss, da, db, dt = oracle.contrastive_loss_and_grads(cb.image, cb.text, cb.temp, ix)
        assert abs(loss.item() - float(g[f"cfg3_{name}_loss"])) <= 2e-6 * abs(float(g[f"cfg3_{name}_loss"]))
        assert abs(dt.item() - float(g[f"cfg3_{name}_dtemp"])) <= 1e-4 * abs(float(g[f"cfg3_{name}_dtemp"]))
        np.testing.assert_allclose(da[:8].numpy(), g[f"cfg3_{name}_dA_rows"], rtol=1e-4, atol=1e-8)
        np.testing.assert_allclose(db[:8].numpy(), g[f"cfg3_{name}_dB_rows"], rtol=1e-4, atol=1e-8)
        assert abs(da.norm().item() - float(g[f"cfg3_{name}_dA_norm"])) <= 1e-4 * float(g[f"cfg3_{name}_dA_norm"])


def test_caption_contrastive_loss_matches_reference(golden):
    """oracle.caption_contrastive_loss vs the reference's get_caption_contrastive_loss run by
    oracle/make_golden_caption.py (loss and autograd gradients)."""
    g = golden("caption_loss.npz")
    for c in "abc":
        cap, text, temp = torch.from_numpy(g[f"{c}_caption"]), torch.from_numpy(g[f"{c}_text"]), float(g[f"{c}_temp"])
        loss, dc, dt, dtemp = oracle.caption_contrastive_loss_and_grads(cap, text, temp)
        assert abs(loss.item() - float(g[f"{c}_loss"])) <= 1e-6 * abs(float(g[f"{c}_loss"]))
        assert np.abs(dc.numpy() - g[f"{c}_dcaption"]).max() <= 1e-6 * max(1e-6, np.abs(g[f"{c}_dcaption"]).max())
        assert np.abs(dt.numpy() - g[f"{c}_dtext"]).max() <= 1e-6 * max(1e-6, np.abs(g[f"{c}_dtext"]).max())
        assert abs(dtemp.item() - float(g[f"{c}_dtemp"])) <= 1e-5 * abs(float(g[f"{c}_dtemp"]))


def test_dstl_loss_matches_reference(golden):
    """oracle.dstl_loss vs the reference's dstl_loss run under gloo with 1 and 2 ranks (oracle/make_golden_dstl.py)."""
    g = golden("dstl_loss.npz")
    for name in ("w1", "w2"):
        world = int(g[f"{name}_world"])
        image, cap = torch.from_numpy(g[f"{name}_image"]), torch.from_numpy(g[f"{name}_caption"])
        ts, tt = torch.from_numpy(g[f"{name}_text_s"]), torch.from_numpy(g[f"{name}_text_t"])
        B = image.shape[0] // world
        for rank in range(world):
            loss, dim, dtt = oracle.dstl_loss_and_grads(image, cap, ts, tt, float(g[f"{name}_alpha"]), rank, B)
            want = float(g[f"{name}_r{rank}_loss"])
            assert abs(loss.item() - want) <= 2e-6 * abs(want) + 1e-9
            for got, ref in ((dim, g[f"{name}_r{rank}_dimage"]), (dtt, g[f"{name}_r{rank}_dtext_t"])):
                assert np.abs(got.numpy() - ref).max() <= 1e-5 * np.abs(ref).max() + 1e-10


def test_caption_vision_loss_matches_reference(golden):
    """oracle.caption_vision_loss vs the reference's caption_vision_loss under gloo with 1 and 2 ranks
    (oracle/make_golden_cv.py): loss, local input gradients and the projection-weight gradients of each rank."""
    g = golden("caption_vision_loss.npz")
    for name in ("w1", "w2"):
        world = int(g[f"{name}_world"])
        t = lambda k: torch.from_numpy(g[f"{name}_{k}"])
        B = t("image").shape[0] // world
        for rank in range(world):
            loss, dim, dcp, dwc, dwv = oracle.caption_vision_loss_and_grads(t("caption"), t("image"), t("idx"), t("Wc"),
                                                                            t("bc"), t("Wv"), t("bv"), rank, B)
            want = float(g[f"{name}_r{rank}_loss"])
            assert abs(loss.item() - want) <= 2e-6 * abs(want)
            for got, key in ((dim, "dimage"), (dcp, "dcaption"), (dwc, "dWc"), (dwv, "dWv")):
                ref = g[f"{name}_r{rank}_{key}"]
                assert np.abs(got.numpy() - ref).max() <= 2e-5 * np.abs(ref).max() + 1e-9, key


def test_gt20_small_every_ground_truth_entry_counts(golden):
    """20 ground-truth texts per image (the MSR-VTT layout): rank = min over ALL of them (:274-278)."""
    g = golden("gt20_small.npz")
    n, m = g["i2t"].shape
    assert m == 20 * n
    txt2img = {t: t // 20 for t in range(m)}
    img2txt = {i: list(range(20 * i, 20 * i + 20)) for i in range(n)}
    want = ev_of(g, "ev_")
    assert_ev_equal(oracle.itm_eval(g["i2t"], g["i2t"].T, txt2img, img2txt), want)
    assert_ev_equal(oracle.itm_eval_by_count(g["i2t"], g["i2t"].T, txt2img, img2txt), want)
    # a reader that stopped after 16 entries would differ on this set (the fixture is only useful if so)
    cut = oracle.itm_eval_by_count(g["i2t"], g["i2t"].T, txt2img, {i: v[:16] for i, v in img2txt.items()})
    assert any(float(cut[k]) != want[k] for k in ("txt_r1", "txt_r5", "txt_r10"))


def test_gallery_small(golden):
    """cfg5 in small: the oracle's gallery_eval against the reference's matmul + itm_eval + argsort."""
    g = golden("gallery_small.npz")
    gal = torch.from_numpy(g["gallery_bf16_bits"]).view(torch.bfloat16)
    qry = torch.from_numpy(g["query_bf16_bits"]).view(torch.bfloat16)
    ev, val, idx = oracle.gallery_eval(gal, qry, g["gt"], k=10)
    for c in (1, 5, 10):
        assert ev[f"img_r{c}"] == float(g[f"ev_img_r{c}"])
    np.testing.assert_allclose(val, g["top10_val"], rtol=0, atol=1e-6)
    assert (idx == g["top10"]).mean() > 0.999   # argsort order at exact ties is unspecified


def test_get_features(golden):
    g = golden("get_features.npz")
    for tok, W, b, feat in (("img_tok", "vW", "vb", "feat_i"), ("txt_tok", "tW", "tb", "feat_t")):
        got = oracle.get_features(torch.from_numpy(g[tok]), torch.from_numpy(g[W]), torch.from_numpy(g[b]))
        np.testing.assert_allclose(got.numpy(), g[feat], rtol=0, atol=1e-6)
