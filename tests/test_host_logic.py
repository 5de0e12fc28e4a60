"""CPU-side checks: the C-ABI library loads and exports what include/leccr_b200.h declares, the host
logic around it, and the N>1 plumbing over gloo (world_size 2).  No compute calls without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

from leccr_b200 import _native as N
from leccr_b200 import evaluation, ops, synth
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "leccr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(leccr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = N.load()
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(N.EXPORTS) == names, "ctypes signatures and header disagree"
    assert lib.leccr_abi_version() == 2
    assert b"sm_100" in lib.leccr_strerror(-3)


def test_topk_problem_struct_matches_header():
    # 21 fields; pointer / int64 / int layout must be what the C struct has (x86-64: 8-byte slots)
    assert [f[0] for f in N.TopkProblem._fields_][:6] == ["rows16", "cols16", "ld_rows16", "ld_cols16", "n_rows", "n_cols"]
    import ctypes

    assert ctypes.sizeof(N.TopkProblem) == 21 * 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    rs = synth.retrieval_set(8, 2, d=64, seed=1)
    with pytest.raises(N.LeccrError):
        evaluation.fused_eval(rs.image, rs.text, rs.txt2img, rs.img2txt)
    with pytest.raises(N.LeccrError):
        evaluation.itm_eval(np.zeros((2, 2), np.float32), np.zeros((2, 2), np.float32), {0: 0, 1: 1}, {0: [0], 1: [1]})
    with pytest.raises(N.LeccrError):
        ops.prep(rs.image)
    lib = N.load()
    assert lib.leccr_check_device() != 0  # no device: refuses


def test_metrics_from_counts_matches_reference_arithmetic():
    want = oracle.metrics_from_recalls(100.0 * 552 / 1000, 100.0 * 841 / 1000, 100.0 * 921 / 1000,
                                       100.0 * 1419 / 5000, 100.0 * 2517 / 5000, 100.0 * 3074 / 5000)
    got = evaluation.metrics_from_counts([552, 841, 921], 1000, [1419, 2517, 3074], 5000)
    assert got == want and tuple(got) == evaluation.EVAL_KEYS


def test_transpose_view_detection():
    a = np.zeros((3, 5), np.float32)
    assert evaluation._is_transpose_view(a, a.T)
    assert not evaluation._is_transpose_view(a, np.ascontiguousarray(a.T))
    assert not evaluation._is_transpose_view(a, np.zeros((5, 3), np.float32))


def test_csr_from_lists():
    off, ids = ops.csr_from_lists([[3, 4], [], [7]], "cpu")
    assert off.tolist() == [0, 2, 2, 3] and ids.tolist() == [3, 4, 7] and off.dtype == torch.int32


def test_synth_is_deterministic():
    a, b = synth.cfg1_multi30k(), synth.cfg1_multi30k()
    assert torch.equal(a.image, b.image) and torch.equal(a.text, b.text)
    assert a.img2txt[7] == [35, 36, 37, 38, 39] and a.txt2img[36] == 7
    assert torch.allclose(a.image.norm(dim=1), torch.ones(1000), atol=1e-5)
    c = synth.cfg4_msrvtt()
    assert c.caption.shape == (2, 1000, 256)


def _allgather_worker(rank, world, port, q):
    import torch.distributed as dist

    from leccr_b200.allgather import allgather

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    full = torch.randn(world * 3, 4, generator=g)
    x = full[rank * 3:(rank + 1) * 3].clone().requires_grad_(True)
    out = allgather(x, rank, world)
    w = torch.arange(out.numel(), dtype=torch.float32).view_as(out)
    (out * w).sum().backward()
    idx = torch.arange(rank * 3, (rank + 1) * 3).view(-1, 1)
    idx_all = allgather(idx, rank, world)
    q.put((rank, torch.equal(out.detach(), full), torch.equal(x.grad, w[rank * 3:(rank + 1) * 3]),
           idx_all.view(-1).tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_two_ranks_gloo():
    """models/xvlm.py:50-67 semantics: rank-ordered concat forward, own slice backward, no reduction."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_allgather_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for rank, fwd_ok, bwd_ok, idx_all in res:
        assert fwd_ok and bwd_ok and idx_all == list(range(6))


def test_query_sharding_covers_everything():
    from leccr_b200 import sharding

    for n, w in ((100000, 8), (5000, 3), (7, 4), (1, 2)):
        spans = [sharding.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_merge_topk_lists_matches_global_topk():
    from leccr_b200 import sharding

    g = torch.Generator().manual_seed(3)
    scores = torch.randn(17, 400, generator=g)
    k = 10
    parts_v, parts_i = [], []
    for (b, e) in (sharding.shard_range(400, r, 4) for r in range(4)):
        v, i = torch.topk(scores[:, b:e], k, dim=1)
        parts_v.append(v)
        parts_i.append(i + b)
    mv, mi = sharding.merge_topk(torch.stack(parts_v), torch.stack(parts_i), k)
    wv, wi = torch.topk(scores, k, dim=1)
    assert torch.equal(mv, wv) and torch.equal(mi, wi)


def test_install_rebinds_the_reference_names():
    """install() / install_caption_loss() / install_scripts() replace exactly the names the
    reference resolves at call time.  Needs the reference tree (build container only)."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference tree not present on this machine")
    import importlib
    import types

    import leccr_b200
    import leccr_b200.install as inst

    ref_loader.load()  # stubs for the reference's missing third-party modules + sys.path
    done = inst.install()
    assert "models.xvlm" in done
    xvlm = importlib.import_module("models.xvlm")
    assert xvlm.AllGather is leccr_b200.AllGather and xvlm.allgather is leccr_b200.allgather
    assert xvlm.XVLMBase.get_contrastive_loss is leccr_b200.get_contrastive_loss
    done = inst.install_caption_loss()
    assert "models.model_retrieval_caption" in done
    mrc = importlib.import_module("models.model_retrieval_caption")
    assert mrc.RetrievalModel.get_caption_contrastive_loss is leccr_b200.get_caption_contrastive_loss
    assert mrc.RetrievalModel.dstl_loss is leccr_b200.dstl_loss
    script = types.ModuleType("fake_task_script")
    inst.install_scripts(image_module=script)
    assert script.itm_eval is leccr_b200.itm_eval and callable(script.evaluation_coarse)
    # the signatures the reference calls with still bind (models/model_retrieval_caption.py:187-193,
    # image_Retrieval_caption.py:453-459)
    import inspect

    assert list(inspect.signature(leccr_b200.get_contrastive_loss).parameters) == ["self", "image_feat", "text_feat", "idx"]
    assert list(inspect.signature(leccr_b200.get_caption_contrastive_loss).parameters) == ["self", "caption_embeds", "text_feats"]
    assert list(inspect.signature(leccr_b200.dstl_loss).parameters) == ["self", "image_embeds", "caption_embeds", "text_embeds_s", "text_embeds_t", "idx", "alpha"]
    assert list(inspect.signature(leccr_b200.itm_eval).parameters) == ["scores_i2t", "scores_t2i", "txt2img", "img2txt"]
    assert list(inspect.signature(script.evaluation_coarse).parameters) == ["model", "data_loader", "tokenizer", "device", "config"]


def _packed_gather_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from leccr_b200.dstl_loss import gather_packed

    g = torch.Generator().manual_seed(3)
    n, B, D = 3, 5, 8
    image = torch.randn(world * B, D, generator=g)
    ts = torch.randn(world * B, D, generator=g)
    tt = torch.randn(world * B, D, generator=g)
    cap = torch.randn(n, world * B, D, generator=g)
    sl = slice(rank * B, (rank + 1) * B)
    loc = [image[sl].clone().requires_grad_(True), cap[:, sl].clone().requires_grad_(True),
           ts[sl].clone().requires_grad_(True), tt[sl].clone().requires_grad_(True)]
    im_all, cap_all, ts_all, tt_all = gather_packed(loc[0], loc[1], loc[2], loc[3], rank, world)
    fwd_ok = bool(torch.equal(im_all, image) and torch.equal(cap_all, cap) and torch.equal(ts_all, ts)
                  and torch.equal(tt_all, tt))
    # backward: every gathered element weighted by a distinct number; each rank must get its own slice back
    w = [torch.arange(t.numel(), dtype=torch.float32).view_as(t) + 1000 * k for k, t in enumerate((image, cap, ts, tt))]
    (im_all * w[0]).sum().add((cap_all * w[1]).sum()).add((ts_all * w[2]).sum()).add((tt_all * w[3]).sum()).backward()
    bwd_ok = bool(torch.equal(loc[0].grad, w[0][sl]) and torch.equal(loc[1].grad, w[1][:, sl])
                  and torch.equal(loc[2].grad, w[2][sl]) and torch.equal(loc[3].grad, w[3][sl]))
    q.put((rank, fwd_ok, bwd_ok))
    dist.barrier()
    dist.destroy_process_group()


def test_dstl_packed_gather_two_ranks_gloo():
    """dstl_loss issues the reference's four all-gathers (models/model_retrieval_caption.py:95-98) as one collective:
    same gathered tensors in rank order, and AllGather's backward (own slice, no reduction) on each of them."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_packed_gather_worker, args=(r, 2, 29613, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for rank, fwd_ok, bwd_ok in res:
        assert fwd_ok and bwd_ok, (rank, fwd_ok, bwd_ok)


def _project_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from leccr_b200.contrastive import _project

    g = torch.Generator().manual_seed(4)
    lin = torch.nn.Linear(6, 6)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(6, 6, generator=g))
        lin.bias.copy_(torch.randn(6, generator=g))
    x_all = torch.randn(world * 4, 3, 6, generator=g)
    x = x_all[rank * 4:(rank + 1) * 4].clone().requires_grad_(True)
    _project(lin, x, world).pow(2).sum().backward()
    # what the reference's module holds after projecting ALL gathered rows: the full parameter gradient
    ref = torch.nn.Linear(6, 6)
    ref.load_state_dict(lin.state_dict())
    xr = x_all.clone().requires_grad_(True)
    ref(xr).pow(2).sum().backward()
    ok = bool(torch.allclose(lin.weight.grad, ref.weight.grad, rtol=1e-5, atol=1e-6)
              and torch.allclose(lin.bias.grad, ref.bias.grad, rtol=1e-5, atol=1e-6)
              and torch.allclose(x.grad, xr.grad[rank * 4:(rank + 1) * 4], rtol=1e-5, atol=1e-6))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_caption_vision_projection_gradients_two_ranks_gloo():
    """caption_vision_loss projects before the gather; the parameter gradients of cproj / vproj must still be the
    reference's (projection after the gather: every rank's module sees all rows), the input gradients local."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_project_worker, args=(r, 2, 29615, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_gallery_plan_layout_helpers():
    """Pure host logic of GallerySearchPlan: window bounds tile the gallery part in whole 256-row column tiles and
    grow geometrically; the (windows, slots) choice respects the 8 list slots per query row; shard ranges tile."""
    from leccr_b200.gallery import _pick_windows, _window_bounds
    from leccr_b200.sharding import shard_range

    for part, w in ((1_000_000, 4), (500_000, 2), (500_000, 3), (70_000, 2), (300, 2), (257, 8), (1, 1)):
        b = _window_bounds(part, w)
        assert b[0][0] == 0 and b[-1][1] == part and len(b) <= w
        assert all(e0 == b1 for (_, e0), (b1, _) in zip(b[:-1], b[1:]))          # contiguous, no overlap
        assert all(e > s for s, e in b) and all(s % 256 == 0 for s, _ in b)      # whole column tiles
        if len(b) > 2:
            sizes = [e - s for s, e in b]
            assert sizes == sorted(sizes)                                         # only the first upload is exposed
    for rb, part in ((782, 1_000_000), (391, 500_000), (196, 500_000), (3, 70_000), (1, 1000)):
        subs = _pick_windows(rb, part)
        assert 1 <= len(subs) <= 4 and all(s >= 1 for s in subs) and sum(subs) <= 8
        assert len(subs) == 1 or part // ((1 << len(subs)) - 1) >= 32768
    assert _pick_windows(196, 500_000) == [2, 3, 3] and _pick_windows(782, 1_000_000) == [2, 2, 2, 2]
    for n, world in ((100_000, 8), (1_000_003, 8), (5, 8), (7, 2)):
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


def test_new_entries_fail_loudly_without_a_gpu():
    import leccr_b200
    from leccr_b200 import _native as N

    with pytest.raises(N.LeccrError):
        leccr_b200.GallerySearchPlan(1000, 10, 64)
    with pytest.raises(N.LeccrError):
        leccr_b200.normalize_rows(torch.randn(4, 8))
    with pytest.raises(N.LeccrError):
        leccr_b200.FeatureGallery(64)


def test_property_gallery_layout_and_merge_random():
    """Randomised (hypothesis) versions of the host-logic checks: window bounds and slot counts for any part size,
    shard ranges for any (n, world), and the k-way merge of partial top-k lists against a global top-k."""
    from hypothesis import given, settings, strategies as st

    from leccr_b200.gallery import _pick_windows, _window_bounds
    from leccr_b200.sharding import merge_topk, shard_range

    @settings(max_examples=200, deadline=None, derandomize=True)
    @given(part=st.integers(1, 3_000_000), w=st.integers(1, 8))
    def bounds(part, w):
        b = _window_bounds(part, w)
        assert b[0][0] == 0 and b[-1][1] == part and 1 <= len(b) <= w
        assert all(e0 == b1 for (_, e0), (b1, _) in zip(b[:-1], b[1:]))
        assert all(e > s for s, e in b) and all(s % 256 == 0 for s, _ in b)

    @settings(max_examples=200, deadline=None, derandomize=True)
    @given(rb=st.integers(1, 4000), part=st.integers(1, 3_000_000))
    def slots(rb, part):
        subs = _pick_windows(rb, part)
        assert 1 <= len(subs) <= 4 and all(s >= 1 for s in subs) and sum(subs) <= 8
        assert len(subs) == 1 or part // ((1 << len(subs)) - 1) >= 32768

    @settings(max_examples=200, deadline=None, derandomize=True)
    @given(n=st.integers(0, 2_000_000), world=st.integers(1, 16))
    def shards(n, world):
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(q=st.integers(1, 40), parts=st.integers(1, 8), per=st.integers(1, 60), k=st.integers(1, 16), seed=st.integers(0, 10**6))
    def merge(q, parts, per, k, seed):
        g = torch.Generator().manual_seed(seed)
        S = torch.randn(q, parts * per, generator=g)
        vals, idxs = [], []
        for p in range(parts):
            kk = min(k, per)
            v, i = S[:, p * per:(p + 1) * per].topk(kk, dim=1)
            if kk < k:   # short lists are padded the way the kernels pad them
                v = torch.cat([v, torch.full((q, k - kk), float("-inf"))], 1)
                i = torch.cat([i, torch.full((q, k - kk), -1, dtype=i.dtype)], 1)
            vals.append(v)
            idxs.append(torch.where(i >= 0, i + p * per, i))
        mv, mi = merge_topk(torch.stack(vals), torch.stack(idxs), k)
        kk = min(k, parts * per)
        wv, wi = S.topk(kk, dim=1)
        assert torch.equal(mv[:, :kk], wv) and torch.equal(mi[:, :kk].long(), wi)

    bounds()
    slots()
    shards()
    merge()


def test_build_staleness_follows_source_contents(tmp_path, monkeypatch):
    """The library is rebuilt when the CONTENT of any source differs from what it was compiled from (mtimes do not
    survive the snapshot copy to the GPU box), and only then."""
    from leccr_b200 import build as B

    srcs = []
    for i in range(3):
        p = tmp_path / f"s{i}.cu"
        p.write_text(f"// source {i}\n")
        srcs.append(str(p))
    lib = tmp_path / "lib.so"
    monkeypatch.setattr(B, "DEPS", srcs)
    monkeypatch.setattr(B, "LIB", str(lib))
    monkeypatch.setattr(B, "HASH", str(lib) + ".srchash")
    assert B.is_stale()                                   # nothing built yet
    lib.write_bytes(b"\x7fELF")
    assert B.is_stale()                                   # a library without a recorded hash is not trusted
    (tmp_path / "lib.so.srchash").write_text(B.source_hash() + "\n")
    assert not B.is_stale()
    os.utime(srcs[1], (1, 1))                             # touching a file changes nothing
    assert not B.is_stale()
    with open(srcs[2], "a") as f:
        f.write("// edited\n")
    assert B.is_stale()


def test_numa_binding_context_leaves_the_affinity_as_it_found_it():
    """Without NVML / a GPU the context manager must be a no-op (bound False) and the thread's CPU affinity must be
    what it was, also when the body raises."""
    from leccr_b200.peer import host_memory_near_device

    before = os.sched_getaffinity(0)
    with host_memory_near_device(0) as near:
        assert near.bound in (False, True)
    assert os.sched_getaffinity(0) == before
    with pytest.raises(RuntimeError):
        with host_memory_near_device(3):
            raise RuntimeError("body failed")
    assert os.sched_getaffinity(0) == before
