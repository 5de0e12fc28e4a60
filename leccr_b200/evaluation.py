"""Drop-ins for the reference's retrieval evaluation (the hot path's eval half).

  itm_eval(scores_i2t, scores_t2i, txt2img, img2txt)            image_Retrieval_caption.py:261-317
                                                                 == video_Retrieval_caption_double_sim.py:194-247
  evaluation_coarse(model, data_loader, tokenizer, device, config)          image_Retrieval_caption.py:83-163
  evaluation_coarse_video(model, data_loader, tokenizer, device, config, alpha=0.9)
                                                                 video_Retrieval_caption_double_sim.py:94-190
plus the additive
  fused_eval(image_embeds, text_embeds, txt2img, img2txt, ...)   similarity + top-k + Recall@1/5/10 in one
                                                                 tensor-core pass; N x M never reaches HBM.

Ranking rule: rank(row) = min over the row's ground truth g of #{j : s_j > s_g}.  This equals the
reference's np.argsort(score)[::-1] position whenever the row has no exactly tied scores (the reference
leaves the order of ties to numpy's unstable sort).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _native as N
from . import ops

EVAL_KEYS = ('txt_r1', 'txt_r5', 'txt_r10', 'txt_r_mean', 'txt_sum_r', 'img_r1', 'img_r5', 'img_r10',
             'img_r_mean', 'r_mean', 'img_sumr', 'sumr_avg', 'sumr_sum')


def _device():
    if not torch.cuda.is_available():
        raise N.LeccrError("leccr_b200 has no CPU path: a CUDA device (B200) is required")
    return torch.device("cuda", torch.cuda.current_device())


def _to_device(x, dev, dtype=None):
    """numpy / CPU tensor / CUDA tensor -> CUDA tensor (async H2D from pinned staging for host data)."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not x.is_cuda:
        x = x.contiguous()
        if not x.is_pinned():
            try:
                x = x.pin_memory()
            except RuntimeError:
                pass
        x = x.to(dev, non_blocking=True)
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    return x


def metrics_from_counts(c_i2t, n_i2t, c_t2i, n_t2i) -> dict:
    """The 13-key dict of image_Retrieval_caption.py:281-316 from #{rank < 1, 5, 10} per direction."""
    tr1, tr5, tr10 = (100.0 * int(c) / n_i2t for c in c_i2t)
    ir1, ir5, ir10 = (100.0 * int(c) / n_t2i for c in c_t2i)
    tr_mean = (tr1 + tr5 + tr10) / 3
    ir_mean = (ir1 + ir5 + ir10) / 3
    r_mean = (tr_mean + ir_mean) / 2
    txt_sumr = tr1 + tr5 + tr10
    img_sumr = ir1 + ir5 + ir10
    sumr_avg = np.round((txt_sumr + img_sumr) / 6, 2)
    return {'txt_r1': tr1, 'txt_r5': tr5, 'txt_r10': tr10, 'txt_r_mean': tr_mean, 'txt_sum_r': txt_sumr,
            'img_r1': ir1, 'img_r5': ir5, 'img_r10': ir10, 'img_r_mean': ir_mean, 'r_mean': r_mean,
            'img_sumr': img_sumr, 'sumr_avg': sumr_avg, 'sumr_sum': (txt_sumr + img_sumr)}


def _gt_lists(txt2img, img2txt, n_img, n_txt):
    return [list(img2txt[i]) for i in range(n_img)], [[txt2img[t]] for t in range(n_txt)]


def _is_transpose_view(a: np.ndarray, b: np.ndarray) -> bool:
    return (isinstance(a, np.ndarray) and isinstance(b, np.ndarray) and a.ndim == 2 and b.shape == a.shape[::-1]
            and b.strides == a.strides[::-1] and np.shares_memory(a, b)
            and b.__array_interface__['data'][0] == a.__array_interface__['data'][0])


# evaluation_coarse returns numpy matrices because the reference's main() expects them (:163), and itm_eval then
# receives the same arrays back: the device copy is kept (a few entries, dropped when the array dies) so the
# N x M matrix crosses PCIe once instead of twice.  The arrays are handed out read-only; a caller that makes one
# writeable again gets the plain host path.
_device_copies = {}


def _remember_device_copy(arr: np.ndarray, dev_tensor: torch.Tensor):
    import weakref

    arr.flags.writeable = False
    key = id(arr)
    _device_copies[key] = (weakref.ref(arr, lambda _r, k=key: _device_copies.pop(k, None)), dev_tensor)
    while len(_device_copies) > 8:
        _device_copies.pop(next(iter(_device_copies)))


def _device_copy_of(arr):
    if not isinstance(arr, np.ndarray):
        return None
    hit = _device_copies.get(id(arr))
    if hit is None or hit[0]() is not arr or arr.flags.writeable or hit[1].shape != arr.shape:
        return None
    return hit[1]


@torch.no_grad()
def itm_eval(scores_i2t, scores_t2i, txt2img, img2txt):
    """Drop-in for itm_eval: ranks every row on the GPU (one pass over the matrix per direction)."""
    dev = _device()
    n_img, n_txt = scores_i2t.shape
    gi, gt = _gt_lists(txt2img, img2txt, n_img, n_txt)
    S = _device_copy_of(scores_i2t)
    if S is None:
        S = _to_device(scores_i2t, dev, torch.float32)
    r_i = ops.rank_rows(S, *ops.csr_from_lists(gi, dev))
    if _is_transpose_view(scores_i2t, scores_t2i) or scores_t2i is None:
        r_t = ops.rank_cols(S, *ops.csr_from_lists(gt, dev))  # the reference's t2i IS i2t.T (:152)
    else:
        St = _to_device(scores_t2i, dev, torch.float32)
        r_t = ops.rank_rows(St, *ops.csr_from_lists(gt, dev))
    c_i = ops.recall_counts(r_i).cpu().tolist()
    c_t = ops.recall_counts(r_t).cpu().tolist()
    return metrics_from_counts(c_i, n_img, c_t, n_txt)


def _precise_operands(rows: torch.Tensor, cols: torch.Tensor, precision: str):
    fmt = ops.fmt_of(precision)
    if precision.endswith("x3"):
        return ops.prep(rows, fmt, N.LAYOUT_X3_ROWS, want_stats=False), ops.prep(cols, fmt, N.LAYOUT_X3_COLS,
                                                                                want_stats=False)
    return ops.prep(rows, fmt, want_stats=False), ops.prep(cols, fmt, want_stats=False)


@torch.no_grad()
def score_matrix(image_embeds, text_embeds, scale: float = 1.0, precision: str = "f16x3") -> torch.Tensor:
    """score_matrix_i2t = image_embeds @ text_embeds.t()  (image_Retrieval_caption.py:151) on the tensor
    cores; the default split-precision operands make the product fp32-accurate."""
    dev = _device()
    a, b = _precise_operands(_to_device(image_embeds, dev, torch.float32), _to_device(text_embeds, dev, torch.float32),
                             precision)
    return ops.sim_matrix(a, b, scale)


@torch.no_grad()
def double_sim_matrix(image_embeds, text_embeds, caption_embeds, alpha: float, fusion: str = "norm",
                      scale: float = 1.0, precision: str = "f16x3") -> torch.Tensor:
    """alpha * f(S) + (1 - alpha) * f(max_n C_n)   (video_...double_sim.py:170-178 / image_...py:235-246)."""
    dev = _device()
    img = _to_device(image_embeds, dev, torch.float32)
    txt = _to_device(text_embeds, dev, torch.float32)
    cap = _to_device(caption_embeds, dev, torch.float32)
    n_cap, n_vid, d = cap.shape
    fmt = ops.fmt_of(precision)
    x3 = precision.endswith("x3")
    t_op = ops.prep(txt, fmt, N.LAYOUT_X3_COLS if x3 else N.LAYOUT_HI, want_stats=False)
    i_op = ops.prep(img, fmt, N.LAYOUT_X3_ROWS if x3 else N.LAYOUT_HI, want_stats=False)
    c_op = ops.prep(cap.reshape(n_cap * n_vid, d), fmt, N.LAYOUT_X3_ROWS if x3 else N.LAYOUT_HI, want_stats=False)
    S = ops.sim_matrix(i_op, t_op)
    Cn = ops.sim_matrix(c_op, t_op).view(n_cap, n_vid, txt.shape[0])
    ops.double_sim_fuse(S, Cn, alpha, N.FUSE_NORM if fusion == "norm" else N.FUSE_RAW)
    if scale != 1.0:
        S.mul_(scale)
    return S


def _dist_scale(distributed: bool) -> float:
    """The reference all-reduces (SUM) score matrices that every rank already holds in full
    (image_Retrieval_caption.py:154-157), i.e. multiplies them by world_size; rankings are unaffected."""
    if distributed and dist.is_available() and dist.is_initialized():
        return float(dist.get_world_size())
    return 1.0


class FeatureGallery:
    """An embedding set that arrives batch by batch, written straight into a preallocated tensor-core operand.

    The reference appends every batch's features to a Python list and `torch.cat`s them at the end
    (image_Retrieval_caption.py:112-118,144-148; video_Retrieval_caption_double_sim.py:124-131,158-162): two
    copies of the set and one allocation per batch.  Here the cast prologue (leccr_prep) writes each batch at its
    row offset of ONE 16-bit buffer in the layout the similarity pass consumes (split-precision [hi|lo|hi] /
    [hi|hi|lo] for the fp32-faithful matrices the drop-ins return), so nothing is kept in fp32 and nothing is
    concatenated.  `capacity` is the number of rows when known (len(dataset)); otherwise the buffer doubles."""

    def __init__(self, dim, precision="f16x3", role="rows", capacity=0, device=None):
        self.dev = device or _device()
        self.D = dim
        self.fmt = ops.fmt_of(precision)
        self.layout = N.LAYOUT_HI if not precision.endswith("x3") else (N.LAYOUT_X3_ROWS if role == "rows" else N.LAYOUT_X3_COLS)
        self.K = dim if self.layout == N.LAYOUT_HI else 3 * dim
        self.dt16 = torch.float16 if self.fmt == N.FMT_F16 else torch.bfloat16
        self.n = 0
        self.buf = torch.empty((max(int(capacity), 0), self.K), dtype=self.dt16, device=self.dev)

    def append(self, feats: torch.Tensor):
        """feats: [b, D] fp32 CUDA rows (e.g. model.get_features(...) of one batch)."""
        if feats.dim() != 2 or feats.shape[1] != self.D:
            raise N.LeccrError(f"FeatureGallery expects [b, {self.D}] rows")
        if not feats.is_cuda:
            raise N.LeccrError("leccr_b200 has no CPU path: features must be CUDA tensors")
        b = feats.shape[0]
        if b == 0:
            return
        if feats.dtype != torch.float32 or feats.stride(1) != 1:
            feats = feats.float().contiguous()
        if self.n + b > self.buf.shape[0]:
            grown = torch.empty((max(2 * self.buf.shape[0], self.n + b, 1024), self.K), dtype=self.dt16, device=self.dev)
            grown[:self.n].copy_(self.buf[:self.n])
            self.buf = grown
        dst = self.buf[self.n:self.n + b]
        N.check(N.load().leccr_prep(N.ptr(feats), b, self.D, feats.stride(0), 0, self.fmt, self.layout, N.ptr(dst), self.K,
                                    None, None, None, N.stream_ptr()), "leccr_prep")
        self.n += b

    def operand(self) -> ops.Operand:
        t16 = self.buf[:self.n]
        return ops.Operand(t16, self.fmt, self.layout, self.n, self.D, None, None, None, t16)


def _len_or_zero(x):
    try:
        return len(x)
    except TypeError:
        return 0


def _collect_text(model, texts, tokenizer, device, config, precision="f16x3"):
    """Text side of evaluation_coarse (image_Retrieval_caption.py:99-118): encoder + get_features per batch as in
    the reference, each batch written into the preallocated column operand."""
    bs = config['batch_size_test_text']
    gal = None
    for i in range(0, len(texts), bs):
        chunk = texts[i: min(len(texts), i + bs)]
        tok = tokenizer(chunk, padding='max_length', truncation=True, max_length=config['max_tokens'],
                        return_tensors="pt").to(device)
        feat = model.get_text_embeds(tok.input_ids, tok.attention_mask)
        f = model.get_features(text_embeds=feat)
        if gal is None:
            gal = FeatureGallery(f.shape[1], precision, "cols", len(texts), f.device)
        gal.append(f)
    return gal


def _caption_inputs(model, generated_captions, tokenizer, device, config, clip_tokenizer):
    if config['caption_encoder_name'] == 'clip':
        if clip_tokenizer is None:
            raise N.LeccrError("caption_encoder_name == 'clip' needs the reference's clip.tokenize (pass clip_tokenizer)")
        captions = clip_tokenizer(generated_captions).to(device)
        return model.get_caption_embeds(captions), torch.zeros_like(captions).masked_fill_(captions == 0, 1).bool()
    tok = tokenizer(generated_captions, padding='max_length', truncation=True, max_length=config['max_tokens'],
                    return_tensors="pt").to(device)
    return model.get_caption_embeds(tok.input_ids, tok.attention_mask), ~tok.attention_mask.bool()


@torch.no_grad()
def evaluation_coarse(model, data_loader, tokenizer, device, config, distributed=False, clip_tokenizer=None):
    """Drop-in for the image evaluation_coarse: encoders run as in the reference (they are out of scope),
    the similarity stage (:147-163) runs on the tensor cores.  Returns (i2t, t2i) numpy, t2i a view of i2t.T."""
    model.eval()
    texts = _collect_text(model, data_loader.dataset.text, tokenizer, device, config)
    images = None
    for image, generated_captions, img_id in data_loader:
        image = image.to(device)
        image_feat, _ = model.get_vision_embeds(image)
        caption_embed, kpm = _caption_inputs(model, generated_captions, tokenizer, device, config, clip_tokenizer)
        image_feat, _, _ = model.interaction_with_caption(image_embeds=image_feat, caption_embeds=caption_embed,
                                                          key_padding_mask=kpm)
        image_feat = image_feat.transpose(0, 1).contiguous()
        f = model.get_features(image_embeds=image_feat)
        if images is None:
            images = FeatureGallery(f.shape[1], "f16x3", "rows", _len_or_zero(data_loader.dataset), f.device)
        images.append(f)
    S = ops.sim_matrix(images.operand(), texts.operand(), _dist_scale(distributed))
    i2t = S.cpu().numpy()
    _remember_device_copy(i2t, S)
    return i2t, i2t.T


@torch.no_grad()
def evaluation_coarse_video(model, data_loader, tokenizer, device, config, alpha=0.9, distributed=False,
                            clip_tokenizer=None):
    """Drop-in for the video evaluation_coarse with the double_sim fusion (:164-190)."""
    model.eval()
    texts = _collect_text(model, data_loader.dataset.text, tokenizer, device, config)
    videos, captions = None, None
    for video, mask_video, generated_captions, img_id in data_loader:
        video = video.to(device)
        mask_video = mask_video.to(device)
        image_feat, image_atts = model.get_vision_embeds(video, mask_video)
        caption_embed, kpm = _caption_inputs(model, generated_captions, tokenizer, device, config, clip_tokenizer)
        image_feat, caption_embed, _ = model.interaction_with_caption(
            image_embeds=image_feat, caption_embeds=caption_embed, key_padding_mask=kpm, video_mask=image_atts)
        image_feat = image_feat.transpose(0, 1).contiguous()
        f = model.get_features(image_embeds=image_feat, vis_mask=mask_video.unsqueeze(-1))
        cap = model.caption_proj1(caption_embed)          # [n, bsz, d], not normalised (:157)
        if videos is None:
            total = _len_or_zero(data_loader.dataset)
            videos = FeatureGallery(f.shape[1], "f16x3", "rows", total, f.device)
            captions = [FeatureGallery(cap.shape[2], "f16x3", "rows", total, f.device) for _ in range(cap.shape[0])]
        videos.append(f)
        for q, g in enumerate(captions):
            g.append(cap[q])
    t_op = texts.operand()
    S = ops.sim_matrix(videos.operand(), t_op)
    Cn = torch.empty((len(captions),) + tuple(S.shape), dtype=torch.float32, device=S.device)
    for q, g in enumerate(captions):
        ops.sim_matrix(g.operand(), t_op, out=Cn[q])
    ops.double_sim_fuse(S, Cn, alpha, N.FUSE_NORM)
    scale = _dist_scale(distributed)
    if scale != 1.0:
        S.mul_(scale)
    i2t = S.cpu().numpy()
    _remember_device_copy(i2t, S)
    return i2t, i2t.T


@torch.no_grad()
def prepare_gt(txt2img, img2txt, n_img, n_txt, device=None):
    """Ground-truth maps (the dataset's txt2img / img2txt dicts) -> device CSR pair, reusable across calls."""
    dev = device or _device()
    gi, gt = _gt_lists(txt2img, img2txt, n_img, n_txt)
    return ops.csr_from_lists(gi, dev), ops.csr_from_lists(gt, dev)


@torch.no_grad()
def _fused_eval_double_sim_materialized(image_embeds, text_embeds, caption_embeds, txt2img, img2txt, k, alpha, fusion,
                                        return_topk, gt):
    """double_sim with the fused matrix alpha * f(S) + (1 - alpha) * f(max_n C_n) MATERIALISED on the device,
    ranked there (leccr_rank_rows / leccr_rank_cols) and its top-k lists read by leccr_topk_dense; only the six
    counts cross PCIe.  The fallback of _fused_eval_double_sim for ground-truth maps the epilogue path does not
    take (texts with several ground-truth videos, maps that are not inverses, more than 7 caption queries)."""
    dev = _device()
    F_ = double_sim_matrix(image_embeds, text_embeds, caption_embeds, alpha, fusion)
    n_img, n_txt = F_.shape
    if gt is None:
        gt = prepare_gt(txt2img, img2txt, n_img, n_txt, dev)
    r_i = ops.rank_rows(F_, *gt[0])
    r_t = ops.rank_cols(F_, *gt[1])
    host = torch.cat([ops.recall_counts(r_i), ops.recall_counts(r_t)]).cpu().tolist()
    ev = metrics_from_counts(host[0:3], n_img, host[3:6], n_txt)
    if not return_topk:
        return ev
    return ev, {'i2t': ops.topk_dense(F_, k), 't2i': ops.topk_dense(F_, k, by_columns=True)}


def _single_gt_per_text(gt, n_img, n_txt):
    """txt_gt [n_txt] int32 when every text has exactly one ground-truth video AND the video CSR is the inverse
    map (what the reference's datasets build, dataset/retrieval_dataset_video.py:201-219); else None.
    The verdict costs two host syncs, so it is remembered ON the ground-truth tensors themselves (an attribute of
    v_off naming the three other tensors by weak reference): a cache keyed by device addresses would answer for a
    different ground truth once the allocator hands the addresses out again."""
    import weakref

    (v_off, v_ids), (t_off, t_ids) = gt
    memo = getattr(v_off, "_leccr_inverse", None)
    if memo is not None:
        r_ids, r_toff, r_tids, m_img, m_txt, ok = memo
        if r_ids() is v_ids and r_toff() is t_off and r_tids() is t_ids and (m_img, m_txt) == (n_img, n_txt):
            return t_ids if ok else None
    ok = False
    if t_ids.numel() == n_txt and v_ids.numel() == n_txt and t_off.numel() == n_txt + 1 and v_off.numel() == n_img + 1:
        counts = (v_off[1:] - v_off[:-1]).long()
        owner = torch.repeat_interleave(torch.arange(n_img, device=v_off.device), counts)
        ok = bool(torch.equal(t_off.long(), torch.arange(n_txt + 1, device=t_off.device))) and \
            bool(torch.equal(t_ids.long()[v_ids.long()], owner))
    v_off._leccr_inverse = (weakref.ref(v_ids), weakref.ref(t_off), weakref.ref(t_ids), n_img, n_txt, ok)
    return t_ids if ok else None


@torch.no_grad()
def _fused_eval_double_sim(image_embeds, text_embeds, caption_embeds, txt2img, img2txt, k, alpha, fusion, return_topk, gt,
                           precision="f16x3"):
    """double_sim evaluation with the fusion IN the tensor-core epilogue (leccr_double_sim_topk): the videos and
    their caption queries are interleaved into one operand, pass 1 reduces the global min / max of S and
    max_n C_n and picks up the ground-truth scores, pass 2 recomputes, fuses with the reference's fp32 operation
    order (video_Retrieval_caption_double_sim.py:87-91,175-179), counts ranks and keeps the top-k lists.
    No N x M buffer exists; six counts cross PCIe."""
    dev = _device()
    img = _to_device(image_embeds, dev, torch.float32)
    txt = _to_device(text_embeds, dev, torch.float32)
    cap = _to_device(caption_embeds, dev, torch.float32)
    n_cap, n_img, d = cap.shape
    n_txt = txt.shape[0]
    if gt is None:
        gt = prepare_gt(txt2img, img2txt, n_img, n_txt, dev)
    txt_gt = _single_gt_per_text(gt, n_img, n_txt)
    if txt_gt is None or n_cap > 7 or d % 8 != 0:
        return _fused_eval_double_sim_materialized(img, txt, cap, txt2img, img2txt, k, alpha, fusion, return_topk, gt)
    lib = N.load()
    fmt = ops.fmt_of(precision)
    x3 = precision.endswith("x3")
    G = 2 if n_cap == 1 else (4 if n_cap <= 3 else 8)
    K = 3 * d if x3 else d
    dt16 = torch.float16 if fmt == N.FMT_F16 else torch.bfloat16
    vc = torch.empty((G * n_img, K), dtype=dt16, device=dev)
    lay_vc = N.LAYOUT_X3_COLS if x3 else N.LAYOUT_HI
    if not img.is_contiguous():
        img = img.contiguous()
    if not cap.is_contiguous():
        cap = cap.contiguous()
    st = N.stream_ptr()
    for m in range(G):  # row G i + m of the interleaved operand: the video, its captions, the last caption again
        src = img if m == 0 else cap[min(m, n_cap) - 1]
        N.check(lib.leccr_prep(N.ptr(src), n_img, d, src.stride(0), 0, fmt, lay_vc, vc.data_ptr() + m * K * 2, G * K,
                               None, None, None, st), "leccr_prep")
    t16 = ops.prep(txt, fmt, N.LAYOUT_X3_ROWS if x3 else N.LAYOUT_HI, want_stats=False).t16
    f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)
    rank_v, rank_t = torch.empty(n_img, **i32), torch.empty(n_txt, **i32)
    counts = torch.empty(6, **i32)
    tv_vc = ti_vc = tv_t = ti_t = None
    if return_topk:
        tv_vc, ti_vc = torch.empty((G * n_img, k), **f32), torch.empty((G * n_img, k), **i32)
        tv_t, ti_t = torch.empty((n_txt, k), **f32), torch.empty((n_txt, k), **i32)
    ws = torch.empty(lib.leccr_double_sim_topk_workspace(n_img, n_txt, G), dtype=torch.uint8, device=dev)
    (v_off, v_ids) = gt[0]
    N.check(lib.leccr_double_sim_topk(N.ptr(vc), N.ptr(t16), n_img, n_txt, K, fmt, G, n_cap, float(alpha),
                                      float(1.0 - alpha), N.FUSE_NORM if fusion == "norm" else N.FUSE_RAW,
                                      N.ptr(txt_gt), N.ptr(v_off), N.ptr(v_ids), k, N.ptr(tv_vc), N.ptr(ti_vc), N.ptr(tv_t),
                                      N.ptr(ti_t), N.ptr(rank_v), N.ptr(rank_t), N.ptr(counts), N.ptr(ws), ws.numel(), st),
            "leccr_double_sim_topk")
    host = counts.cpu().tolist()
    ev = metrics_from_counts(host[0:3], n_img, host[3:6], n_txt)
    if not return_topk:
        return ev
    return ev, {'i2t': (tv_vc[::G], ti_vc[::G]), 't2i': (tv_t, ti_t), 'rank_i2t': rank_v, 'rank_t2i': rank_t}


@torch.no_grad()
def fused_eval(image_embeds, text_embeds, txt2img=None, img2txt=None, k=10, precision="f16", tiles_per_chunk=0,
               return_topk=True, gt=None, caption_embeds=None, alpha=0.9, fusion="none"):
    """Similarity + per-row top-k + exact Recall@1/5/10 for both directions in one tensor-core launch.

    image_embeds [N, D], text_embeds [M, D]: numpy / CPU / CUDA, fp32 (cast to 16-bit operands here) or
    fp16 / bf16 (used as they are).  gt: optional prepare_gt(...) result (else built from the dicts).
    caption_embeds [n, N, D] with fusion "norm" (video_Retrieval_caption_double_sim.py:175-179, alpha 0.9) or
    "raw" (image_Retrieval_caption.py:239-246, alpha 0.8) evaluates the double_sim fusion instead.
    Returns (eval dict with the reference's 13 keys, topk) where topk is
    {'i2t': (val [N, k], idx [N, k]), 't2i': (val [M, k], idx [M, k])} CUDA tensors (approximate scores,
    ties inside the 16-bit rounding may be ordered differently from fp32).
    """
    if fusion not in ("none", "raw", "norm"):
        raise ValueError("fusion must be 'none', 'raw' or 'norm'")
    if caption_embeds is not None and fusion != "none":
        return _fused_eval_double_sim(image_embeds, text_embeds, caption_embeds, txt2img, img2txt, k, alpha, fusion,
                                      return_topk, gt)
    dev = _device()
    img = _to_device(image_embeds, dev)
    txt = _to_device(text_embeds, dev)
    fmt = ops.fmt_of(precision)
    I, T = ops.prep(img, fmt), ops.prep(txt, fmt)
    if gt is None:
        gt = prepare_gt(txt2img, img2txt, img.shape[0], txt.shape[0], dev)
    if return_topk:
        r_i, r_t = ops.sim_topk([(I, T, gt[0]), (T, I, gt[1])], k=k, tiles_per_chunk=tiles_per_chunk)
    else:  # Recall only (all itm_eval returns): counting epilogue, no candidate lists (leccr_sim_rank)
        r_i, r_t = ops.sim_rank([(I, T, gt[0]), (T, I, gt[1])])
    # the step's single D2H read: 6 Recall counts + the two operand range flags (32 bytes)
    host = torch.cat([r_i.recall_counts.float(), r_t.recall_counts.float(), I.stats[3:4], T.stats[3:4]]).cpu().tolist()
    if host[6] != 0.0 or host[7] != 0.0:
        raise N.LeccrError("embeddings overflow the fp16 operand format; call fused_eval(precision='bf16')")
    counts = [[int(c) for c in host[0:3]], [int(c) for c in host[3:6]]]
    ev = metrics_from_counts(counts[0], img.shape[0], counts[1], txt.shape[0])
    if not return_topk:
        return ev
    return ev, {'i2t': (r_i.val, r_i.idx), 't2i': (r_t.val, r_t.idx)}


class FusedEvalPlan:
    """Repeated fused evaluations of one shape (serving / per-epoch validation): static device buffers and
    the whole step (cast -> tensor-core pass -> finalize -> Recall counts) captured once in a CUDA graph, so
    a run is one H2D (if the inputs are on the host), one graph launch and one 96-byte D2H.  The graph holds
    only the library's own launches plus ONE zero-fill: no allocation, no packing kernels.

        plan = FusedEvalPlan(n_img, n_txt, dim, txt2img, img2txt)
        ev = plan.run(image_embeds, text_embeds)          # host (ideally pinned) or device fp32 tensors
        ev, topk = plan.run(..., return_topk=True)        # topk tensors are overwritten by the next run
    """

    def __init__(self, n_img, n_txt, dim, txt2img=None, img2txt=None, k=10, precision="f16", gt=None,
                 tiles_per_chunk=0, lists=True):
        """lists=False: Recall only -- the counting epilogue (leccr_sim_rank) instead of the list epilogue; `run`
        then has no top-k to return.  That is all the reference's evaluation computes (itm_eval)."""
        self.dev = dev = _device()
        self.lib = lib = N.load()
        self.lists = bool(lists)
        self.n_img, self.n_txt, self.dim, self.k = n_img, n_txt, dim, k
        self.fmt = ops.fmt_of(precision)
        self.tpc = tiles_per_chunk
        if dim % 8 != 0:
            raise N.LeccrError("embedding dimension must be a multiple of 8 (TMA 16-byte rows)")
        self.gt = gt if gt is not None else prepare_gt(txt2img, img2txt, n_img, n_txt, dev)
        dt16 = torch.float16 if self.fmt == N.FMT_F16 else torch.bfloat16
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.img = torch.zeros((n_img, dim), **f32)
        self.txt = torch.zeros((n_txt, dim), **f32)
        self.img[:, 0] = 1.0  # harmless unit rows for the capture run
        self.txt[:, 0] = 1.0
        self.img16 = torch.empty((n_img, dim), dtype=dt16, device=dev)
        self.txt16 = torch.empty((n_txt, dim), dtype=dt16, device=dev)
        self.rn = torch.empty((4, max(n_img, n_txt)), **f32)
        # one block of state, zeroed by ONE fill per step: words 0-2 / 4-6 Recall counts (int32) of the two
        # directions, words 8-11 / 12-15 operand statistics (fp32) of images / texts
        self.state = torch.zeros(24, **i32)
        self.state_f = self.state.view(torch.float32)
        self.host = torch.empty(24, dtype=torch.int32).pin_memory()
        self.topk = {'i2t': (torch.empty((n_img, k), **f32), torch.empty((n_img, k), **i32)),
                     't2i': (torch.empty((n_txt, k), **f32), torch.empty((n_txt, k), **i32))}
        self.ranks = (torch.empty(n_img, **i32), torch.empty(n_txt, **i32))
        self.gts = (torch.empty(max(1, self.gt[0][1].numel()), **f32), torch.empty(max(1, self.gt[1][1].numel()), **f32))
        self.probs = (N.TopkProblem * 2)()
        sp = self.state.data_ptr()
        for p, (rows, cols, rows16, cols16, nr, nc, key, d, rn0, st_cols) in enumerate((
                (self.img, self.txt, self.img16, self.txt16, n_img, n_txt, 'i2t', 0, 0, sp + 48),
                (self.txt, self.img, self.txt16, self.img16, n_txt, n_img, 't2i', 1, 2, sp + 32))):
            q = self.probs[p]
            q.rows16, q.cols16 = rows16.data_ptr(), cols16.data_ptr()
            q.ld_rows16 = q.ld_cols16 = dim
            q.n_rows, q.n_cols = nr, nc
            q.topk_val, q.topk_idx = self.topk[key][0].data_ptr(), self.topk[key][1].data_ptr()
            q.gt_off, q.gt_ids = self.gt[d][0].data_ptr(), self.gt[d][1].data_ptr()
            q.rows_x, q.cols_x = rows.data_ptr(), cols.data_ptr()
            q.ld_rows_x = q.ld_cols_x = dim
            q.x_dtype = N.F32
            q.rn_hi, q.rn_lo = self.rn[rn0].data_ptr(), self.rn[rn0 + 1].data_ptr()
            q.col_stats = st_cols
            q.rank = self.ranks[d].data_ptr()
            q.recall_counts = sp + 16 * d
            q.gt_score = self.gts[d].data_ptr()
        ws_bytes = lib.leccr_sim_topk_workspace(self.probs, 2, self.tpc) if self.lists else lib.leccr_sim_rank_workspace(self.probs, 2)
        self.ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        self._step()  # warm-up outside capture: lazy module load, kernel attributes
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step()

    def _step(self):
        st = torch.cuda.current_stream().cuda_stream
        lib, d, sp = self.lib, self.dim, self.state.data_ptr()
        self.state.zero_()
        N.check(lib.leccr_prep_pair(self.img.data_ptr(), self.n_img, d, self.img16.data_ptr(), d, self.rn[0].data_ptr(),
                                    self.rn[1].data_ptr(), sp + 32, self.txt.data_ptr(), self.n_txt, d,
                                    self.txt16.data_ptr(), d, self.rn[2].data_ptr(), self.rn[3].data_ptr(), sp + 48, d, 0,
                                    self.fmt, N.LAYOUT_HI, st), "leccr_prep_pair")
        if self.lists:
            N.check(lib.leccr_sim_topk(self.probs, 2, d, self.fmt, self.k, self.tpc, self.ws.data_ptr(), self.ws.numel(), st),
                    "leccr_sim_topk")
        else:
            N.check(lib.leccr_sim_rank(self.probs, 2, d, self.fmt, self.ws.data_ptr(), self.ws.numel(), st), "leccr_sim_rank")

    def launch(self, image_embeds=None, text_embeds=None):
        """Asynchronous part of a run: stage the inputs (if given) and replay the graph."""
        if image_embeds is not None:
            self.img.copy_(torch.as_tensor(image_embeds), non_blocking=True)
        if text_embeds is not None:
            self.txt.copy_(torch.as_tensor(text_embeds), non_blocking=True)
        self.graph.replay()

    @torch.no_grad()
    def run(self, image_embeds=None, text_embeds=None, return_topk=False):
        self.launch(image_embeds, text_embeds)
        self.host.copy_(self.state, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        h = self.host.tolist()
        hf = self.host.view(torch.float32)
        if float(hf[11]) != 0.0 or float(hf[15]) != 0.0:
            raise N.LeccrError("embeddings overflow the fp16 operand format; build the plan with precision='bf16'")
        ev = metrics_from_counts(h[0:3], self.n_img, h[4:7], self.n_txt)
        if return_topk and not self.lists:
            raise N.LeccrError("this plan was built with lists=False: it computes the Recall dict only")
        return (ev, self.topk) if return_topk else ev


class StreamedEvalPlan:
    """Fused evaluation of HOST-resident embeddings with the PCIe transfer overlapped: the text set crosses
    in `windows` pieces on a copy stream while the tensor cores already rank what has arrived
    (leccr_sim_topk_stream).  Per window one launch carries two problems: image rows x the window's text
    columns (candidates accumulate in the image problem's list slots, finalised after the last window) and
    the window's text rows x all images (complete rows, finalised at once).  Same results as fused_eval.

        plan = StreamedEvalPlan(n_img, n_txt, dim, txt2img, img2txt)
        ev = plan.run(image_embeds_host, text_embeds_host)      # pinned fp32 host tensors (or numpy)
    """

    def __init__(self, n_img, n_txt, dim, txt2img=None, img2txt=None, k=10, precision="f16", gt=None,
                 windows=3, img_subs=2, txt_subs=3):
        self.dev = _device()
        lib = N.load()
        self.lib = lib
        self.n_img, self.n_txt, self.dim, self.k = n_img, n_txt, dim, k
        self.fmt = ops.fmt_of(precision)
        if dim % 8 != 0:
            raise N.LeccrError("embedding dimension must be a multiple of 8 (TMA 16-byte rows)")
        if isinstance(windows, int):
            fracs = [1.0 / max(1, windows)] * max(1, windows)
        else:
            fracs = [float(f) / sum(windows) for f in windows]  # shrinking windows shorten the un-overlapped tail
        # column chunks of a streamed launch must stay short (<= 32 tiles of 256 columns, the dense two-warpgroup
        # epilogue): raise the sub-list counts when a window or the image set is longer, within 8 slots per row
        max_win = max(fracs) * n_txt
        img_subs = max(img_subs, -(-int(max_win + 255) // 256 // 32))
        txt_subs = max(txt_subs, -(-((n_img + 255) // 256) // 32))
        if img_subs * len(fracs) > 8:
            fracs = [1.0 / max(1, 8 // img_subs)] * max(1, 8 // img_subs)
            img_subs = max(img_subs, -(-int(max(fracs) * n_txt + 255) // 256 // 32))
        # Larger sets (a window or the image set beyond 8 slots x 32 tiles = 65,536 columns): every call of the
        # plan switches to the long-chunk list shape (LECCR_TOPK_LONG: the filter epilogue of the large-gallery
        # search, which has no chunk-length limit); the slot counts then only balance the load.
        self.long = img_subs * len(fracs) > 8 or txt_subs > 8
        if self.long:
            img_subs = max(1, min(img_subs, 8 // len(fracs)))
            txt_subs = min(txt_subs, 8)
        long_flag = N.TOPK_LONG if self.long else 0
        self.gt = gt if gt is not None else prepare_gt(txt2img, img2txt, n_img, n_txt, self.dev)
        dev = self.dev
        dt16 = torch.float16 if self.fmt == N.FMT_F16 else torch.bfloat16
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.img = torch.empty((n_img, dim), **f32)
        self.txt = torch.empty((n_txt, dim), **f32)
        self.img16 = torch.empty((n_img, dim), dtype=dt16, device=dev)
        self.txt16 = torch.empty((n_txt, dim), dtype=dt16, device=dev)
        self.rn = torch.empty((4, max(n_img, n_txt)), **f32)       # rn_hi / rn_lo of images, of texts
        self.small = torch.zeros(16, **f32)                          # [0:4] image stats, [4:8] text stats
        self.counts = torch.zeros(8, **i32)                          # [0:3] i2t, [4:7] t2i Recall counts
        self.out = torch.empty(8, **f32)
        self.host = torch.empty(8, dtype=torch.float32).pin_memory()
        self.topk = {'i2t': (torch.empty((n_img, k), **f32), torch.empty((n_img, k), **i32)),
                     't2i': (torch.empty((n_txt, k), **f32), torch.empty((n_txt, k), **i32))}
        self.ranks = (torch.empty(n_img, **i32), torch.empty(n_txt, **i32))
        self.gts = (torch.empty(max(1, self.gt[0][1].numel()), **f32), torch.empty(max(1, self.gt[1][1].numel()), **f32))
        # window boundaries: equal pieces, multiples of 256 columns except the last
        self.bounds = []
        b, acc = 0, 0.0
        for i, f in enumerate(fracs):
            acc += f
            e = n_txt if i == len(fracs) - 1 else min(n_txt, (int(round(acc * n_txt)) + 255) // 256 * 256)
            if e > b:
                self.bounds.append((b, e))
            b = max(b, e)
        if b < n_txt:
            self.bounds[-1] = (self.bounds[-1][0], n_txt)
        per = max(e - b for b, e in self.bounds)
        W = len(self.bounds)
        self.sub_total = W * img_subs
        self.ws_img = torch.empty(lib.leccr_sim_topk_stream_workspace(n_img, self.sub_total), dtype=torch.uint8, device=dev)
        self.ws_txt = torch.empty(lib.leccr_sim_topk_stream_workspace(per, txt_subs), dtype=torch.uint8, device=dev)
        esz = 2
        gi_off, gi_ids = self.gt[0]
        gt_off, gt_ids = self.gt[1]
        # ground-truth CSR of a text window: offsets rebased on the host once (one GT image per text in the
        # reference's datasets, but any CSR works)
        off_host = gt_off.cpu()
        self._keep = []
        self.calls = []
        for w, (b, e) in enumerate(self.bounds):
            pr = (N.TopkProblem * 2)()
            so = (N.TopkStream * 2)()
            # problem 0: image rows x this window's text columns (tensor-core phase only)
            p = pr[0]
            p.rows16, p.cols16 = self.img16.data_ptr(), self.txt16.data_ptr() + b * dim * esz
            p.ld_rows16 = p.ld_cols16 = dim
            p.n_rows, p.n_cols = n_img, e - b
            o = so[0]
            o.phases = long_flag | N.TOPK_GEMM | (N.TOPK_INIT if w == 0 else 0)
            o.sub_begin, o.sub_count, o.sub_total = w * img_subs, img_subs, self.sub_total
            o.col_begin = b
            o.workspace, o.workspace_bytes = self.ws_img.data_ptr(), self.ws_img.numel()
            # problem 1: this window's text rows x all images, complete
            woff = (off_host[b:e + 1] - off_host[b]).to(torch.int32).to(dev)
            wids = gt_ids[int(off_host[b]):int(off_host[e])]
            self._keep += [woff, wids]
            p = pr[1]
            p.rows16, p.cols16 = self.txt16.data_ptr() + b * dim * esz, self.img16.data_ptr()
            p.ld_rows16 = p.ld_cols16 = dim
            p.n_rows, p.n_cols = e - b, n_img
            p.topk_val = self.topk['t2i'][0].data_ptr() + b * k * 4
            p.topk_idx = self.topk['t2i'][1].data_ptr() + b * k * 4
            p.gt_off, p.gt_ids = woff.data_ptr(), wids.data_ptr()
            p.rows_x, p.cols_x = self.txt.data_ptr() + b * dim * 4, self.img.data_ptr()
            p.ld_rows_x = p.ld_cols_x = dim
            p.x_dtype = N.F32
            p.rn_hi, p.rn_lo = self.rn[2].data_ptr() + b * 4, self.rn[3].data_ptr() + b * 4
            p.col_stats = self.small.data_ptr()
            p.rank = self.ranks[1].data_ptr() + b * 4
            p.recall_counts = self.counts.data_ptr() + 16
            p.gt_score = self.gts[1].data_ptr() + int(off_host[b]) * 4
            o = so[1]
            o.phases = long_flag | N.TOPK_INIT | N.TOPK_GEMM | N.TOPK_FINALIZE
            o.sub_begin, o.sub_count, o.sub_total = 0, txt_subs, txt_subs
            o.workspace, o.workspace_bytes = self.ws_txt.data_ptr(), self.ws_txt.numel()
            self.calls.append((pr, so, 2))
        # final call: merge the image problem's slots
        pr = (N.TopkProblem * 1)()
        so = (N.TopkStream * 1)()
        p = pr[0]
        p.rows16, p.cols16 = self.img16.data_ptr(), self.txt16.data_ptr()
        p.ld_rows16 = p.ld_cols16 = dim
        p.n_rows, p.n_cols = n_img, n_txt
        p.topk_val, p.topk_idx = self.topk['i2t'][0].data_ptr(), self.topk['i2t'][1].data_ptr()
        p.gt_off, p.gt_ids = gi_off.data_ptr(), gi_ids.data_ptr()
        p.rows_x, p.cols_x = self.img.data_ptr(), self.txt.data_ptr()
        p.ld_rows_x = p.ld_cols_x = dim
        p.x_dtype = N.F32
        p.rn_hi, p.rn_lo = self.rn[0].data_ptr(), self.rn[1].data_ptr()
        p.col_stats = self.small.data_ptr() + 16
        p.rank = self.ranks[0].data_ptr()
        p.recall_counts = self.counts.data_ptr()
        p.gt_score = self.gts[0].data_ptr()
        o = so[0]
        o.phases = long_flag | N.TOPK_FINALIZE
        o.sub_begin, o.sub_count, o.sub_total = 0, self.sub_total, self.sub_total
        o.n_cols_total = n_txt
        o.workspace, o.workspace_bytes = self.ws_img.data_ptr(), self.ws_img.numel()
        self.calls.append((pr, so, 1))
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ev_img = torch.cuda.Event()
        self.ev_win = [torch.cuda.Event() for _ in self.bounds]

    def _prep(self, src, dst16, row0, row1, rn_hi, rn_lo, stats_ptr, st):
        d = self.dim
        N.check(self.lib.leccr_prep(src.data_ptr() + row0 * d * 4, row1 - row0, d, d, 0, self.fmt, N.LAYOUT_HI,
                                    dst16.data_ptr() + row0 * d * 2, d, rn_hi.data_ptr() + row0 * 4,
                                    rn_lo.data_ptr() + row0 * 4, stats_ptr, st), "leccr_prep")

    def _issue(self, img_h, txt_h):
        """Everything of one evaluation, asynchronously: windowed H2D on the copy stream, casts, tensor-core
        passes and finalizes on the current stream, D2H of the counts into self.host."""
        cur = torch.cuda.current_stream()
        cs = self.copy_stream
        cs.wait_stream(cur)  # the previous run has consumed the staging buffers
        with torch.cuda.stream(cs):
            self.img.copy_(img_h, non_blocking=True)
            self.ev_img.record(cs)
            for w, (b, e) in enumerate(self.bounds):
                self.txt[b:e].copy_(txt_h[b:e], non_blocking=True)
                self.ev_win[w].record(cs)
        st = cur.cuda_stream
        self.small.zero_()
        self.counts.zero_()
        cur.wait_event(self.ev_img)
        self._prep(self.img, self.img16, 0, self.n_img, self.rn[0], self.rn[1], self.small.data_ptr(), st)
        for w, (b, e) in enumerate(self.bounds):
            cur.wait_event(self.ev_win[w])
            self._prep(self.txt, self.txt16, b, e, self.rn[2], self.rn[3], self.small.data_ptr() + 16, st)
            pr, so, n = self.calls[w]
            N.check(self.lib.leccr_sim_topk_stream(pr, so, n, self.dim, self.fmt, self.k, st), "leccr_sim_topk_stream")
        pr, so, n = self.calls[-1]
        N.check(self.lib.leccr_sim_topk_stream(pr, so, n, self.dim, self.fmt, self.k, st), "leccr_sim_topk_stream")
        cur.wait_stream(cs)
        torch.cat([self.counts[0:3].float(), self.counts[4:7].float(), self.small[3:4], self.small[7:8]], out=self.out)
        self.host.copy_(self.out, non_blocking=True)

    @torch.no_grad()
    def run(self, image_embeds, text_embeds, return_topk=False, graph=True):
        """One evaluation.  With pinned host tensors the whole sequence (H2D windows included) is captured in
        a CUDA graph on first use and replayed whenever the SAME pinned buffers are passed again (serving:
        refill the buffers in place); other inputs take the eager path."""
        img_h = torch.as_tensor(image_embeds)
        txt_h = torch.as_tensor(text_embeds)
        if img_h.shape != self.img.shape or txt_h.shape != self.txt.shape or img_h.dtype != torch.float32 \
                or txt_h.dtype != torch.float32:
            raise N.LeccrError("StreamedEvalPlan.run needs fp32 [n_img, D] and [n_txt, D] inputs of the planned shape")
        pinned = (not img_h.is_cuda) and (not txt_h.is_cuda) and img_h.is_pinned() and txt_h.is_pinned() \
            and img_h.is_contiguous() and txt_h.is_contiguous()
        key = (img_h.data_ptr(), txt_h.data_ptr())
        graphs = self.__dict__.setdefault("_graphs", {})  # one graph per pair of pinned buffers (a few are kept)
        if graph and pinned and key in graphs:
            graphs[key][0].replay()
        elif graph and pinned:
            self._issue(img_h, txt_h)  # eager once: lazy initialisation must not happen under capture
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._issue(img_h, txt_h)
            while len(graphs) >= 4:
                graphs.pop(next(iter(graphs)))
            graphs[key] = (g, img_h, txt_h)  # the buffers stay referenced: the graph holds their addresses
            g.replay()
        else:
            self._issue(img_h, txt_h)
        torch.cuda.current_stream().synchronize()
        h = self.host.tolist()
        if h[6] != 0.0 or h[7] != 0.0:
            raise N.LeccrError("embeddings overflow the fp16 operand format; build the plan with precision='bf16'")
        ev = metrics_from_counts([int(c) for c in h[0:3]], self.n_img, [int(c) for c in h[3:6]], self.n_txt)
        return (ev, self.topk) if return_topk else ev


# ----------------------------------------------------------------------------- multi-GPU
@torch.no_grad()
def fused_eval_sharded(image_embeds, text_embeds, txt2img, img2txt, k=10, precision="f16", group=None):
    """Query-sharded fused evaluation: every rank holds the full embedding sets (as in the reference, where
    every rank evaluates everything, image_Retrieval_caption.py:453), ranks only ITS slice of the rows of
    each direction, and the six Recall counts are summed with one all_reduce.  Returns the same dict on
    every rank plus this rank's slices of the top-k lists."""
    from . import sharding

    dev = _device()
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    img = _to_device(image_embeds, dev)
    txt = _to_device(text_embeds, dev)
    fmt = ops.fmt_of(precision)
    I, T = ops.prep(img, fmt), ops.prep(txt, fmt)
    bi, ei = sharding.shard_range(img.shape[0], rank, world)
    bt, et = sharding.shard_range(txt.shape[0], rank, world)
    probs = []
    if ei > bi:
        probs.append((I.rows(bi, ei), T, ops.csr_from_lists([list(img2txt[i]) for i in range(bi, ei)], dev)))
    if et > bt:
        probs.append((T.rows(bt, et), I, ops.csr_from_lists([[txt2img[t]] for t in range(bt, et)], dev)))
    res = ops.sim_topk(probs, k=k) if probs else []
    counts = torch.zeros(6, dtype=torch.int32, device=dev)
    j = 0
    if ei > bi:
        counts[0:3] = res[j].recall_counts
        j += 1
    if et > bt:
        counts[3:6] = res[j].recall_counts
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    c = counts.cpu().tolist()
    ev = metrics_from_counts(c[0:3], img.shape[0], c[3:6], txt.shape[0])
    return ev, {"rows_i2t": (bi, ei), "rows_t2i": (bt, et), "results": res}


@torch.no_grad()
def topk_gallery_sharded(queries, gallery_shard, shard_offset: int, k=10, precision="bf16", group=None):
    """Row-partitioned gallery: every rank holds all queries and ITS gallery rows [shard_offset, ...).
    Local fused similarity + top-k, global column = local + shard_offset, then the exchange: on one NVLink
    node every rank pulls the other ranks' (Q, k) lists through peer pointers inside the merge kernel
    (leccr_topk_merge_peers); otherwise one all-gather and a merge (leccr_b200.sharding).
    Returns the same (val, idx int64) on every rank."""
    from . import sharding

    dev = _device()
    q = _to_device(queries, dev)
    g = _to_device(gallery_shard, dev)
    fmt = ops.fmt_of(precision)
    Q, G = ops.prep(q, fmt, want_stats=False), ops.prep(g, fmt, want_stats=False)
    res, = ops.sim_topk([(Q, G, None)], k=k)
    if group is None:
        from . import peer

        merged = peer.merge_topk_peers(res.val, res.idx, shard_offset, k)  # pull + merge in one kernel
        if merged is not None:
            return merged[0], merged[1].long()
    return sharding.allgather_topk(res.val, res.idx.long() + shard_offset, k, group)
