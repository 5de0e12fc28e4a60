"""TEST INFRASTRUCTURE -- CPU restatement of the reference's dense cross-modal similarity path.

This is the checker the parity tests, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg use.
It is never imported by the product (leccr_b200/), which has no CPU path.

Parity status: PINNED.  The reference holds no golden vectors (it has no tests at all), so the oracle
is pinned against outputs of the reference's OWN functions executed in the build container on seeded
inputs: `oracle/make_golden.py` imports them from /root/reference/LECCR through `oracle/ref_loader.py`
and commits inputs + outputs under tests/golden/; tests/test_oracle_golden.py replays them.

Every function cites the reference lines it restates (paths relative to /root/reference/LECCR/).
The arithmetic is the reference's: fp32 torch / numpy on the host.
"""
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- AllGather
def allgather_forward(per_rank: Sequence[torch.Tensor]) -> torch.Tensor:
    """models/xvlm.py:53-59: all_gather into world_size buffers, torch.cat(dim 0) in rank order."""
    return torch.cat(list(per_rank), 0)


def allgather_backward(grad_output: torch.Tensor, rank: int, batch_size: int) -> torch.Tensor:
    """models/xvlm.py:62-67: the gradient of the gather is the rank's own slice, no reduction."""
    return grad_output[batch_size * rank: batch_size * (rank + 1)]


# ----------------------------------------------------------------------------- contrastive loss
def contrastive_loss(image_feat_all: torch.Tensor, text_feat_all: torch.Tensor, temp,
                     idx_all: Optional[torch.Tensor] = None) -> torch.Tensor:
    """models/xvlm.py:273-292 on the already gathered features (single-process concatenated batch).

    logits = A B^T / temp; idx None -> arange labels + cross_entropy both ways (:277-280); else soft
    labels pos / pos.sum(1) from idx equality and -sum(log_softmax * labels).mean() both ways (:283-290).
    """
    logits = image_feat_all @ text_feat_all.t() / temp
    bsz = image_feat_all.shape[0]
    if idx_all is None:
        labels = torch.arange(bsz, device=image_feat_all.device)
        loss_i2t = F.cross_entropy(logits, labels)
        loss_t2i = F.cross_entropy(logits.t(), labels)
    else:
        idx_all = idx_all.view(-1, 1)
        pos_idx = torch.eq(idx_all, idx_all.t()).to(logits.dtype)
        labels = pos_idx / pos_idx.sum(1, keepdim=True)
        loss_i2t = -torch.sum(F.log_softmax(logits, dim=1) * labels, dim=1).mean()
        loss_t2i = -torch.sum(F.log_softmax(logits.t(), dim=1) * labels, dim=1).mean()
    return (loss_i2t + loss_t2i) / 2


def contrastive_loss_and_grads(image_feat_all, text_feat_all, temp: float, idx_all=None, rank: int = 0,
                               batch_size: Optional[int] = None, dtype=torch.float32):
    """Loss plus what autograd hands back to rank `rank`: the local slices of dA, dB (AllGather.backward)
    and the full d loss / d temp.  dtype=float64 gives the high-precision answer tolerances are set from."""
    a = image_feat_all.detach().to(dtype).clone().requires_grad_(True)
    b = text_feat_all.detach().to(dtype).clone().requires_grad_(True)
    t = torch.tensor(float(temp), dtype=dtype, requires_grad=True)
    loss = contrastive_loss(a, b, t, idx_all)
    loss.backward()
    n = a.shape[0]
    bs = n if batch_size is None else batch_size
    return (loss.detach(), allgather_backward(a.grad, rank, bs), allgather_backward(b.grad, rank, bs),
            t.grad.detach())


def caption_contrastive_loss(caption_embeds: torch.Tensor, text_feats: torch.Tensor, temp) -> torch.Tensor:
    """models/model_retrieval_caption.py:145-152 (== video_model_retrieval_caption.py:171-178), statement for
    statement: local batch only, max over the n caption queries, arange labels, symmetric cross entropy."""
    n, bsz, d = caption_embeds.shape
    sim = caption_embeds.reshape(-1, d) @ text_feats.transpose(0, 1)
    logits = torch.max(sim.reshape(n, bsz, bsz), dim=0)[0] / temp
    labels = torch.arange(bsz, device=caption_embeds.device)
    loss_i2t = F.cross_entropy(logits, labels)
    loss_t2i = F.cross_entropy(logits.t(), labels)
    return (loss_t2i + loss_i2t) / 2


def caption_contrastive_loss_and_grads(caption_embeds, text_feats, temp: float, dtype=torch.float32):
    """Loss and autograd gradients (d caption, d text, d temp); dtype=float64 for tolerance setting."""
    c = caption_embeds.detach().to(dtype).clone().requires_grad_(True)
    t = text_feats.detach().to(dtype).clone().requires_grad_(True)
    tp = torch.tensor(float(temp), dtype=dtype, requires_grad=True)
    loss = caption_contrastive_loss(c, t, tp)
    loss.backward()
    return loss.detach(), c.grad, t.grad, tp.grad.detach()


def norm_score_pos(score: torch.Tensor) -> torch.Tensor:
    """models/model_retrieval_caption.py:87-90 (the POSITIVE variant used in training): (x - min) / max(x - min)."""
    score = score - torch.min(score)
    score = score / torch.max(score)
    return score


def dstl_loss(image_all, caption_all, text_s_all, text_t_all, alpha: float = 0.8) -> torch.Tensor:
    """models/model_retrieval_caption.py:94-116 on the ALL-GATHERED tensors (the gathers themselves are
    allgather_forward), statement for statement, including the reference's mixed orientation of the two label
    terms (logits_sv is text x image, logits_sc is caption-sample x text)."""
    logits_tv = text_t_all @ image_all.t()
    logits_sv = text_s_all @ image_all.t()
    n, bsz, d = caption_all.shape
    sim = caption_all.reshape(-1, d) @ text_s_all.transpose(0, 1)
    logits_sc = torch.max(sim.reshape(n, bsz, bsz), dim=0)[0]
    logits_sc = norm_score_pos(logits_sc)
    logits_sv = norm_score_pos(logits_sv)
    labels = alpha * logits_sv + (1. - alpha) * logits_sc
    labels = F.softmax(labels, 1)
    logits_tv = F.log_softmax(logits_tv, 1)
    return F.kl_div(logits_tv, labels.detach(), reduction='batchmean')


def dstl_loss_and_grads(image_all, caption_all, text_s_all, text_t_all, alpha: float = 0.8, rank: int = 0,
                        batch_size: Optional[int] = None, dtype=torch.float32):
    """Loss and what autograd hands back to rank `rank`: the local rows of d image and d text_t (labels are
    detached, so text_s and the captions get no gradient)."""
    im = image_all.detach().to(dtype).clone().requires_grad_(True)
    tt = text_t_all.detach().to(dtype).clone().requires_grad_(True)
    loss = dstl_loss(im, caption_all.detach().to(dtype), text_s_all.detach().to(dtype), tt, alpha)
    loss.backward()
    bs = im.shape[0] if batch_size is None else batch_size
    return loss.detach(), allgather_backward(im.grad, rank, bs), allgather_backward(tt.grad, rank, bs)


def caption_vision_loss(caption_all, image_all, idx_all, cproj, vproj) -> torch.Tensor:
    """models/model_retrieval_caption.py:122-143 on the ALL-GATHERED token tensors (caption_all is [N, cn, d], i.e.
    after the reference's transpose and gather), statement for statement: F.normalize keeps the reference's
    default dim=1 (the token axis), the token-pair similarities are formed and averaged as the reference does."""
    caption = F.normalize(cproj(caption_all))
    image = F.normalize(vproj(image_all))
    bsz, vn, d = image.shape
    _, cn, _ = caption.shape
    _image = image.reshape(-1, d)
    _caption = caption.reshape(-1, d)
    sim = _caption @ _image.t()
    sim = sim.reshape(bsz, cn, bsz, vn).transpose(1, 2)
    sim = torch.mean(torch.mean(sim, dim=-1), dim=-1)
    idx = idx_all.view(-1, 1)
    pos_idx = torch.eq(idx, idx.t()).float()
    labels = pos_idx / pos_idx.sum(1, keepdim=True)
    return -torch.sum(F.log_softmax(sim, dim=1) * labels, dim=1).mean()


def caption_vision_loss_and_grads(caption, image, idx, Wc, bc, Wv, bv, rank: int = 0, batch_size: Optional[int] = None,
                                  dtype=torch.float32):
    """caption: [cn, N, d] as the model holds it.  Returns loss, the local rows of d image and d caption
    (AllGather.backward keeps the local slices) and this rank's gradients of the two projection weights."""
    N_ = image.shape[0]
    bs = N_ if batch_size is None else batch_size
    sl = slice(rank * bs, (rank + 1) * bs)
    d = image.shape[2]
    cproj, vproj = torch.nn.Linear(d, d).to(dtype), torch.nn.Linear(d, d).to(dtype)
    with torch.no_grad():
        cproj.weight.copy_(Wc.to(dtype)); cproj.bias.copy_(bc.to(dtype))
        vproj.weight.copy_(Wv.to(dtype)); vproj.bias.copy_(bv.to(dtype))
    im_loc = image[sl].detach().to(dtype).clone().requires_grad_(True)
    cp_loc = caption[:, sl].detach().to(dtype).clone().requires_grad_(True)
    # other ranks' rows take part in the forward but return no gradient to THIS rank
    im_all = torch.cat([image[:sl.start].to(dtype), im_loc, image[sl.stop:].to(dtype)], 0)
    cp_all = torch.cat([caption[:, :sl.start].to(dtype), cp_loc, caption[:, sl.stop:].to(dtype)], 1).transpose(0, 1)
    # the reference projects AFTER the gather (:123-124): every rank's module sees all N rows, so its parameter
    # gradients are the full ones; only the gradients of the gathered INPUTS are cut to the local slice
    loss = caption_vision_loss(cp_all, im_all, idx, cproj, vproj)
    loss.backward()
    return loss.detach(), im_loc.grad, cp_loc.grad, cproj.weight.grad, vproj.weight.grad


# ----------------------------------------------------------------------------- evaluation score matrices
def score_matrices(image_embeds: torch.Tensor, text_embeds: torch.Tensor):
    """image_Retrieval_caption.py:151-152,163: i2t = image @ text.T, t2i = its transpose VIEW, as numpy."""
    i2t = image_embeds @ text_embeds.t()
    t2i = i2t.t()
    return i2t.cpu().numpy(), t2i.cpu().numpy()


def norm_score(x: torch.Tensor) -> torch.Tensor:
    """video_Retrieval_caption_double_sim.py:87-91 == (x - max x) / (max x - min x), whole-matrix min/max."""
    s = -x
    s = s - torch.min(s)
    s = s / torch.max(s)
    return -s


def caption_scores(caption_embeds: torch.Tensor, text_embeds: torch.Tensor) -> torch.Tensor:
    """video_...double_sim.py:173-175: max over the n caption queries of caption @ text.T.
    The reference's reshape(n, bsz, bsz) is only valid when #videos == #texts; this is the general form
    (n, N_vid, N_txt) that coincides with it on the square case."""
    n, nv, d = caption_embeds.shape
    c_sim = caption_embeds.reshape(-1, d) @ text_embeds.t()
    return torch.max(c_sim.reshape(n, nv, text_embeds.shape[0]), dim=0)[0]


def double_sim_matrices(image_embeds, text_embeds, caption_embeds, alpha: float = 0.9, fusion: str = "norm"):
    """fusion 'norm': video_...double_sim.py:170-179 (alpha = 0.9, :95).
    fusion 'raw' : image_Retrieval_caption.py:235-246 (alpha = 0.8, no norm_score)."""
    s_i2t = image_embeds @ text_embeds.t()
    s_t2i = s_i2t.t()
    c_i2t = caption_scores(caption_embeds, text_embeds)
    c_t2i = c_i2t.t()
    if fusion == "norm":
        i2t = alpha * norm_score(s_i2t) + (1. - alpha) * norm_score(c_i2t)
        t2i = alpha * norm_score(s_t2i) + (1. - alpha) * norm_score(c_t2i)
    elif fusion == "raw":
        i2t = alpha * s_i2t + (1 - alpha) * c_i2t
        t2i = alpha * s_t2i + (1 - alpha) * c_t2i
    else:
        raise ValueError(fusion)
    return i2t.cpu().numpy(), t2i.cpu().numpy()


# ----------------------------------------------------------------------------- ranking / Recall@K
def itm_eval(scores_i2t: np.ndarray, scores_t2i: np.ndarray, txt2img: Dict[int, int],
             img2txt: Dict[int, List[int]]) -> dict:
    """image_Retrieval_caption.py:261-317 (== video_...double_sim.py:194-247), statement by statement."""
    ranks = np.zeros(scores_i2t.shape[0])
    for index, score in enumerate(scores_i2t):
        inds = np.argsort(score)[::-1]
        rank = 1e20
        for i in img2txt[index]:
            tmp = np.where(inds == i)[0][0]
            if tmp < rank:
                rank = tmp
        ranks[index] = rank
    tr1 = 100.0 * len(np.where(ranks < 1)[0]) / len(ranks)
    tr5 = 100.0 * len(np.where(ranks < 5)[0]) / len(ranks)
    tr10 = 100.0 * len(np.where(ranks < 10)[0]) / len(ranks)
    ranks = np.zeros(scores_t2i.shape[0])
    for index, score in enumerate(scores_t2i):
        inds = np.argsort(score)[::-1]
        ranks[index] = np.where(inds == txt2img[index])[0][0]
    ir1 = 100.0 * len(np.where(ranks < 1)[0]) / len(ranks)
    ir5 = 100.0 * len(np.where(ranks < 5)[0]) / len(ranks)
    ir10 = 100.0 * len(np.where(ranks < 10)[0]) / len(ranks)
    return metrics_from_recalls(tr1, tr5, tr10, ir1, ir5, ir10)


def metrics_from_recalls(tr1, tr5, tr10, ir1, ir5, ir10) -> dict:
    """image_Retrieval_caption.py:297-316: the seven derived numbers and the 13-key dict."""
    tr_mean = (tr1 + tr5 + tr10) / 3
    ir_mean = (ir1 + ir5 + ir10) / 3
    r_mean = (tr_mean + ir_mean) / 2
    txt_sumr = tr1 + tr5 + tr10
    img_sumr = ir1 + ir5 + ir10
    sumr_avg = np.round((txt_sumr + img_sumr) / 6, 2)
    return {'txt_r1': tr1, 'txt_r5': tr5, 'txt_r10': tr10, 'txt_r_mean': tr_mean, 'txt_sum_r': txt_sumr,
            'img_r1': ir1, 'img_r5': ir5, 'img_r10': ir10, 'img_r_mean': ir_mean, 'r_mean': r_mean,
            'img_sumr': img_sumr, 'sumr_avg': sumr_avg, 'sumr_sum': (txt_sumr + img_sumr)}


def ranks_by_count(scores: np.ndarray, gt: Sequence[Sequence[int]]) -> np.ndarray:
    """Closed form of the argsort position when no two scores in a row tie:
    rank(row) = min over the row's ground truth g of #{j : s_j > s_g}   (:268-278, :288-290)."""
    out = np.zeros(scores.shape[0], dtype=np.int64)
    for r in range(scores.shape[0]):
        row = scores[r]
        out[r] = min(int(np.count_nonzero(row > row[g])) for g in gt[r])
    return out


def recall_counts(ranks: np.ndarray) -> List[int]:
    return [int(np.count_nonzero(ranks < k)) for k in (1, 5, 10)]


def itm_eval_by_count(scores_i2t: np.ndarray, scores_t2i: np.ndarray, txt2img, img2txt) -> dict:
    """Same dict as itm_eval via ranks_by_count (identical whenever rows have no exact ties)."""
    r_i = ranks_by_count(scores_i2t, [img2txt[i] for i in range(scores_i2t.shape[0])])
    r_t = ranks_by_count(scores_t2i, [[txt2img[t]] for t in range(scores_t2i.shape[0])])
    n_i, n_t = len(r_i), len(r_t)
    c_i, c_t = recall_counts(r_i), recall_counts(r_t)
    return metrics_from_recalls(*(100.0 * c / n_i for c in c_i), *(100.0 * c / n_t for c in c_t))


def topk(scores: np.ndarray, k: int):
    """Top-k columns of every row, descending, ties by lower column (the order np.argsort(...)[::-1]
    leaves unspecified at ties)."""
    order = np.lexsort((np.arange(scores.shape[1])[None, :].repeat(scores.shape[0], 0), -scores), axis=1)[:, :k]
    return np.take_along_axis(scores, order, 1), order


# ----------------------------------------------------------------------------- large-gallery retrieval (cfg5)
def gallery_eval(gallery: torch.Tensor, queries: torch.Tensor, gt: Sequence[int], k: int = 10):
    """BASELINE.json configs[4]: every query ranks the whole gallery, one ground-truth gallery row per query.
    The reference would run it as its text->image direction: score = gallery @ queries.T
    (image_Retrieval_caption.py:151), t2i = its transpose (:152), then per query a full np.argsort and the
    position of the ground truth (:288-290), Recall@1/5/10 (:293-295).  16-bit inputs are scored in fp32, as
    `.float()` embeddings would be.  Returns (recall dict img_r1/5/10, top-k values, top-k gallery rows)."""
    t2i = (gallery.float() @ queries.float().t()).t().cpu().numpy()
    ranks = np.zeros(t2i.shape[0])
    top = np.zeros((t2i.shape[0], k), dtype=np.int64)
    for index, score in enumerate(t2i):
        inds = np.argsort(score)[::-1]
        ranks[index] = np.where(inds == int(gt[index]))[0][0]
        top[index] = inds[:k]
    ev = {f"img_r{c}": 100.0 * len(np.where(ranks < c)[0]) / len(ranks) for c in (1, 5, 10)}
    return ev, np.take_along_axis(t2i, top, 1), top


# ----------------------------------------------------------------------------- get_features
def get_features(tokens: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """models/xvlm.py:241-256 for one modality with the default 'cls' pooling:
    F.normalize(proj(embeds[:, 0, :]), dim=-1)."""
    return F.normalize(F.linear(tokens[:, 0, :], weight, bias), dim=-1)
