"""Seeded synthetic embeddings for the five BASELINE.json configurations.

The hot path consumes (B, 256) unit vectors plus the ground-truth maps the reference's datasets build
(dataset/retrieval_dataset.py:208-226): txt2img[t] -> image, img2txt[i] -> list of texts.  Everything is
generated on the CPU with an explicit torch.Generator so tests, bench and the golden script agree.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

EMBED_DIM = 256  # configs/cclm-base-ft/*.yaml: embed_dim
TEMP = 0.07      # models/xvlm.py:177


@dataclass
class RetrievalSet:
    image: torch.Tensor                 # [N, D] unit rows
    text: torch.Tensor                  # [M, D] unit rows
    txt2img: Dict[int, int]
    img2txt: Dict[int, List[int]]
    caption: Optional[torch.Tensor] = None  # [n, N, D] un-normalised (video double_sim)


def _noise_scale(d):
    return 1.5 * 4.0 / d ** 0.5


def retrieval_set(n_img: int, texts_per_img: int, d: int = EMBED_DIM, seed: int = 1234,
                  n_caption_queries: int = 0) -> RetrievalSet:
    """Images = normalize(randn); texts = normalize(image[t // per] + sigma * randn): Recall is unsaturated."""
    g = torch.Generator().manual_seed(seed)
    img = F.normalize(torch.randn(n_img, d, generator=g), dim=-1)
    m = n_img * texts_per_img
    owner = torch.arange(m) // texts_per_img
    txt = F.normalize(img[owner] + _noise_scale(d) * torch.randn(m, d, generator=g), dim=-1)
    txt2img = {t: int(owner[t]) for t in range(m)}
    img2txt = {i: list(range(i * texts_per_img, (i + 1) * texts_per_img)) for i in range(n_img)}
    cap = None
    if n_caption_queries:
        cap = img.unsqueeze(0) + 0.1 * torch.randn(n_caption_queries, n_img, d, generator=g)
    return RetrievalSet(img, txt, txt2img, img2txt, cap)


def cfg1_multi30k(d: int = EMBED_DIM) -> RetrievalSet:
    """Multi30K-like: 1,000 images x 5,000 captions."""
    return retrieval_set(1000, 5, d, seed=1234)


def cfg2_mscoco5k(d: int = EMBED_DIM) -> RetrievalSet:
    """MSCOCO-5K-like: 5,000 images x 25,000 captions."""
    return retrieval_set(5000, 5, d, seed=1235)


def cfg4_msrvtt(d: int = EMBED_DIM) -> RetrievalSet:
    """MSR-VTT-CN-like: 1,000 videos x 1,000 queries, n = 2 caption queries (Retrieval_msrvtt.yaml:47)."""
    return retrieval_set(1000, 1, d, seed=1236, n_caption_queries=2)


@dataclass
class ContrastiveBatch:
    image: torch.Tensor  # [N, D] global batch, rank r owns rows [r * B, (r + 1) * B)
    text: torch.Tensor
    idx: torch.Tensor    # [N] int64, about two positives per class
    temp: float = TEMP


def cfg3_itc(n_global: int = 4096, d: int = EMBED_DIM, seed: int = 7) -> ContrastiveBatch:
    """ITC step: global batch 4096 (8 ranks x 512); text correlated with image so the loss is not ln N."""
    g = torch.Generator().manual_seed(seed)
    a = F.normalize(torch.randn(n_global, d, generator=g), dim=-1)
    b = F.normalize(a + 0.5 * 4.0 / d ** 0.5 * torch.randn(n_global, d, generator=g), dim=-1)
    idx = torch.randint(0, max(1, n_global // 2), (n_global,), generator=g)
    return ContrastiveBatch(a, b, idx)


def cfg5_gallery(n_gallery: int, n_query: int, d: int = EMBED_DIM, seed: int = 1237, device="cpu",
                 dtype=torch.bfloat16):
    """Large-gallery sweep: queries = normalize(gallery[gt] + sigma * randn); returns (gallery, queries, gt)."""
    g = torch.Generator(device=device).manual_seed(seed)
    gal = F.normalize(torch.randn(n_gallery, d, generator=g, device=device), dim=-1)
    gt = torch.randint(0, n_gallery, (n_query,), generator=g, device=device)
    qry = F.normalize(gal[gt] + _noise_scale(d) * torch.randn(n_query, d, generator=g, device=device), dim=-1)
    return gal.to(dtype), qry.to(dtype), gt
