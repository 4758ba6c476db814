"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU restatement of the reference's retrieval and
zero-shot scoring.

Follows retrieval.py:143 (similarity matrix), retrieval.py:151-209 (``itm_eval``: per-row / per-column rank of
the ground truth through a descending argsort, recall@1/5/10 and their means) and zero_shot.py:155 (row argmax).
Pinned against the reference's own ``itm_eval`` executed in the build container: tests/golden/retrieval_*.npz,
written by tests/golden/make_golden_retrieval.py.
"""
from __future__ import annotations

from typing import Dict, Mapping, Sequence

import numpy as np
import torch


def similarity(image_embeds: np.ndarray, text_embeds: np.ndarray) -> np.ndarray:
    """retrieval.py:143: ``image_embeds @ text_embeds.t()`` (fp32 in the reference; any float dtype here)."""
    return image_embeds @ text_embeds.T


def rank_above(scores: np.ndarray, row_targets: Sequence[Sequence[int]]) -> np.ndarray:
    """Position of a row's best ground truth in the descending order of that row = number of entries strictly
    above the largest target score (retrieval.py:163-176: min over targets of ``np.where(argsort[::-1] == t)``)."""
    out = np.zeros(scores.shape[0], dtype=np.int64)
    for i, tg in enumerate(row_targets):
        thr = max(scores[i, t] for t in tg)
        out[i] = int((scores[i] > thr).sum())
    return out


def recall_dict(rank_i2t: np.ndarray, rank_t2i: np.ndarray) -> Dict[str, float]:
    """retrieval.py:178-209."""
    def at(r, k):
        return 100.0 * float((r < k).sum()) / len(r)
    tr = [at(rank_i2t, k) for k in (1, 5, 10)]
    ir = [at(rank_t2i, k) for k in (1, 5, 10)]
    tr_mean, ir_mean = sum(tr) / 3, sum(ir) / 3
    return {"txt_r1": tr[0], "txt_r5": tr[1], "txt_r10": tr[2], "txt_r_mean": tr_mean,
            "img_r1": ir[0], "img_r5": ir[1], "img_r10": ir[2], "img_r_mean": ir_mean,
            "r_mean": (tr_mean + ir_mean) / 2}


def itm_eval(scores_i2t: np.ndarray, scores_t2i: np.ndarray, txt2img: Mapping[int, int],
             img2txt: Mapping[int, Sequence[int]], image_ids: Sequence[int]) -> Dict[str, float]:
    """Same signature and result as the reference's itm_eval (retrieval.py:151-209)."""
    ids = [int(x) for x in image_ids]
    img2idx = {img_id: idx for idx, img_id in enumerate(ids)}
    r_i2t = rank_above(scores_i2t, [img2txt[i] for i in ids])
    r_t2i = rank_above(scores_t2i, [[img2idx[int(txt2img[j])]] for j in range(scores_t2i.shape[0])])
    return recall_dict(r_i2t, r_t2i)


def split3(x: torch.Tensor, side: int) -> torch.Tensor:
    """The bf16 hi/lo operand split of the scoring kernel, as float64 [rows, 3 D]: side 0 = (hi, hi, lo),
    side 1 = (hi, lo, hi); the row dot of a side-0 with a side-1 operand is <a, b> up to ~2^-17."""
    x = x.float()
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    parts = (hi, hi, lo) if side == 0 else (hi, lo, hi)
    return torch.cat([p.double() for p in parts], dim=1)


def synth_retrieval(n_img: int, caps_per_img: int, dim: int, seed: int, noise: float = 1.0):
    """Seeded toy retrieval set: unit image embeddings, `caps_per_img` noisy captions each, shuffled caption
    order, non-trivial image ids.  Returns (img [n_img, D], txt [n_txt, D], txt2img, img2txt, image_ids)."""
    gen = torch.Generator("cpu").manual_seed(seed)
    img = torch.nn.functional.normalize(torch.randn(n_img, dim, generator=gen), dim=-1)
    n_txt = n_img * caps_per_img
    owner = torch.arange(n_img).repeat_interleave(caps_per_img)[torch.randperm(n_txt, generator=gen)]
    txt = torch.nn.functional.normalize(img[owner] + noise * torch.randn(n_txt, dim, generator=gen) / dim ** 0.5,
                                        dim=-1)
    image_ids = (1000 + 7 * torch.randperm(n_img, generator=gen)).tolist()
    txt2img = {j: image_ids[int(owner[j])] for j in range(n_txt)}
    img2txt = {i: [] for i in image_ids}
    for j in range(n_txt):
        img2txt[txt2img[j]].append(j)
    return img, txt, txt2img, img2txt, image_ids
