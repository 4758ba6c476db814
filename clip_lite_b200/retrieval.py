"""Matrix-free retrieval / zero-shot scoring.

The reference evaluates retrieval by materialising ``sims_matrix = image_embeds @ text_embeds.t()``
(retrieval.py:143), copying it to the host and running a NumPy ``argsort`` per row and per column
(``itm_eval``, retrieval.py:151-209); zero-shot classification takes ``torch.max(features @ prompts.t(), 1)``
(zero_shot.py:155).  Here the same quantities come out of the tensor-core score kernel's epilogue:

    rank_i2t[i] = #{ j : S_ij > max_{t in img2txt[i]} S_it }      (position of the best ground-truth caption)
    rank_t2i[j] = #{ i : S_ij > S_{img(j), j} }                  (position of the ground-truth image)
    argmax_j S_ij

so the N x M matrix never exists, in HBM or on the host.  ``itm_eval`` below returns the reference's metric
dictionary (same keys, same formulae).  Scores are fp32-faithful by default (``precision="bf16x3"``: each
operand is split into bf16 hi + lo parts, three products on the tensor cores, error ~2^-17); ties are resolved
as "not ranked above", where NumPy's unstable argsort is order-dependent.

No CPU path: inputs must be CUDA tensors.
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Sequence, Tuple

import torch

from . import _lib
from . import kernels as K


NO_TARGET_RANK = 2 ** 31 - 1     # rank reported for a row / column that has no ground truth (reference: 1e20)


def prepare_operand(x: torch.Tensor, side: int, normalize: bool = False, precision: str = "bf16x3") -> torch.Tensor:
    """bf16 operand of the score kernel: [rows, 3 D] split form (side 0 = images/rows, 1 = texts/columns) or,
    with precision="bf16", the plain bf16 rounding [rows, D]."""
    K._req(x, "embeddings", ndim=2)
    rows, d = x.shape
    if precision == "bf16x3":
        if d % 8:
            raise ValueError(f"the embedding dimension must be a multiple of 8 (got {d})")
        out = torch.empty(rows, 3 * d, dtype=torch.bfloat16, device=x.device)
        with K._on_device(x.device):
            _lib.call("jsd_split_bf16x3", x.data_ptr(), K._code(x), rows, d, side, int(normalize), out.data_ptr(),
                      K._stream())
        return out
    if precision == "bf16":
        if d % 8:
            raise ValueError(f"the embedding dimension must be a multiple of 8 (got {d})")
        if normalize:
            return K.normalize_cast(x)[0]
        return x.to(torch.bfloat16).contiguous()
    raise ValueError(f"unknown precision {precision!r}; expected 'bf16x3' or 'bf16'")


def _csr(lists: Sequence[Sequence[int]], device) -> Tuple[torch.Tensor, torch.Tensor]:
    ptr = [0]
    idx = []
    for row in lists:
        idx.extend(int(t) for t in row)
        ptr.append(len(idx))
    return (torch.tensor(ptr, dtype=torch.int32, device=device),
            torch.tensor(idx if idx else [0], dtype=torch.int32, device=device))


def retrieval_ranks(image_embeds: torch.Tensor, text_embeds: torch.Tensor,
                    row_targets: Optional[Sequence[Sequence[int]]] = None,
                    col_targets: Optional[torch.Tensor] = None, normalize: bool = False,
                    precision: str = "bf16x3") -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """(rank_i2t [n_img] int32, rank_t2i [n_txt] int32); either is None when its targets are not given.
    row_targets[i] = text indices that are ground truth for image i; col_targets[j] = image index of text j."""
    if image_embeds.shape[1] != text_embeds.shape[1]:
        raise ValueError("image and text embeddings disagree on the dimension")
    if row_targets is None and col_targets is None:
        raise ValueError("no targets given")
    dev = image_embeds.device
    a = prepare_operand(image_embeds, 0, normalize, precision)
    b = prepare_operand(text_embeds, 1, normalize, precision)
    m, n, k = a.shape[0], b.shape[0], a.shape[1]
    ptr = idx = ct = thr_r = thr_c = rank_r = rank_c = None
    if row_targets is not None:
        if len(row_targets) != m:
            raise ValueError(f"row_targets has {len(row_targets)} entries for {m} images")
        ptr, idx = _csr(row_targets, dev)
        if int(idx.max()) >= n or int(idx.min()) < 0:
            raise ValueError("row target out of range")
        thr_r = torch.empty(m, dtype=torch.int32, device=dev)
        rank_r = torch.empty(m, dtype=torch.int32, device=dev)
    if col_targets is not None:
        ct = torch.as_tensor(col_targets).to(device=dev, dtype=torch.int32).contiguous()
        if ct.numel() != n:
            raise ValueError(f"col_targets has {ct.numel()} entries for {n} texts")
        if int(ct.max()) >= m or int(ct.min()) < -1:
            raise ValueError("column target out of range (valid: image rows 0..M-1, or -1 for 'no ground-truth image')")
        thr_c = torch.empty(n, dtype=torch.float32, device=dev)
        rank_c = torch.empty(n, dtype=torch.int32, device=dev)
    with K._on_device(dev):
        _lib.call("jsd_score_ranks", a.data_ptr(), b.data_ptr(), m, n, k, K._ptr(ptr), K._ptr(idx), K._ptr(ct),
                  K._ptr(thr_r), K._ptr(thr_c), K._ptr(rank_r), K._ptr(rank_c), K._stream())
    # a row / column without any ground truth never counts as retrieved: the reference leaves such an image at
    # rank = 1e20 (retrieval.py:166), i.e. outside every recall@k
    if rank_r is not None:
        rank_r.masked_fill_(ptr[1:] == ptr[:-1], NO_TARGET_RANK)
    if rank_c is not None:
        rank_c.masked_fill_(ct < 0, NO_TARGET_RANK)
    return rank_r, rank_c


def score_argmax(features: torch.Tensor, prompt_features: torch.Tensor, normalize: bool = False,
                 precision: str = "bf16x3") -> Tuple[torch.Tensor, torch.Tensor]:
    """(max_j S_ij, argmax_j S_ij) per row -- ``torch.max(features @ prompt_features.t(), 1)`` of zero_shot.py:155
    without the matrix; ties go to the smallest column."""
    a = prepare_operand(features, 0, normalize, precision)
    b = prepare_operand(prompt_features, 1, normalize, precision)
    m, n, k = a.shape[0], b.shape[0], a.shape[1]
    best = torch.empty(m, dtype=torch.int64, device=a.device)
    with K._on_device(a.device):
        _lib.call("jsd_score_argmax", a.data_ptr(), b.data_ptr(), m, n, k, best.data_ptr(), K._stream())
    col = 0xFFFFFFFF - (best & 0xFFFFFFFF)
    enc = (best >> 32) & 0xFFFFFFFF                     # order-preserving encoding of the fp32 maximum
    bits = torch.where(enc >= 0x80000000, enc ^ 0x80000000, enc ^ 0xFFFFFFFF)
    val = torch.where(bits >= 0x80000000, bits - (1 << 32), bits).to(torch.int32).view(torch.float32)
    return val, col


def zero_shot_predict(features: torch.Tensor, prompt_features: torch.Tensor, normalize: bool = False,
                      precision: str = "bf16x3") -> torch.Tensor:
    """Predicted class per row (zero_shot.py:155)."""
    return score_argmax(features, prompt_features, normalize, precision)[1]


def recall_metrics(rank_i2t: torch.Tensor, rank_t2i: torch.Tensor) -> Dict[str, float]:
    """The reference's metric dictionary (retrieval.py:178-209) from the two rank vectors."""
    def at(r, k):
        return 100.0 * float((r < k).sum()) / r.numel()
    tr1, tr5, tr10 = (at(rank_i2t, k) for k in (1, 5, 10))
    ir1, ir5, ir10 = (at(rank_t2i, k) for k in (1, 5, 10))
    tr_mean = (tr1 + tr5 + tr10) / 3
    ir_mean = (ir1 + ir5 + ir10) / 3
    return {"txt_r1": tr1, "txt_r5": tr5, "txt_r10": tr10, "txt_r_mean": tr_mean,
            "img_r1": ir1, "img_r5": ir5, "img_r10": ir10, "img_r_mean": ir_mean,
            "r_mean": (tr_mean + ir_mean) / 2}


def itm_eval(image_embeds: torch.Tensor, text_embeds: torch.Tensor, txt2img: Mapping[int, int],
             img2txt: Mapping[int, Sequence[int]], image_ids, normalize: bool = False,
             precision: str = "bf16x3") -> Dict[str, float]:
    """Drop-in for ``itm_eval(scores_i2t, scores_t2i, txt2img, img2txt, image_ids)`` (retrieval.py:151-209) that
    takes the embeddings instead of the two host score matrices.  ``image_ids[idx]`` is the dataset id of image
    row idx; ``img2txt[id]`` its caption rows; ``txt2img[j]`` the image id of caption row j."""
    ids = [int(x) for x in (image_ids.tolist() if hasattr(image_ids, "tolist") else image_ids)]
    img2idx = {img_id: idx for idx, img_id in enumerate(ids)}
    rows = [list(img2txt[i]) for i in ids]
    n_txt = text_embeds.shape[0]
    cols = torch.tensor([img2idx[int(txt2img[j])] for j in range(n_txt)], dtype=torch.int32)
    r_i2t, r_t2i = retrieval_ranks(image_embeds, text_embeds, rows, cols, normalize, precision)
    return recall_metrics(r_i2t, r_t2i)
