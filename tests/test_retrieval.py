"""Retrieval / zero-shot scoring.

CPU tier: the oracle restatement of the reference's itm_eval against the golden metric dictionaries that the
reference's own itm_eval produced (tests/golden/make_golden_retrieval.py), and against the live reference
function when /root/reference is present.
GPU tier: the matrix-free tensor-core path against the oracle (ranks, metric dictionaries, row argmax)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import retrieval_oracle as ro


def _load(path):
    z = np.load(path, allow_pickle=True)
    img2txt = {int(k): v for k, v in json.loads(str(z["img2txt"])).items()}
    txt2img = {j: int(v) for j, v in enumerate(z["txt2img"])}
    return (torch.from_numpy(z["image_embeds"]), torch.from_numpy(z["text_embeds"]), txt2img, img2txt,
            [int(x) for x in z["image_ids"]], json.loads(str(z["metrics"])))


def _cases(golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "retrieval_*.npz")))
    assert files, "retrieval golden vectors missing"
    return files


def test_oracle_matches_reference_golden(golden_dir):
    for path in _cases(golden_dir):
        img, txt, txt2img, img2txt, ids, want = _load(path)
        sims = (img @ txt.t()).numpy()
        got = ro.itm_eval(sims, sims.T, txt2img, img2txt, ids)
        assert got.keys() == want.keys()
        for k in want:
            assert abs(got[k] - want[k]) < 1e-9, (os.path.basename(path), k)


def test_oracle_matches_live_reference_itm_eval():
    if not os.path.isfile("/root/reference/retrieval.py"):
        pytest.skip("reference tree not present")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_retrieval import reference_itm_eval
    ref = reference_itm_eval()
    img, txt, txt2img, img2txt, ids = ro.synth_retrieval(64, 4, 32, seed=11, noise=5.0)
    sims = (img @ txt.t()).numpy()
    want = ref(sims, sims.T, txt2img, img2txt, torch.tensor(ids))
    got = ro.itm_eval(sims, sims.T, txt2img, img2txt, ids)
    for k in want:
        assert abs(got[k] - want[k]) < 1e-9, k


def test_split3_reproduces_fp32_scores():
    a = torch.nn.functional.normalize(torch.randn(50, 96), dim=-1)
    b = torch.nn.functional.normalize(torch.randn(70, 96), dim=-1)
    s3 = ro.split3(a, 0) @ ro.split3(b, 1).t()
    assert float((s3 - a.double() @ b.double().t()).abs().max()) < 3e-5
    # plain bf16 operands are ~100x worse: the reason the split exists
    s1 = a.bfloat16().double() @ b.bfloat16().double().t()
    assert float((s1 - a.double() @ b.double().t()).abs().max()) > 1e-4


# ------------------------------------------------------------------ GPU tier
def _kernel_vs_oracle_ranks(img, txt, row_targets, col_targets):
    """Kernel ranks vs fp64 ranks on the very operands the kernel multiplies; a mismatch is only legal where the
    fp32 accumulation of the tensor core can flip a comparison (a score within 1e-5 of the threshold)."""
    from clip_lite_b200 import retrieval as R
    r_row, r_col = R.retrieval_ranks(img.cuda(), txt.cuda(), row_targets, torch.tensor(col_targets))
    s = (ro.split3(img, 0) @ ro.split3(txt, 1).t()).numpy()
    want_row = ro.rank_above(s, row_targets)
    want_col = ro.rank_above(s.T, [[c] for c in col_targets])
    for got, want, mat, tg in ((r_row.cpu().numpy(), want_row, s, row_targets),
                               (r_col.cpu().numpy(), want_col, s.T, [[c] for c in col_targets])):
        bad = np.nonzero(got != want)[0]
        for i in bad:
            thr = max(mat[i, t] for t in tg[i])
            others = np.delete(mat[i], tg[i])
            near = int((np.abs(others - thr) < 1e-5).sum())
            assert abs(int(got[i]) - int(want[i])) <= near, (i, got[i], want[i], near)
        assert len(bad) <= max(1, len(got) // 100)
    return r_row, r_col


@pytest.mark.gpu
def test_ranks_match_oracle_on_golden_sets(golden_dir):
    from clip_lite_b200 import retrieval as R
    for path in _cases(golden_dir):
        img, txt, txt2img, img2txt, ids, want = _load(path)
        img2idx = {i: k for k, i in enumerate(ids)}
        rows = [img2txt[i] for i in ids]
        cols = [img2idx[txt2img[j]] for j in range(txt.shape[0])]
        _kernel_vs_oracle_ranks(img, txt, rows, cols)
        got = R.itm_eval(img.cuda(), txt.cuda(), txt2img, img2txt, torch.tensor(ids))
        assert got.keys() == want.keys()
        for k in want:                       # one near-tie may move one item across a recall cut-off
            assert abs(got[k] - want[k]) <= 100.0 / min(img.shape[0], txt.shape[0]) + 1e-9, (os.path.basename(path), k)


@pytest.mark.gpu
@pytest.mark.parametrize("n_img,caps,dim", [(1000, 5, 256), (777, 3, 512), (130, 1, 64), (5000, 5, 256)])
def test_ranks_ragged_and_coco_sized(n_img, caps, dim):
    """Edge tiles (sizes that are no multiple of the 256 x 256 tile), one caption per image, and the COCO 5k
    test-set shape (5000 images x 25000 captions)."""
    img, txt, txt2img, img2txt, ids = ro.synth_retrieval(n_img, caps, dim, seed=3, noise=8.0)
    img2idx = {i: k for k, i in enumerate(ids)}
    rows = [img2txt[i] for i in ids]
    cols = [img2idx[txt2img[j]] for j in range(txt.shape[0])]
    r_row, r_col = _kernel_vs_oracle_ranks(img, txt, rows, cols)
    assert int(r_row.min()) >= 0 and int(r_row.max()) < txt.shape[0]
    assert int(r_col.min()) >= 0 and int(r_col.max()) < n_img


@pytest.mark.gpu
def test_one_sided_and_missing_targets():
    from clip_lite_b200 import retrieval as R
    img, txt, txt2img, img2txt, ids = ro.synth_retrieval(300, 2, 64, seed=5, noise=6.0)
    img2idx = {i: k for k, i in enumerate(ids)}
    rows = [img2txt[i] for i in ids]
    cols = [img2idx[txt2img[j]] for j in range(txt.shape[0])]
    both = R.retrieval_ranks(img.cuda(), txt.cuda(), rows, torch.tensor(cols))
    only_r, none_c = R.retrieval_ranks(img.cuda(), txt.cuda(), rows, None)
    none_r, only_c = R.retrieval_ranks(img.cuda(), txt.cuda(), None, torch.tensor(cols))
    assert none_c is None and none_r is None
    assert torch.equal(only_r, both[0]) and torch.equal(only_c, both[1])      # deterministic integer counts
    # a row / column without ground truth is never "retrieved" (reference: rank = 1e20, retrieval.py:166)
    rows[7] = []
    cols[3] = -1
    r, c = R.retrieval_ranks(img.cuda(), txt.cuda(), rows, torch.tensor(cols))
    assert int(r[7]) == R.NO_TARGET_RANK and int(c[3]) == R.NO_TARGET_RANK
    keep_r = torch.ones(len(rows), dtype=torch.bool); keep_r[7] = False
    keep_c = torch.ones(len(cols), dtype=torch.bool); keep_c[3] = False
    assert torch.equal(r.cpu()[keep_r], both[0].cpu()[keep_r]) and torch.equal(c.cpu()[keep_c], both[1].cpu()[keep_c])
    assert R.recall_metrics(r, c)["txt_r10"] <= R.recall_metrics(*both)["txt_r10"]
    cols[3] = -2                                                              # invalid sentinel
    with pytest.raises(ValueError):
        R.retrieval_ranks(img.cuda(), txt.cuda(), rows, torch.tensor(cols))
    cols[3] = len(rows)                                                       # beyond the last image row
    with pytest.raises(ValueError):
        R.retrieval_ranks(img.cuda(), txt.cuda(), rows, torch.tensor(cols))
    cols[3] = -1
    with pytest.raises(ValueError):
        R.retrieval_ranks(img.cuda(), txt.cuda(), None, None)
    with pytest.raises(RuntimeError):
        R.retrieval_ranks(img, txt, rows, None)                               # no CPU path


@pytest.mark.gpu
@pytest.mark.parametrize("n,classes,dim", [(2048, 1000, 512), (333, 10, 64), (64, 100, 1024)])
def test_zero_shot_argmax_matches_torch_max(n, classes, dim):
    from clip_lite_b200 import retrieval as R
    gen = torch.Generator("cpu").manual_seed(n)
    feats = torch.nn.functional.normalize(torch.randn(n, dim, generator=gen), dim=-1)
    prompts = torch.nn.functional.normalize(torch.randn(classes, dim, generator=gen), dim=-1)
    val, pred = R.score_argmax(feats.cuda(), prompts.cuda())
    s = ro.split3(feats, 0) @ ro.split3(prompts, 1).t()
    want_val, want_pred = torch.max(s, 1)                                     # zero_shot.py:155
    pred, val = pred.cpu(), val.cpu().double()
    assert float((val - want_val).abs().max()) < 1e-5
    bad = torch.nonzero(pred != want_pred).flatten()
    for i in bad:                                                             # only a numerical tie may differ
        assert abs(float(s[i, pred[i]] - want_val[i])) < 1e-5
    assert len(bad) <= max(1, n // 200)
    # F.normalize fused into the operand split gives the same prediction from un-normalised features
    _, pred2 = R.score_argmax((3.0 * feats).cuda(), (0.5 * prompts).cuda(), normalize=True)
    assert float((pred2.cpu() != pred).float().mean()) < 0.01


# ------------------------------------------------------------------ host logic of the drop-in (CPU tier)
def test_itm_eval_host_logic_with_oracle_ranks(monkeypatch, golden_dir):
    """clip_lite_b200.retrieval.itm_eval = id bookkeeping + rank kernel + recall formulae.  With the rank kernel
    replaced by the oracle's ranks (the CUDA library cannot run here) the bookkeeping and the formulae must
    reproduce the reference's golden metric dictionaries exactly."""
    from clip_lite_b200 import retrieval as R
    seen = {}

    def fake_ranks(image_embeds, text_embeds, row_targets, col_targets, normalize, precision):
        s = (image_embeds.double() @ text_embeds.double().t()).numpy()
        seen["rows"], seen["cols"] = row_targets, col_targets
        r = ro.rank_above(s, row_targets)
        c = ro.rank_above(s.T, [[int(x)] for x in col_targets])
        return torch.from_numpy(r), torch.from_numpy(c)

    monkeypatch.setattr(R, "retrieval_ranks", fake_ranks)
    for path in _cases(golden_dir):
        img, txt, txt2img, img2txt, ids, want = _load(path)
        got = R.itm_eval(img, txt, txt2img, img2txt, torch.tensor(ids))
        assert got.keys() == want.keys()
        for k in want:
            assert abs(got[k] - want[k]) < 1e-9, (os.path.basename(path), k)
        assert len(seen["rows"]) == img.shape[0] and len(seen["cols"]) == txt.shape[0]


def test_csr_and_argument_checks():
    from clip_lite_b200 import retrieval as R
    ptr, idx = R._csr([[3, 1], [], [2]], "cpu")
    assert ptr.tolist() == [0, 2, 2, 3] and idx.tolist() == [3, 1, 2]
    assert ptr.dtype == torch.int32 and idx.dtype == torch.int32
    with pytest.raises(RuntimeError):                       # CPU tensors are refused: there is no CPU path
        R.prepare_operand(torch.zeros(4, 8), 0)
    m = R.recall_metrics(torch.tensor([0, 3, 7, 12]), torch.tensor([0, 0, 5, 9]))
    assert m["txt_r1"] == 25.0 and m["txt_r5"] == 50.0 and m["txt_r10"] == 75.0
    assert m["img_r1"] == 50.0 and m["img_r5"] == 50.0 and m["img_r10"] == 100.0
    assert abs(m["r_mean"] - ((25 + 50 + 75) / 3 + (50 + 50 + 100) / 3) / 2) < 1e-12
