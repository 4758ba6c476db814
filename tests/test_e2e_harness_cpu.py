"""The drop-in claim end to end (VERDICT r1 missing #6): the train-step harness of
tools/e2e_harness.py (restating model.py:32-113, factories.py:464-487, train.py:210-227 around an arbitrary loss module)
is run twice on the same tiny encoders, the same batches and the same seeds -- once with the UNMODIFIED reference
loss.py, once with clip_lite_b200.loss.JSDInfoMaxLoss loaded from the reference's state_dict -- and must leave the same
parameters and buffers behind after two optimiser steps.  The drop-in's row-wise entry points run in the CPU emulation
of the kernel source (tests/_emu_backend.py); with fused_heads=True the heads' tail goes through jsd_heads.cuh."""
import copy
import os
import sys

import pytest
import torch
from torch import nn

from oracle import reference_loader as rl
from tests import _emu_backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import e2e_harness as H  # noqa: E402

pytestmark = [pytest.mark.skipif(not rl.reference_available(), reason="needs /root/reference (build container)"),
              pytest.mark.skipif(not _emu_backend.available(), reason="needs g++ and the CUDA headers")]

IMG_DIM, TXT_DIM, B = 32, 24, 8


def tiny_encoders():
    from transformers import BertConfig
    torch.manual_seed(0)
    cnn = nn.Sequential(nn.Conv2d(3, 8, 3, stride=2, padding=1), nn.BatchNorm2d(8), nn.ReLU(),
                        nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(8, IMG_DIM))
    cfg = BertConfig(vocab_size=100, hidden_size=TXT_DIM, num_hidden_layers=1, num_attention_heads=2,
                     intermediate_size=32, max_position_embeddings=16, hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    return H.ImageEncoder(cnn), H.TextEncoder(config=cfg)


def batches(cluster):
    out = []
    for seed in (1, 2):
        b = H.synthetic_batch(B, "cpu", seed=seed, image_px=16, tokens=6, vocab=100)
        if cluster:
            n = H.synthetic_batch(B, "cpu", seed=seed + 10, image_px=16, tokens=6, vocab=100)
            b.update({"neg_image": n["image"], "neg_input_ids": n["input_ids"], "neg_attention_mask": n["attention_mask"]})
        out.append(b)
    return out


def run(loss_module, encoders, cluster, ctx):
    img, txt = copy.deepcopy(encoders)
    model = H.VLInfoStep(txt, img, loss_module, is_amp=True)
    opt = H.make_optimizer(model)
    scaler = torch.amp.GradScaler("cpu", enabled=False)
    losses = []
    with ctx:
        for i, batch in enumerate(batches(cluster)):
            torch.manual_seed(100 + i)                      # the prior terms draw uniform noise (loss.py:186-200)
            out = H.train_step(model, opt, scaler, batch)
            losses.append({k: float(v) for k, v in out["loss_components"].items()})
    return model, losses


@pytest.mark.parametrize("fused_heads", [False, True])
@pytest.mark.parametrize("cluster", [False, True])
def test_train_step_with_dropin_loss_equals_reference_loss(monkeypatch, cluster, fused_heads):
    from clip_lite_b200 import loss as L
    ref = rl.load_reference_loss()
    encoders = tiny_encoders()
    torch.manual_seed(3)
    ref_loss = ref.JSDInfoMaxLoss(image_dim=IMG_DIM, text_dim=TXT_DIM, type="dot", image_prior=True, text_prior=True)
    mine = L.JSDInfoMaxLoss(image_dim=IMG_DIM, text_dim=TXT_DIM, type="dot", image_prior=True, text_prior=True,
                            fused_heads=fused_heads)
    mine.load_state_dict(ref_loss.state_dict(), strict=True)          # the checkpoint contract (checkpointing.py:198-211)

    m_ref, l_ref = run(ref_loss, encoders, cluster, rl.cuda_calls_neutralised())

    import contextlib
    calls = _emu_backend.install(monkeypatch)
    monkeypatch.setattr(L.JSDInfoMaxLoss, "_require_cuda", staticmethod(lambda t: None))
    m_new, l_new = run(mine, encoders, cluster, contextlib.nullcontext())
    assert "jsd_index_fwd_bwd" in calls and (("jsd_ln_normalize_pair" in calls) == fused_heads)

    for a, b in zip(l_ref, l_new):
        assert set(a) == set(b) == {"total_loss", "cross_modal_loss", "visual_loss", "textual_loss"}
        for k in a:
            assert abs(a[k] - b[k]) <= 1e-4 * max(abs(a[k]), 1e-6), (k, a[k], b[k])
    ref_state, new_state = m_ref.state_dict(), m_new.state_dict()
    assert list(ref_state) == list(new_state)                         # same parameter / buffer names, same order
    for k in ref_state:
        a, b = ref_state[k].double(), new_state[k].double()
        assert float((a - b).abs().max()) <= 2e-4 * max(float(a.abs().max()), 1e-6), k


def test_param_groups_follow_the_reference_rule():
    img, txt = tiny_encoders()
    from clip_lite_b200 import loss as L
    model = H.VLInfoStep(txt, img, L.JSDInfoMaxLoss(image_dim=IMG_DIM, text_dim=TXT_DIM))
    opt = H.make_optimizer(model)
    lrs = {name: g["lr"] for (name, _), g in zip(model.named_parameters(), opt.param_groups)}
    assert all(lr == 0.2 for n, lr in lrs.items() if "image_encoder" in n)
    assert all(lr == 1e-3 for n, lr in lrs.items() if "image_encoder" not in n)
    assert lrs["loss.global_d.temperature"] == 1e-3                   # the loss parameters train with OPTIM.LR
