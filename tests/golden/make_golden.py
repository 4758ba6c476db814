#!/usr/bin/env python
"""Generate the golden vectors in this directory from the UNMODIFIED reference
(/root/reference/loss.py, imported through oracle/reference_loader.py).

Run once in the build container:   python tests/golden/make_golden.py
The reference tree does not exist on the GPU box, so the tests only ever read
the committed .npz / .json files written here.

Cases
-----
estimator_*.npz   reference JSDInfoMaxLoss with nn.Identity projection heads, i.e.
                  loss.py:94-105,204-254 on raw (B, D) embeddings; normal mode,
                  cluster mode and the two self-supervised call sites.  The
                  reference runs in float64 on float32 inputs; stored are the four
                  loss components and d(total_loss)/d(inputs, temperature).
module_*.npz      the full reference module (projection heads, BatchNorm in train
                  mode, priors off) with weights from tests/golden/_weights.py;
                  stored are losses, input grads, per-parameter grad checksums
                  and the BatchNorm buffers after the step.
state_dict_keys.json  names/shapes of the default-config module's state_dict.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import reference_loader as rl          # noqa: E402
from oracle.jsd_oracle import synth_embeddings      # noqa: E402
from _weights import seeded_state_dict              # noqa: E402


def _np(x):
    return x.detach().cpu().numpy()


def estimator_case(name, b, d, seed, correlated, t, mode="normal", ssl=False):
    f, g = synth_embeddings(b, d, seed, correlated)
    kw = {}
    m = rl.reference_estimator_module(visual_self_supervised=ssl, textual_self_supervised=ssl)
    if ssl:
        for blk in (m.visual_d, m.textual_d):
            blk.img_block = torch.nn.Identity()
            blk.text_block = torch.nn.Identity()
    m = m.double()
    m.global_d.temperature.data.fill_(t)
    if ssl:
        m.visual_d.temperature.data.fill_(t - 0.3)
        m.textual_d.temperature.data.fill_(t - 0.3)
    inputs = {"image_features": f, "text_features": g}
    if mode == "cluster":
        nf, ng = synth_embeddings(b, d, seed + 100, correlated)
        inputs["neg_image_features"], inputs["neg_text_features"] = nf, ng
    if ssl:
        af, ag = synth_embeddings(b, d, seed + 200, correlated)
        inputs["aug_image_features"] = 0.7 * f + 0.5 * af
        inputs["aug_text_features"] = 0.7 * g + 0.5 * ag
    leaves = {k: v.double().requires_grad_(True) for k, v in inputs.items()}
    with rl.cuda_calls_neutralised():
        out = m(**leaves)
    out["total_loss"].backward()
    rec = {"t": np.float64(t), "mode": mode, "ssl": ssl}
    for k, v in inputs.items():
        rec["in_" + k] = _np(v)
        rec["grad_" + k] = _np(leaves[k].grad)
    for k, v in out.items():
        rec["out_" + k] = _np(v)
    rec["grad_temperature"] = _np(m.global_d.temperature.grad)
    if ssl:
        rec["t_ssl"] = np.float64(t - 0.3)
        rec["grad_temperature_visual"] = _np(m.visual_d.temperature.grad)
        rec["grad_temperature_textual"] = _np(m.textual_d.temperature.grad)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(f"{name}: cross={float(out['cross_modal_loss']):.9f} total={float(out['total_loss']):.9f}")


def module_case(name, b, image_dim, text_dim, seed, mode="normal", train=True):
    ref = rl.load_reference_loss()
    m = ref.JSDInfoMaxLoss(image_dim=image_dim, text_dim=text_dim, type="dot",
                           image_prior=False, text_prior=False)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(seeded_state_dict(shapes, seed))
    m = m.double()
    m.train(train)
    gen = torch.Generator("cpu").manual_seed(seed)
    img = torch.randn(b, image_dim, generator=gen)
    txt = 0.5 * torch.randn(b, text_dim, generator=gen)
    txt[:, : min(image_dim, text_dim)] += 0.5 * img[:, : min(image_dim, text_dim)]
    inputs = {"image_features": img, "text_features": txt}
    if mode == "cluster":
        inputs["neg_image_features"] = torch.randn(b, image_dim, generator=gen)
        inputs["neg_text_features"] = torch.randn(b, text_dim, generator=gen)
    leaves = {k: v.double().requires_grad_(True) for k, v in inputs.items()}
    with rl.cuda_calls_neutralised():
        out = m(**leaves)
    out["total_loss"].backward()
    rec = {"seed": seed, "image_dim": image_dim, "text_dim": text_dim, "mode": mode, "train": train}
    for k, v in inputs.items():
        rec["in_" + k] = _np(v)
        rec["grad_" + k] = _np(leaves[k].grad)
    for k, v in out.items():
        rec["out_" + k] = _np(v)
    for k, p in m.named_parameters():
        gr = p.grad if p.grad is not None else torch.zeros_like(p)
        rec["pgrad_sum/" + k] = _np(gr.sum())
        rec["pgrad_abs/" + k] = _np(gr.abs().sum())
        # a fixed pseudo-random projection catches permutation/transposition errors
        w = torch.from_numpy(np.random.RandomState(7).standard_normal(tuple(gr.shape) or (1,))).reshape(gr.shape)
        rec["pgrad_proj/" + k] = _np((gr * w).sum())
    for k, v in m.state_dict().items():
        if "running_" in k or "num_batches" in k:
            rec["buf/" + k] = _np(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(f"{name}: cross={float(out['cross_modal_loss']):.9f}")


def state_dict_keys():
    ref = rl.load_reference_loss()
    m = ref.JSDInfoMaxLoss(image_dim=2048, text_dim=768, type="dot", image_prior=True, text_prior=True)
    keys = {k: list(v.shape) for k, v in m.state_dict().items()}
    m2 = ref.JSDInfoMaxLoss(image_dim=32, text_dim=24, type="dot", image_prior=False, text_prior=False,
                            visual_self_supervised=True, textual_self_supervised=True)
    keys_ssl = {k: list(v.shape) for k, v in m2.state_dict().items()}
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as fh:
        json.dump({"default_2048_768_priors": keys, "ssl_32_24_nopriors": keys_ssl,
                   "n_params_default": int(sum(p.numel() for p in m.parameters()))}, fh, indent=1)
    print("state_dict_keys.json:", len(keys), "entries")


def main():
    torch.manual_seed(0)
    t0 = float(np.log(1 / 0.07))
    estimator_case("estimator_b2_d8", 2, 8, 0, True, t0)
    estimator_case("estimator_b3_d8", 3, 8, 1, False, t0)
    estimator_case("estimator_b16_d8", 16, 8, 2, True, 1.5)
    estimator_case("estimator_b16_d128", 16, 128, 0, False, t0)
    estimator_case("estimator_b128_d128", 128, 128, 1, True, t0)
    estimator_case("estimator_b64_d100_hot", 64, 100, 3, True, 3.4)   # tau ~ 30: exercises softplus threshold
    estimator_case("estimator_cluster_b8_d16", 8, 16, 0, True, t0, mode="cluster")
    estimator_case("estimator_cluster_b32_d64", 32, 64, 1, False, t0, mode="cluster")
    estimator_case("estimator_ssl_b16_d32", 16, 32, 2, True, t0, ssl=True)
    estimator_case("estimator_ssl_cluster_b8_d32", 8, 32, 4, True, t0, mode="cluster", ssl=True)
    module_case("module_b8_train", 8, 24, 16, 11)
    module_case("module_b8_eval", 8, 24, 16, 12, train=False)
    module_case("module_cluster_b6_train", 6, 24, 16, 13, mode="cluster")
    state_dict_keys()


if __name__ == "__main__":
    main()
