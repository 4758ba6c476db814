#!/usr/bin/env python
"""End-to-end train-step harness around the loss (BASELINE.json configs[0] / configs[4]).

The encoders, the data pipeline and the trainer of the reference are OUT of scope and its
training stack cannot be imported here (nltk, fvcore, albumentations, sentence_transformers ... are not installed).
What the two e2e configs need from them is small, so this file RESTATES it around an arbitrary loss module:

  VLInfoStep.forward       model.py:32-113, mode "train_sbert": image encoder, text encoder, keyword call of the loss,
                           {"loss", "loss_components"} out, everything under amp.autocast
  ImageEncoder             encoder.py:14-64: torchvision resnet50(zero_init_residual=False), fc = Identity -> [B, 2048]
  TextEncoder              encoder.py:125-197, mode "train_sbert", random init: BertModel(BertConfig(num_hidden_layers)),
                           pooler_output -> [B, 768]
  param groups             factories.py:464-487: lr by substring of the parameter name (CNN 0.2, transformer 1e-3, the
                           rest -- the loss module -- 1e-3), weight decay 1e-4, SGD momentum 0.9 (Lookahead, a wrapper
                           class local to the reference, is left out)
  train_step               train.py:210-227: zero_grad, autocast forward, GradScaler backward, unscale, clip_grad_norm 10,
                           step, update

    python tools/e2e_harness.py --loss reference --device cpu --batch 32          C1: the reference's own loss.py
                                                                                  (needs /root/reference: build container)
    python tools/e2e_harness.py --loss b200 --device cuda --batch 1024            C5 per GPU: the drop-in loss swapped in
    torchrun --nproc-per-node 8 ... tools/e2e_harness.py --loss b200 --batch 128  C5: bs 1024 over 8 GPUs under DDP

Prints one JSON line: images/s of the whole train step (synthetic 224-px images U[0,1), random token ids of length 30).
The parity of the swap (same model, same batch, reference loss vs drop-in loss -> same parameters after the step) is
tested on the CPU in tests/test_e2e_harness_cpu.py.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import sys
import time

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class ImageEncoder(nn.Module):
    def __init__(self, net: nn.Module = None):
        super().__init__()
        if net is None:
            import torchvision
            net = torchvision.models.resnet50(weights=None, zero_init_residual=False)
            net.fc = nn.Identity()
        self.img_encoder = net

    def forward(self, image):
        x = self.img_encoder(image)
        return x.view(x.size(0), x.size(1))


class TextEncoder(nn.Module):
    def __init__(self, num_hidden_layers: int = 12, config=None):
        super().__init__()
        from transformers import BertConfig, BertModel
        self.strans = BertModel(config if config is not None else BertConfig(num_hidden_layers=num_hidden_layers))

    def forward(self, x):
        return self.strans(**x).pooler_output


class VLInfoStep(nn.Module):
    """model.py:15-113 for mode "train_sbert" (negatives / augmentations when the batch carries them)."""

    def __init__(self, text_encoder, image_encoder, loss, is_amp: bool = True):
        super().__init__()
        self.text_encoder, self.image_encoder, self.loss, self.is_amp = text_encoder, image_encoder, loss, is_amp

    def _text(self, batch, prefix):
        return self.text_encoder({"input_ids": batch[prefix + "input_ids"],
                                  "attention_mask": batch[prefix + "attention_mask"]})

    def forward(self, batch):
        device_type = batch["image"].device.type
        with torch.autocast(device_type, enabled=self.is_amp and device_type == "cuda"):
            kw = dict(image_features=self.image_encoder(batch["image"]), text_features=self._text(batch, ""),
                      neg_image_features=None, neg_text_features=None, aug_image_features=None,
                      aug_text_features=None)
            if "neg_input_ids" in batch:
                kw["neg_image_features"] = self.image_encoder(batch["neg_image"])
                kw["neg_text_features"] = self._text(batch, "neg_")
            if "aug_image" in batch:
                kw["aug_image_features"] = self.image_encoder(batch["aug_image"])
            if "aug_input_ids" in batch:
                kw["aug_text_features"] = self._text(batch, "aug_")
            loss_dict = self.loss(**kw)
            return {"loss": loss_dict["total_loss"],
                    "loss_components": {k: v.clone().detach() for k, v in loss_dict.items()}}


NO_DECAY = ".*textual.(embedding|transformer).*(norm.*|bias)"      # config.py:172


def make_optimizer(model: nn.Module, cnn_lr=0.2, trans_lr=1e-3, lr=1e-3, weight_decay=1e-4, momentum=0.9):
    groups = []
    for name, p in model.named_parameters():
        wd = 0.0 if re.match(NO_DECAY, name) else weight_decay
        group_lr = cnn_lr if "image_encoder" in name else (trans_lr if "text_encoder" in name else lr)
        groups.append({"params": [p], "lr": group_lr, "weight_decay": wd})
    return torch.optim.SGD(groups, momentum=momentum)


def train_step(model, optimizer, scaler, batch, clip=10.0):
    optimizer.zero_grad()
    out = model(batch)
    loss = out["loss"]
    scaler.scale(loss).backward()
    scaler.unscale_(optimizer)
    torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
    scaler.step(optimizer)
    scaler.update()
    return out


def synthetic_batch(batch: int, device, seed: int = 0, image_px: int = 224, tokens: int = 30, vocab: int = 30522):
    gen = torch.Generator("cpu").manual_seed(seed)
    return {"image": torch.rand(batch, 3, image_px, image_px, generator=gen).to(device),
            "input_ids": torch.randint(0, vocab, (batch, tokens), generator=gen).to(device),
            "attention_mask": torch.ones(batch, tokens, dtype=torch.long).to(device)}


REFERENCE_ROOT = os.environ.get("CLIPLITE_REFERENCE_ROOT", "/root/reference")


def load_reference_loss_module():
    """The UNMODIFIED reference loss.py, imported from where it lies (build container only; loss.py:9 does
    `from utils import *`, which must resolve to the reference's own empty utils package)."""
    import importlib.util
    path = os.path.join(REFERENCE_ROOT, "loss.py")
    if not os.path.isfile(path):
        raise FileNotFoundError(f"{path} not found: --loss reference runs where the reference tree is present")
    had_utils = sys.modules.pop("utils", None)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        spec = importlib.util.spec_from_file_location("cliplite_reference_loss_e2e", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(REFERENCE_ROOT)
        sys.modules.pop("utils", None)
        if had_utils is not None:
            sys.modules["utils"] = had_utils
    return mod


class cuda_calls_as_identity:
    """loss.py:186,257,280 hard-code `.cuda()`: on a host without a GPU make Tensor.cuda the identity while the
    reference runs (its files are never edited)."""

    def __enter__(self):
        self.orig = torch.Tensor.cuda
        if not torch.cuda.is_available():
            torch.Tensor.cuda = lambda t, *a, **k: t

    def __exit__(self, *a):
        torch.Tensor.cuda = self.orig


def make_loss(kind: str, **kw):
    """kind "reference": the unmodified reference loss.py (build container only); "b200": the drop-in."""
    if kind == "reference":
        return load_reference_loss_module().JSDInfoMaxLoss(**kw)
    from clip_lite_b200.loss import JSDInfoMaxLoss
    return JSDInfoMaxLoss(**kw)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--loss", choices=("reference", "b200"), default="b200")
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--batch", type=int, default=32, help="per process")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--text-layers", type=int, default=12)
    ap.add_argument("--neg-mode", default="shift1")
    ap.add_argument("--fused-heads", action="store_true")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl" if args.device == "cuda" else "gloo")
        if args.device == "cuda":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    device = torch.device(args.device, torch.cuda.current_device()) if args.device == "cuda" else torch.device("cpu")
    import contextlib
    ctx = cuda_calls_as_identity() if args.loss == "reference" else contextlib.nullcontext()
    torch.manual_seed(0)
    extra = {} if args.loss == "reference" else {"neg_mode": args.neg_mode, "fused_heads": args.fused_heads}
    loss = make_loss(args.loss, image_dim=2048, text_dim=768, type="dot", image_prior=True, text_prior=True, **extra)
    model = VLInfoStep(TextEncoder(args.text_layers), ImageEncoder(), loss).to(device)
    if world > 1:
        model = nn.parallel.DistributedDataParallel(
            model, device_ids=[device.index] if device.type == "cuda" else None, find_unused_parameters=True)
    optimizer = make_optimizer(model)
    scaler = torch.amp.GradScaler(device.type, enabled=device.type == "cuda")
    batch = synthetic_batch(args.batch, device, seed=rank)

    def sync():
        if device.type == "cuda":
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    with ctx:
        for _ in range(args.warmup):
            out = train_step(model, optimizer, scaler, batch)
        sync()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = train_step(model, optimizer, scaler, batch)
        sync()
        dt = (time.perf_counter() - t0) / args.steps
    if rank == 0:
        print(json.dumps({"metric": "e2e train step images/s (ResNet-50 + BERT, synthetic 224px / 30 tokens)",
                          "value": world * args.batch / dt, "unit": "images/s", "s_per_step": dt, "impl": args.loss,
                          "device": args.device, "n_procs": world, "batch_per_proc": args.batch,
                          "text_layers": args.text_layers, "threads": torch.get_num_threads(),
                          "total_loss": float(out["loss"].detach()),
                          "config": {"neg_mode": args.neg_mode if args.loss == "b200" else "reference",
                                     "fused_heads": bool(args.fused_heads)}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
