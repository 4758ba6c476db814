"""Deterministic, torch-version-independent weights for the full-module golden
cases: every state_dict entry is filled from numpy's legacy RandomState stream
so that make_golden.py (reference module, build container) and the tests (the
B200 module, GPU box) load bit-identical parameters without storing them."""
from __future__ import annotations

import numpy as np
import torch


def seeded_state_dict(shapes, seed: int):
    """shapes: ordered mapping name -> shape (as in module.state_dict())."""
    rs = np.random.RandomState(seed)
    out = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[name] = torch.tensor(3, dtype=torch.long)
            continue
        if len(shape) == 0:                       # temperature
            out[name] = torch.tensor(float(np.log(1 / 0.07) + 0.1 * rs.standard_normal()))
            continue
        x = rs.standard_normal(shape)
        if leaf == "weight" and len(shape) == 2:
            x = x / np.sqrt(shape[1])
        elif leaf == "weight":                    # BatchNorm / LayerNorm scale
            x = 1.0 + 0.1 * x
        elif leaf == "running_var":
            x = 1.0 + 0.1 * np.abs(x)
        else:                                     # biases, running_mean
            x = 0.1 * x
        out[name] = torch.from_numpy(x.astype(np.float32))
    return out
