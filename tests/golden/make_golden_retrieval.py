#!/usr/bin/env python
"""Golden vectors for the retrieval scoring path, produced by the reference's OWN ``itm_eval``.

retrieval.py cannot be imported here (it pulls in the training stack: factories, albumentations, ...), so this
script parses /root/reference/retrieval.py with ``ast``, compiles just the ``itm_eval`` function object from the
reference's own source text (nothing is copied into this repository) and runs it on seeded toy retrieval sets.

Run once in the build container:   python tests/golden/make_golden_retrieval.py
Writes retrieval_*.npz: the embeddings, the id maps and the reference's metric dictionary.
"""
from __future__ import annotations

import ast
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.retrieval_oracle import synth_retrieval   # noqa: E402

REFERENCE = os.environ.get("CLIPLITE_REFERENCE_ROOT", "/root/reference")


def reference_itm_eval():
    path = os.path.join(REFERENCE, "retrieval.py")
    tree = ast.parse(open(path).read(), filename=path)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "itm_eval"]
    assert len(fn) == 1, "itm_eval not found in the reference"
    mod = ast.Module(body=fn, type_ignores=[])
    ns = {"np": np, "torch": torch}
    exec(compile(mod, path, "exec"), ns)
    return ns["itm_eval"]


CASES = [  # name, images, captions per image, D, seed, caption noise
    ("retrieval_i40_c5_d64", 40, 5, 64, 0, 6.0),
    ("retrieval_i96_c3_d128", 96, 3, 128, 1, 9.0),
    ("retrieval_i200_c5_d128", 200, 5, 128, 2, 12.0),
]


def main():
    itm_eval = reference_itm_eval()
    for name, n_img, caps, dim, seed, noise in CASES:
        img, txt, txt2img, img2txt, image_ids = synth_retrieval(n_img, caps, dim, seed, noise)
        sims = img @ txt.t()                                        # retrieval.py:143
        res = itm_eval(sims.cpu().numpy(), sims.t().cpu().numpy(), txt2img, img2txt, torch.tensor(image_ids))
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), image_embeds=img.numpy(), text_embeds=txt.numpy(),
            image_ids=np.asarray(image_ids), txt2img=np.asarray([txt2img[j] for j in range(len(txt2img))]),
            img2txt=json.dumps({str(k): v for k, v in img2txt.items()}), metrics=json.dumps(res))
        print(name, res)


if __name__ == "__main__":
    main()
