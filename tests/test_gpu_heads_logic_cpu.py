"""The GPU parity tests of the projection-head tail (tests/test_zz_gpu_heads.py) have not run on hardware yet.  So
that at least their own logic -- argument tuples, shapes, reference computations, tolerances -- is known to be right,
this test rewrites that file for the CPU (`.cuda()` dropped, device "cpu", the big shapes shrunk, the row-wise entry
points routed to the CPU emulation of the kernel source, the tensor-core entry points to the oracle stand-ins) and
runs every case of it in a sub-process.  A failure of the GPU file on hardware is then a finding about the kernels
or the real tensor-core path, not about the test."""
import os
import subprocess
import sys

import pytest

from tests import _emu_backend

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "test_zz_gpu_heads.py")
GEN = os.path.join(HERE, "_generated_heads_as_cpu.py")

SHRINK = {"(8192, 2048)": "(16, 2048)", "(1000, 2048)": "(10, 2048)", "(300, 1024)": "(9, 1024)",
          "(1024, 2048, 1, False)": "(12, 2048, 1, False)", "(128, 2048, 2, True)": "(8, 2048, 2, True)",
          "(200, 256, 3, True)": "(20, 256, 3, True)", "(512, 2048)": "(16, 2048)", "(1024, 2048)]": "(24, 2048)]",
          "(512, 1024)": "(16, 1024)", "(1000, 256)": "(40, 256)", "(256, 128)": "(32, 128)",
          '("dense", 512, 128)': '("dense", 32, 128)', '("dense", 256, 2048)': '("dense", 16, 2048)',
          '("index", 512, 2048)': '("index", 16, 2048)'}

# CUDA graphs do not exist on the CPU: the rewrite replays the eager step through the same interface
EAGER_GRAPHED_STEP = '''    class GraphedStep:
        def __init__(self, loss_fn, f, g, t):
            self.fn, self.t = loss_fn, t

        def __call__(self, f, g):
            f, g = f.clone().requires_grad_(True), g.clone().requires_grad_(True)
            loss = self.fn(f, g, self.t)[0]
            return (loss,) + tuple(torch.autograd.grad(loss, (f, g, self.t)))
'''

PREAMBLE = '''pytestmark = []
from tests import _emu_backend


@pytest.fixture(autouse=True)
def _emu(monkeypatch):
    from clip_lite_b200 import loss as L
    _emu_backend.install(monkeypatch, bwd_blocks=5)
    monkeypatch.setattr(L.JSDInfoMaxLoss, "_require_cuda", staticmethod(lambda t: None))
'''


@pytest.mark.skipif(not _emu_backend.available(), reason="needs g++ and the CUDA headers")
def test_gpu_head_tests_pass_on_the_cpu_emulation():
    s = open(SRC).read()
    marker = "pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]"
    assert marker in s
    s = s.replace(".cuda()", "").replace('torch.Generator(device="cuda")', "torch.Generator()")
    s = s.replace('device="cuda"', 'device="cpu"').replace(marker, PREAMBLE)
    graph_import = "    from clip_lite_b200.graph import GraphedStep\n"
    assert graph_import in s
    s = s.replace(graph_import, EAGER_GRAPHED_STEP)
    for big, small in SHRINK.items():
        assert big in s, f"{big} no longer appears in test_zz_gpu_heads.py: update SHRINK"
        s = s.replace(big, small)
    _emu_backend.build()
    try:
        with open(GEN, "w") as fh:
            fh.write(s)
        # (the bad-argument case checks the real library's error channel and the "no CPU tensors" rule: not emulated)
        res = subprocess.run([sys.executable, "-m", "pytest", GEN, "-q", "-x", "-p", "no:cacheprovider",
                              "-k", "not test_bad_arguments_fail_loudly"],
                             capture_output=True, text=True, timeout=1500, cwd=os.path.dirname(HERE))
    finally:
        if os.path.exists(GEN):
            os.remove(GEN)
    tail = res.stdout[-1500:]
    assert res.returncode == 0, tail + res.stderr[-1500:]
    assert " passed" in tail and "failed" not in tail and "skipped" not in tail, tail
