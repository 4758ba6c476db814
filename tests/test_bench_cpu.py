"""Host logic of bench.py and of the path-selection policies (no GPU): both arms print the same `config`, the CPU
arm runs on a tiny workload with every key the contract names, and the fused-kernel policy reads as documented."""
import argparse
import importlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    sys.path.insert(0, ROOT)
    return importlib.import_module("bench")


def _args(**kw):
    base = dict(gpus=1, steps=2, warmup=1, impl="reference", workload="dense_b8192_d1024", no_cpu_baseline=False,
                no_parity=False, cuda_graph=1, route="reduce", exchange="peer")
    base.update(kw)
    return argparse.Namespace(**base)


def test_config_is_a_function_of_workload_and_gpus_only(bench):
    for name, wl in bench.WORKLOADS.items():
        for gpus in (1, 2, 8):
            a = bench.workload_config(_args(workload=name, gpus=gpus, impl="b200"), wl)
            b = bench.workload_config(_args(workload=name, gpus=gpus, impl="reference", steps=7, warmup=3), wl)
            assert a == b and a["workload"] == name and a["n_gpus"] == gpus
            assert a["global_batch"] == wl["batch"] * (gpus if wl["weak"] else 1)
            json.dumps(a)


@pytest.mark.parametrize("workload", ["dense_tiny", "index_tiny"])
def test_reference_arm_prints_the_contract_keys(bench, monkeypatch, capsys, workload):
    monkeypatch.setitem(bench.WORKLOADS, "dense_tiny", dict(batch=256, dim=64, weak=False, mode="dense", dtype="bf16"))
    monkeypatch.setitem(bench.WORKLOADS, "index_tiny", dict(batch=256, dim=64, weak=True, mode="index", dtype="f32"))
    monkeypatch.setattr(bench, "time_cpu", lambda wl, b, d, budget_s, steps=None, warmup=1:
                        bench.__dict__["_orig_time_cpu"](wl, b, d, 0.2, steps=1, warmup=1))
    args = _args(workload=workload, steps=1, warmup=1)
    bench.run_reference(args, bench.WORKLOADS[workload])
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["config"] == bench.workload_config(args, bench.WORKLOADS[workload])
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["value"] == line["value"] and cb["cores"] >= 1 and cb["kind"] in ("port", "reference") and cb["sample"]
    if workload == "index_tiny":
        from oracle import reference_loader as rl
        assert cb["kind"] == ("reference" if rl.reference_available() else "port")


def test_fused_kernel_policy(monkeypatch):
    from clip_lite_b200 import _lib, kernels
    lib = _lib.load()
    for d in (64, 128, 192, 256):
        assert lib.jsd_dense_fused_supported(1024, d) == 1
    for b, d in ((1024, 1024), (1024, 96), (1024, 320), (1, 128), (1 << 17, 128)):
        assert lib.jsd_dense_fused_supported(b, d) == 0
    for b in (2, 128, 1024, 4096, 65536):
        ns = lib.jsd_dense_fused_splits(b, 128)
        assert 1 <= ns <= 8 and ns <= (b + 127) // 128
    monkeypatch.delenv("JSD_FUSED", raising=False)
    assert kernels.fused_supported(1024, 128) and kernels.fused_supported(4096, 256)
    assert not kernels.fused_supported(8192, 128)            # default: staged above the measured crossover
    assert not kernels.fused_supported(1024, 1024)
    monkeypatch.setenv("JSD_FUSED", "1")
    assert kernels.fused_supported(65536, 128) and not kernels.fused_supported(1024, 1024)
    monkeypatch.setenv("JSD_FUSED", "0")
    assert not kernels.fused_supported(1024, 128)


def test_loss_module_validates_the_new_options():
    from clip_lite_b200.loss import JSDInfoMaxLoss
    with pytest.raises(ValueError):
        JSDInfoMaxLoss(8, 8, neg_mode="dense", gather=True, exchange="peer", grad_partials="fp16")
    m = JSDInfoMaxLoss(8, 8, neg_mode="dense", gather=True, exchange="peer", grad_partials="fp32")
    assert m.grad_partials == "fp32" and m.exchange == "peer"
