"""The C-ABI shared library builds, loads, and exports exactly what include/jsd_b200.h
declares.  No compute calls (there is no GPU in the CPU test tier)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "jsd_b200.h")


@pytest.fixture(scope="module")
def lib():
    from clip_lite_b200 import _lib, build
    build.build_library()
    return _lib.load()


def _lib_mod():
    from clip_lite_b200 import _lib
    return _lib


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jsd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("jsd_index_fwd_bwd", "jsd_dense_fwd", "jsd_dense_bwd_du", "jsd_dense_bwd_dv",
                 "jsd_normalize_cast", "jsd_normalize_bwd", "jsd_last_error", "jsd_abi_version"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in jsd_b200.h but not exported"


def test_python_binding_covers_every_declared_symbol():
    from clip_lite_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_abi_version_and_error_channel(lib):
    assert lib.jsd_abi_version() == _lib_mod().ABI_VERSION
    assert lib.jsd_last_error() is not None
    # argument validation happens before any CUDA call: a null pointer is refused with a message
    rc = lib.jsd_dense_fwd(None, None, 8, 8, 8, 0, None, None, 0, None, None, None, None, None)
    assert rc != 0 and b"null pointer" in lib.jsd_last_error()
    rc = lib.jsd_index_fwd_bwd(None, None, 0, 4, 4, None, None, None, None, None, None, None, None, None, 1.0, None, None)
    assert rc != 0
    assert lib.jsd_index_workspace_bytes(1024) == 1024 * 16
    assert lib.jsd_dense_workspace_bytes() > 0


def test_sass_uses_blackwell_tensor_core_and_tma_paths():
    """cuobjdump evidence that the dense kernels are tcgen05 + TMA, not legacy mma.sync."""
    import shutil
    import subprocess
    from clip_lite_b200 import build
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", build.build_library()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA.16816" not in sass


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from clip_lite_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.JSDLibraryError):
        _lib.load()


def test_peer_ctx_layout_matches_the_header(tmp_path):
    """The ctypes mirror of struct jsd_peer_ctx has the size and field offsets the C compiler gives the header's."""
    import subprocess
    from clip_lite_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "jsd_b200.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %d %d\\n", sizeof(jsd_peer_ctx), '
        'offsetof(jsd_peer_ctx, rank), offsetof(jsd_peer_ctx, world), offsetof(jsd_peer_ctx, rows), '
        'offsetof(jsd_peer_ctx, dim), offsetof(jsd_peer_ctx, v_all), offsetof(jsd_peer_ctx, stage), '
        'offsetof(jsd_peer_ctx, flags), JSD_MAX_PEERS, JSD_PEER_HANDLE_BYTES); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    c = _lib.PeerCtx
    want = [ctypes.sizeof(c), c.rank.offset, c.world.offset, c.rows.offset, c.dim.offset, c.v_all.offset,
            c.stage.offset, c.flags.offset, _lib.MAX_PEERS, _lib.PEER_HANDLE_BYTES]
    assert got == want


def test_header_is_plain_c():
    """include/jsd_b200.h compiles as C (no C++-isms, no CUDA or torch types in the boundary)."""
    import subprocess
    subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", HEADER], check=True)


def test_integration_stub_matches_the_binding():
    """The ctypes stub a maintainer would paste from INTEGRATION.md section 2 declares every entry point it uses
    with exactly the argument types of clip_lite_b200._lib.SIGNATURES (which test_python_binding_covers... ties
    to the header), and its call passes as many arguments as it declares."""
    import ast
    from clip_lite_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "ctypes.CDLL" in b)
    tree = ast.parse(stub)
    declared = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Attribute):
            tgt = node.targets[0]
            if tgt.attr in ("argtypes", "restype") and isinstance(tgt.value, ast.Attribute):
                value = eval(compile(ast.Expression(node.value), "<stub>", "eval"), {"ctypes": ctypes})
                declared.setdefault(tgt.value.attr, {})[tgt.attr] = value
    assert "jsd_index_fwd_bwd" in declared
    for name, d in declared.items():
        res, args = _lib.SIGNATURES[name]
        assert d["restype"] is res, name
        if "argtypes" in d:
            assert list(d["argtypes"]) == list(args), f"{name}: stub argtypes differ from the binding"
    calls = [n for n in ast.walk(tree) if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute)
             and n.func.attr == "jsd_index_fwd_bwd"]
    assert calls and all(len(c.args) == len(_lib.SIGNATURES["jsd_index_fwd_bwd"][1]) for c in calls)


def test_gpu_verified_kernels_are_unchanged():
    """Every kernel of the build that last ran green on a B200 (profiles/sass_manifest_gpu_verified_r02.txt: one md5
    of the disassembly per kernel) is byte-identical in the current build: kernels added since then (the
    projection-head tail, written without a GPU at hand) did not touch the verified machine code.  After a GPU run
    of a deliberate kernel change, regenerate the manifest with tools/sass_diff.py --manifest."""
    import shutil
    import subprocess
    import sys
    from clip_lite_b200 import build
    manifest = os.path.join(ROOT, "profiles", "sass_manifest_gpu_verified_r02.txt")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not (shutil.which("cuobjdump") and os.path.exists(nvcc)):
        pytest.skip("CUDA toolkit not available")
    if "V12.9.86" not in subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout:
        pytest.skip("manifest was written by nvcc 12.9.86")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import sass_diff
    finally:
        sys.path.pop(0)
    changed, added, removed = sass_diff.compare(sass_diff.read_manifest(manifest),
                                                sass_diff.kernels(build.build_library()))
    assert not changed and not removed, (changed, removed)
    assert all("ln_normalize" in k or "ln_bwd_finalize" in k for k in added), added


def test_nvtx_ranges_wrap_library_calls(monkeypatch):
    """JSD_NVTX=1: one NVTX range per entry point, popped even when the call fails."""
    import torch
    from clip_lite_b200 import _lib
    events = []
    monkeypatch.setattr(_lib, "NVTX", True)
    monkeypatch.setattr(torch.cuda.nvtx, "range_push", lambda name: events.append(("push", name)))
    monkeypatch.setattr(torch.cuda.nvtx, "range_pop", lambda: events.append(("pop",)))
    with pytest.raises(_lib.JSDLibraryError):
        _lib.call("jsd_dense_fwd", None, None, 8, 8, 8, 0, None, None, 0, None, None, None, None, None)
    assert events == [("push", "jsd_dense_fwd"), ("pop",)]


def test_head_tail_entry_points_validate_before_launching(lib):
    """Argument checks of the projection-head tail run before any CUDA call (so they can be exercised without a GPU):
    null pointers, incomplete second row set, oversized D, workspace size."""
    fake = 0x1000                                             # never dereferenced: every case fails validation first
    rc = lib.jsd_ln_normalize_pair(None, None, 0, 8, 64, None, None, 1e-5, None, None, 1e-5, 0, None, None, None, None,
                                   None)
    assert rc != 0 and b"null pointer" in lib.jsd_last_error()
    rc = lib.jsd_ln_normalize_pair(fake, fake, 0, 8, 64, None, None, 1e-5, None, None, 1e-5, 0, fake, None, fake, None,
                                   None)
    assert rc != 0 and b"together" in lib.jsd_last_error()
    args = [fake, None, 0, 4, 4104, None, None, None, None, fake, None, fake, None, 1, 0, 0.0, None, 0, None, 0, None,
            None, None, 4, fake, fake, None, None, None, None, None, None, None, None]
    rc = lib.jsd_ln_normalize_bwd_pair(*args)
    assert rc != 0 and b"too large" in lib.jsd_last_error()
    args[13] = 0                                              # n_slices = 0
    rc = lib.jsd_ln_normalize_bwd_pair(*args)
    assert rc != 0 and b"slices" in lib.jsd_last_error()
    args[13], args[20] = 1, fake                              # gdiag without partner rows / temperature
    rc = lib.jsd_ln_normalize_bwd_pair(*args)
    assert rc != 0 and b"positive-pair" in lib.jsd_last_error()
    assert lib.jsd_ln_workspace_bytes(4, 64) == 2 * 4 * 2 * 64 * 4          # never more blocks than rows
    assert lib.jsd_ln_workspace_bytes(0, 64) == 0


def test_product_build_never_enables_the_cpu_emulation():
    """JSD_HOST_EMU (tests/emu: kernel source compiled for the CPU) is a test-tier switch: the product's nvcc command
    does not define it, no product module mentions it, and the shipped library exports no emulation entry point."""
    from clip_lite_b200 import build
    assert not any("JSD_HOST_EMU" in f for f in build.NVCC_FLAGS)
    pkg = os.path.join(ROOT, "clip_lite_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            assert "emu" not in open(os.path.join(pkg, name)).read().lower(), name
    lib = ctypes.CDLL(build.build_library())
    assert not hasattr(lib, "emu_ln_normalize_pair") and not hasattr(lib, "emu_set_bwd_blocks")
