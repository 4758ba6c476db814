"""Kernel-by-kernel comparison of builds of libjsd_b200.so (cuobjdump -sass, one md5 per kernel).

    python tools/sass_diff.py old.so new.so          kernels whose machine code differs / was added / removed
    python tools/sass_diff.py --manifest lib.so      "md5  mangled-name" per kernel (profiles/sass_manifest_*.txt)
    python tools/sass_diff.py --check manifest lib   every kernel of the manifest is byte-identical in lib

Used to show, without a GPU, that a change which adds kernels leaves every GPU-verified kernel untouched
(tests/test_abi.py::test_gpu_verified_kernels_are_unchanged)."""
import hashlib
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    parts = re.split(r"^\s*Function : (\S+)\s*$", out, flags=re.M)
    return {name: hashlib.md5(body.encode()).hexdigest() for name, body in zip(parts[1::2], parts[2::2])}


def read_manifest(path):
    res = {}
    for line in open(path):
        if line.strip() and not line.startswith("#"):
            md5, name = line.split()
            res[name] = md5
    return res


def demangle(names):
    if not names:
        return []
    out = subprocess.run(["c++filt"] + list(names), capture_output=True, text=True).stdout.splitlines()
    return [o[:150] for o in out]


def compare(ka, kb):
    changed = sorted(k for k in ka if k in kb and ka[k] != kb[k])
    added = sorted(k for k in kb if k not in ka)
    removed = sorted(k for k in ka if k not in kb)
    return changed, added, removed


def main(argv):
    if argv[0] == "--manifest":
        for name, md5 in sorted(kernels(argv[1]).items()):
            print(md5, name)
        return 0
    ka = read_manifest(argv[1]) if argv[0] == "--check" else kernels(argv[0])
    kb = kernels(argv[2] if argv[0] == "--check" else argv[1])
    changed, added, removed = compare(ka, kb)
    print(f"{len(ka)} kernels before, {len(kb)} after; identical: {len(ka) - len(changed) - len(removed)}")
    for title, lst in (("CHANGED", changed), ("added", added), ("removed", removed)):
        print(f"{title}: {len(lst)}")
        for k in demangle(lst):
            print("   ", k)
    return 1 if changed or removed else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
