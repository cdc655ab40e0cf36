"""TEST / MEASUREMENT INFRASTRUCTURE ONLY — snapshot the pure-Python files of the reference that the denoising path
imports into oracle/_ref/ (git-ignored, NOT gpurun-ignored), so the GPU box can run the UNMODIFIED reference modules
(`bench.py --impl reference`, the `gpu_reference` field, tests/test_zz_speed_gpu.py) without /root/reference.

    python -m oracle.make_ref_snapshot        # needs /root/reference (the build container)

Nothing is edited: files are copied byte for byte, keeping their paths relative to the reference root
(magicdrive/**.py and third_party/diffusers/src/diffusers/**.py).  oracle/ref_shim.py then loads them from oracle/_ref
when /root/reference is absent.  The snapshot never enters the git history (.gitignore: oracle/_ref/) and nothing in the
product (magicdrive_b200/) imports it.
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(ROOT, "_ref")
SRC = os.environ.get("MAGICDRIVE_REFERENCE_SRC", "/root/reference")
TREES = ["magicdrive", "third_party/diffusers/src/diffusers"]


def snapshot(verbose=True):
    if not os.path.isdir(os.path.join(SRC, TREES[0])):
        if verbose:
            print(f"[ref-snapshot] {SRC} not present: keeping whatever oracle/_ref already holds")
        return os.path.isdir(os.path.join(DST, TREES[0]))
    n, nbytes, h = 0, 0, hashlib.sha256()
    for tree in TREES:
        for dirpath, dirnames, filenames in os.walk(os.path.join(SRC, tree)):
            dirnames[:] = sorted(d for d in dirnames if d != "__pycache__")
            for f in sorted(filenames):
                if not f.endswith(".py"):
                    continue
                src = os.path.join(dirpath, f)
                rel = os.path.relpath(src, SRC)
                dst = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                data = open(src, "rb").read()
                h.update(rel.encode())
                h.update(data)
                n += 1
                nbytes += len(data)
    with open(os.path.join(DST, "SNAPSHOT.txt"), "w") as fh:
        fh.write(f"byte-for-byte copy of {n} .py files ({nbytes} bytes) of {SRC} ({', '.join(TREES)})\nsha256 {h.hexdigest()}\n")
    if verbose:
        print(f"[ref-snapshot] {n} files, {nbytes / 1e6:.1f} MB -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if snapshot() else 1)
