"""TEST INFRASTRUCTURE.  Recipe that makes the UNMODIFIED reference travel to the GPU box.

    python oracle/vendor_ref.py          (build container only; __graft_entry__.build() calls `vendor()`)

The reference (Nandan-Upadhyaya/GAN-codes, Deep_Fusion_GAN/src) is pure Python without a build, so "compiling it
where it lies" is a byte-for-byte copy of its `*.py` files into `oracle/_ref/src/` -- git-ignored (never in history),
not gpurun-ignored (ships with the snapshot like the built .so).  Nothing is edited: `verify()` compares every copied
file with its source by SHA-256 whenever /root/reference is present.  Consumers (oracle/ref_runner.py) put
`oracle/_ref` on sys.path and import `src.deep_fusion_gan.model` etc. exactly as the reference's own train.py does.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/Deep_Fusion_GAN/src"
DST = os.path.join(HERE, "_ref")


def _py_files(root: str):
    for d, _, files in os.walk(root):
        if "__pycache__" in d:
            continue
        for f in sorted(files):
            if f.endswith(".py"):
                yield os.path.relpath(os.path.join(d, f), root)


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def vendor() -> bool:
    """Copy the reference's Python sources to oracle/_ref/src.  Returns False (and leaves an existing copy alone)
    when /root/reference is not present (GPU box)."""
    if not os.path.isdir(REF_SRC):
        return False
    dst_src = os.path.join(DST, "src")
    if os.path.isdir(dst_src):
        shutil.rmtree(dst_src)
    n = 0
    for rel in _py_files(REF_SRC):
        out = os.path.join(dst_src, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(REF_SRC, rel), out)
        n += 1
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as f:
        f.write(f"# byte-for-byte copy of {REF_SRC} (*.py), made by oracle/vendor_ref.py; do not edit\n")
        for rel in _py_files(dst_src):
            f.write(f"{_sha(os.path.join(dst_src, rel))}  src/{rel}\n")
    return n > 0


def verify() -> bool:
    """Every file listed in the manifest still has the recorded hash (and equals the source when it is present)."""
    man = os.path.join(DST, "MANIFEST.txt")
    if not os.path.exists(man):
        return False
    for line in open(man):
        if line.startswith("#") or not line.strip():
            continue
        sha, rel = line.split()
        p = os.path.join(DST, rel)
        if not os.path.exists(p) or _sha(p) != sha:
            return False
        src = os.path.join(os.path.dirname(REF_SRC), rel)
        if os.path.exists(src) and _sha(src) != sha:
            return False
    return True


def available() -> bool:
    return os.path.isdir(os.path.join(DST, "src", "deep_fusion_gan")) and verify()


if __name__ == "__main__":
    ok = vendor()
    print("vendored" if ok else "reference not present", "->", DST, "| verified:", verify())
    sys.exit(0 if ok else 1)
