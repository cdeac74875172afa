"""Recipe: make the UNMODIFIED reference importable on the GPU box (test infrastructure, never product code).

The reference is pure Python, so there is nothing to compile: this script copies the `gym_socks`
package verbatim from /root/reference into oracle/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box
with the snapshot exactly like a built .so; it never enters the repository history).  `__graft_entry__.build()` runs it
whenever /root/reference is present.  Consumers: `bench.py --impl reference`, `bench.py`'s `cpu_baseline` leg and
tests/ -- through oracle/ref_loader.py, which also puts the test-only gym shim on the path.

    python oracle/vendor_ref.py [/root/reference]
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def vendor(src_root="/root/reference"):
    src = os.path.join(src_root, "gym_socks")
    if not os.path.isdir(src):
        return None
    dst = os.path.join(DEST, "gym_socks")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(DEST, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "tests"))
    h, count = hashlib.sha256(), 0
    for base, _, files in sorted(os.walk(dst)):
        for f in sorted(files):
            with open(os.path.join(base, f), "rb") as fh:
                h.update(fh.read())
            count += 1
    with open(os.path.join(DEST, "VENDORED_FROM.txt"), "w") as fh:
        fh.write(f"verbatim copy of {src} ({count} files, sha256 {h.hexdigest()[:16]}); made by oracle/vendor_ref.py\n")
    return dst


if __name__ == "__main__":
    out = vendor(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("vendored to", out if out else "(reference tree not found: nothing done)")
