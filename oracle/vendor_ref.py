"""Recipe: make the UNMODIFIED reference importable on the GPU box (test infrastructure, never product code).

The reference is pure Python, so there is nothing to compile: this script packs the `gym_socks` package
verbatim from /root/reference into ONE archive, oracle/_ref/gym_socks_ref.zip (git-ignored, NOT gpurun-ignored: it travels
to the GPU box with the snapshot exactly like a built .so; it never enters the repository history and no reference source
file is unpacked into the tree -- Python imports straight from the archive).  `__graft_entry__.build()` runs it whenever
/root/reference is present.  Consumers: `bench.py --impl reference`, `bench.py`'s `cpu_baseline` leg and tests/ -- through
oracle/ref_loader.py, which also puts the test-only gym shim on the path.

    python oracle/vendor_ref.py [/root/reference]
"""
import hashlib
import os
import shutil
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(DEST, "gym_socks_ref.zip")


def vendor(src_root="/root/reference"):
    src = os.path.join(src_root, "gym_socks")
    if not os.path.isdir(src):
        return None
    os.makedirs(DEST, exist_ok=True)
    stale = os.path.join(DEST, "gym_socks")
    if os.path.isdir(stale):  # an unpacked copy from an earlier recipe
        shutil.rmtree(stale)
    h, count = hashlib.sha256(), 0
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        for base, dirs, files in sorted(os.walk(src)):
            dirs[:] = sorted(d for d in dirs if d not in ("__pycache__", "tests"))
            for f in sorted(files):
                if not f.endswith(".py"):
                    continue
                full = os.path.join(base, f)
                with open(full, "rb") as fh:
                    data = fh.read()
                h.update(data)
                z.writestr(os.path.relpath(full, src_root), data)
                count += 1
    with open(os.path.join(DEST, "VENDORED_FROM.txt"), "w") as fh:
        fh.write(f"verbatim archive of {src} ({count} files, sha256 {h.hexdigest()[:16]}); made by oracle/vendor_ref.py\n")
    return ARCHIVE


if __name__ == "__main__":
    out = vendor(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("vendored to", out if out else "(reference tree not found: nothing done)")
