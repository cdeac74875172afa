import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")
    # the CUDA library is a build artefact (git-ignored).  `make` is incremental: it rebuilds when any csrc/*.cu,
    # *.cuh or the public header is newer than the .so, so a stale binary is never tested; a failed build fails loudly.
    import shutil
    import subprocess

    csrc = os.path.join(ROOT, "socks_b200", "csrc")
    lib = os.path.join(ROOT, "socks_b200", "libsocksb200.so")
    if shutil.which("nvcc") and shutil.which("make"):
        r = subprocess.run(["make", "-C", csrc, "-j", "8"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise pytest.UsageError("building socks_b200/libsocksb200.so failed:\n" + r.stdout[-4000:])
    elif not os.path.exists(lib):
        raise pytest.UsageError("socks_b200/libsocksb200.so is missing and nvcc/make are not available to build it")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def tubes_from(arr):
    """(T, 2, d) float64 array of float32-representable bounds -> list of Box."""
    from socks_b200.spaces import Box

    return [Box(low=a[0].astype(np.float32), high=a[1].astype(np.float32), shape=a[0].shape) for a in arr]


@pytest.fixture(scope="session")
def golden():
    return load_golden
