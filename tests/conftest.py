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
    # the CUDA library is a build artefact (git-ignored): build it once if this is a fresh checkout
    lib = os.path.join(ROOT, "socks_b200", "libsocksb200.so")
    if not os.path.exists(lib):
        import subprocess

        subprocess.run(["make", "-C", os.path.join(ROOT, "socks_b200", "csrc"), "-j", "8"], check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def tubes_from(arr):
    """(T, 2, d) float64 array of float32-representable bounds -> list of Box."""
    from socks_b200.spaces import Box

    return [Box(low=a[0].astype(np.float32), high=a[1].astype(np.float32), shape=a[0].shape) for a in arr]


@pytest.fixture(scope="session")
def golden():
    return load_golden
