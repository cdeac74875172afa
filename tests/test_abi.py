"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol that
include/socks_b200.h declares, and argument validation fails with the documented codes before any
CUDA call is made.  No compute here."""
import os
import re

import pytest

from conftest import ROOT
from socks_b200 import _lib

HEADERS = [os.path.join(ROOT, "include", "socks_b200.h"), os.path.join(ROOT, "include", "socks_b200_dev.h")]


def _declared(headers=HEADERS):
    names = set()
    for h in headers:
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(socks_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in socks_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in socks_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)
    # the product header carries no probe / tuning exports
    product = _declared(HEADERS[:1])
    assert not [n for n in product if "probe" in n or "_dev_" in n or n == "socks_potrf_inv_set_block"]
    for composite in ("socks_sr_fit", "socks_srmax_fit", "socks_srmax_predict", "socks_fwd_scores", "socks_median_sq_dist",
                      "socks_ctx_create", "socks_ctx_destroy", "socks_comm_init", "socks_bcast", "socks_allgather"):
        assert composite in product


def test_version_and_padding():
    lib = _lib.load()
    assert lib.socks_version() >= 100
    assert lib.socks_padded_dim(1) == 128
    assert lib.socks_padded_dim(128) == 128
    assert lib.socks_padded_dim(129) == 256
    assert lib.socks_padded_dim(20000) == 20096
    assert lib.socks_potrf_work_bytes(1000) == 1024 * 128 * 8
    assert lib.socks_trtri_work_bytes(100) == 16
    assert lib.socks_trtri_work_bytes(1000) >= 512 * 512 * 8  # merge scratch + the digit form of one block-column panel
    assert lib.socks_sr_predict_work_bytes(20000, 250000) >= (20000 * 20 + 10 * 16 * 250000) * 8


def test_argument_validation_codes():
    lib = _lib.load()
    # d out of range -> argument 5 of socks_gram_rbf
    assert lib.socks_gram_rbf(None, 4, None, 4, 0, -1.0, 0, None, None, 4, None) == -5
    assert b"invalid argument" in lib.socks_last_error()
    # ldk < m -> argument 10
    assert lib.socks_gram_rbf(None, 4, None, 4, 2, -1.0, 0, None, None, 2, None) == -10
    # bad problem code -> argument 12 of socks_sr_predict
    assert lib.socks_sr_predict(1, 1, 1, 8, 8, 2, -1.0, 0, 1, 8, 1, 7, 3, 1, 8, 16, 0, None) == -12
    # non-multiple-of-128 m for the GEMM -> argument 7
    assert lib.socks_gemm_tn(16, 128, 16, 128, 16, 128, 100, 128, 128, None) == -7
    with pytest.raises(_lib.SocksError):
        _lib.call("socks_potrf", None, 10, 128, None, None, None)


def test_no_cpu_fallback_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from socks_b200.kernel.metrics import rbf_kernel
    import numpy as np

    with pytest.raises(_lib.SocksError):
        rbf_kernel(np.zeros((3, 2)), sigma=1.0)
