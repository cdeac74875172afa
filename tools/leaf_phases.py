"""Cycle breakdown of the 128 x 128 leaf factorisation (socks_dev_potf2_phases)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from socks_b200 import _device as dv, _lib

rng = np.random.default_rng(0)
Z = rng.standard_normal((128, 160))
G = Z @ Z.T / 160 + 0.1 * np.eye(128)
names = ["load", "chol0", "solve0", "upd0", "chol1", "solve1", "upd1", "chol2", "solve2", "upd2", "chol3", "inv+writeL",
         "assemble", "write_dinv"]
for rep in range(3):
    A = dv.to_dev(G)
    dinv = dv.empty(128, 128)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    cyc = torch.zeros(16, dtype=torch.int64, device="cuda")
    _lib.call("socks_dev_potf2_phases", dv.ptr(A), 128, dv.ptr(dinv), info.data_ptr(), cyc.data_ptr(), dv.stream())
    torch.cuda.synchronize()
c = cyc.cpu().numpy()
order = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14]
ticks = [c[i] for i in order if c[i] != 0]
ticks = [c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], c[9], c[10], c[11], c[12], c[13], c[14]]
print("total cycles", ticks[-1] - ticks[0])
for nm, a, b in zip(names, ticks[:-1], ticks[1:]):
    print(f"{nm:12s} {b - a:8d}")
L = np.tril(A.cpu().numpy())
print("residual", np.abs(L @ L.T - G).max(), "inv err", np.abs(np.tril(dinv.cpu().numpy()) @ L - np.eye(128)).max())
