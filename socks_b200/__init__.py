"""socks_b200 — B200-native (sm_100a) implementation of SOCKS's kernel-embedding hot path.

Drop-in for `gym_socks.kernel.metrics.{rbf_kernel, regularized_inverse}` and the
`KernelSR` / `KernelMaximalSR` / `KernelControlFwd` algorithms.  All arithmetic
runs in hand-written fp64 CUDA kernels behind the C ABI declared in `include/socks_b200.h`
(`socks_b200/libsocksb200.so`); there is no CPU fallback: the first call that needs the
library raises if it cannot be loaded or no B200 is visible.
"""

__version__ = "0.1.0"
