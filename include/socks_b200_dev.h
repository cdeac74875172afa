/*
 * socks_b200_dev.h -- measurement / development exports of libsocksb200.so (tools/, bench.py's live peak probe).
 * Not part of the product contract in socks_b200.h.
 */
#ifndef SOCKS_B200_DEV_H
#define SOCKS_B200_DEV_H

#include "socks_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* register-resident FP64 peak probes: mode 0 DFMA, 1 DMMA, 2 both, 3 exp, 4 DFMA with three register operands, 6 IMMA */
int socks_probe_fp64(int mode, int iters, double* out_dev, int blocks, socks_stream_t stream);
/* clock64() at the phase boundaries of the 128 x 128 leaf factorisation (cycles: 16 int64 on the device). */
int socks_dev_potf2_phases(double* A, int64_t lda, double* dinv, int* info_dev, long long* cycles,
                           socks_stream_t stream);

/* tcgen05.mma kind::i8 check: C (M x N int32) = A (M x K int8) * B (N x K int8)^T, both K-contiguous; M % 128 == 0,
 * N % 64 == 0, K % 32 == 0.  variant swaps the descriptor's leading / stride byte offsets (layout-convention probe). */
int socks_dev_i8gemm_check(const void* A, const void* B, void* C, int M, int N, int K, int variant,
                           socks_stream_t stream);
/* Block-column width of socks_potrf_inv's look-ahead factorisation (see socks_b200.h); returns the previous value. */
int socks_potrf_inv_set_block(int nb);

#ifdef __cplusplus
}
#endif
#endif
