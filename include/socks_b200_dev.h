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
/* FP64-equivalent C (m x n) -= A (m x K) B (n x K)^T on the integer tensor cores (Ozaki scheme: 7 balanced base-256
 * digits per operand, 28 exact int8 GEMMs on tcgen05.mma kind::i8, fp64 recombination); m, n % 128 == 0, K % 32 == 0
 * (longer K than one int32 accumulation chunk is handed over chunk by chunk); syrk bit 0: B = A, bit 1: 8 digits
 * (36 GEMMs, 256x smaller truncation error), bit 2: balance the two operands along K with power-of-two scales first.  The factorisation's trailing updates (7 digits) and the merges of
 * inv(L) (8 digits) run through the same kernels. */
int64_t socks_dev_ozaki_gemm_work_bytes(int64_t m, int64_t n, int64_t K);
int socks_dev_ozaki_gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t m,
                            int64_t n, int64_t K, int syrk, int lower_only, void* work, socks_stream_t stream);
/* Block-column width of socks_potrf_inv's look-ahead factorisation (multiple of 128, 0 = pure recursion); returns the
 * previous value.  Negative codes switch measurement variants: -1/-2 half-width tail off/on, -3/-4 trailing updates on
 * FP64 DMMA / integer tensor cores, -5/-6 large merges likewise, -7/-8 seven / eight digits in the merges. */
int socks_potrf_inv_set_block(int nb);
/* Timeline of one socks_potrf_inv call (CUDA events on both streams): call with NULL pointers to arm it for the next
 * factorisation; afterwards it returns the number of events and fills code = 1000 * kind + step (kinds: 0 start, 1 diagonal
 * block done, 2 panel done, 3 small merges done [panel stream]; 4 column update done, 5 rest update done, 6 join, 7 large
 * merge done [caller's stream]) and the time in ms since the start. */
int socks_dev_potrf_trace(int* codes, double* ms, int cap);

#ifdef __cplusplus
}
#endif
#endif
