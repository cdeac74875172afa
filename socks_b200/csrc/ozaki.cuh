// Host entry points of the integer tensor-core GEMM (csrc/ozaki.cu) used by the factorisation (csrc/chol.cu).
#pragma once
#include "common.cuh"

namespace socks {

// bytes of the sliced form (digits + row scales) of a (rows x K) fp64 operand
int64_t ozaki_operand_bytes(int64_t rows, int64_t K, int S);
// S = 7 or 8 digits per entry (one more digit = 256x smaller truncation error, 36 instead of 28 MMAs per K-block).
// operand(r, k) = P[r][k] (trans == 0) or P[k][r] (trans != 0) -> buf (1024-byte aligned): S balanced base-256 digits per
// entry + per-row power-of-two scales.  tri: 0 none, 1 operand(r, k) = 0 for k >= 128 (floor(r/128) + 1), 2 operand(r, k) = 0
// for k < 128 floor(r/128) (those regions of P are never read: they hold garbage in the factorisation's buffers).
int ozaki_slice_operand(const double* P, int64_t rows, int64_t K, int64_t ld, int trans, int tri, uint8_t* buf, int S,
                        const double* kscale, cudaStream_t st);
// Power-of-two scales along K balancing the two operands of one product (A B^T = (A D^-1)(B D)^T): kwork (2 K doubles)
// receives the scale of A in [0, K) and of B in [K, 2K); pass them as `kscale` above (nullptr: no balancing).
int ozaki_kbalance(const double* PA, int64_t rows_a, int64_t lda, int trans_a, int tri_a, const double* PB, int64_t rows_b,
                   int64_t ldb, int trans_b, int tri_b, int64_t K, double* kwork, cudaStream_t st);
// C (m x n, ldc) (op)= A B^T with A = operand rows [128 a_tile0, ..) of abuf, B likewise (m % 128 == 0, n % 128 == 0);
// mode 0: C -= AB^T, 1: C = AB^T, 2: C = -AB^T; tri_a / tri_b clip each tile's K range (tri_a 1: A(i,k) = 0 for k > i,
// tri_a 2: A(i,k) = 0 for k < i; tri_b: B(k,j) = 0 for k < j), at 128-tile granularity.
int ozaki_gemm_nt_sliced(const uint8_t* abuf, int64_t a_rows_total, int64_t a_tile0, const uint8_t* bbuf, int64_t b_rows_total,
                         int64_t b_tile0, int64_t K, double* C, int64_t ldc, int64_t m, int64_t n, int lower_only, int mode,
                         int tri_a, int tri_b, int S, cudaStream_t st);

}  // namespace socks
