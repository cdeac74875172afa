// Host entry points of the integer tensor-core GEMM (csrc/ozaki.cu) used by the factorisation (csrc/chol.cu).
#pragma once
#include "common.cuh"

namespace socks {

// bytes of the sliced form (digits + row scales) of a (rows x K) fp64 operand
int64_t ozaki_operand_bytes(int64_t rows, int64_t K);
// P (rows x K, ld) -> buf (1024-byte aligned): 7 balanced base-256 digits per entry, per-row power-of-two scales
int ozaki_slice_operand(const double* P, int64_t rows, int64_t K, int64_t ld, uint8_t* buf, cudaStream_t st);
// C (m x n, ldc) -= A B^T with A = rows [128 a_tile0, ..) of the operand in abuf, B likewise; m, n multiples of 128 / 64
int ozaki_gemm_nt_sliced(const uint8_t* abuf, int64_t a_rows_total, int64_t a_tile0, const uint8_t* bbuf, int64_t b_rows_total,
                         int64_t b_tile0, int64_t K, double* C, int64_t ldc, int64_t m, int64_t n, int lower_only,
                         cudaStream_t st);

}  // namespace socks
