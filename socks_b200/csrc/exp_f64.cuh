// Table-driven fp64 exp for the fused kernels: exp(x) = 2^e * T[j] * (1 + p(r)),
//   k = rint(x * 64/ln2) = 64 e + j,  r = x - k ln2/64 (Cody-Waite hi/lo),  |r| <= ln2/128,
//   p(r) = r + r^2/2 + ... + r^5/120  (truncation r^6/720 < 2^-54),  T[j] = 2^(j/64) in shared memory.
// 10 FP64-pipe instructions (CUDA's exp: 14 DFMA + ~4 more), max error ~2 ulp (2.2e-16 relative, checked
// offline against 60-digit arithmetic on 2e5 arguments in [-700, 0]); results below 2^-1022 go through a
// two-step scaling so subnormals round like IEEE and x < -746 gives exactly 0, as numpy's exp does.
// The materialising Gram kernels keep CUDA's exp (1 ulp): they are HBM-write bound anyway.
#pragma once
#include "common.cuh"

namespace socks {

constexpr int EXP_TBL = 64;

__device__ __forceinline__ void exp_table_init(double* tbl, int tid, int nthreads) {
    // 2^(j/64), correctly rounded
    static const double kTbl[EXP_TBL] = {
        1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
        1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
        1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
        1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
        1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
        1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
        1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
        1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
        1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
        1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
        1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
        1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
        1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
        1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
        1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
        1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};
    for (int j = tid; j < EXP_TBL; j += nthreads) tbl[j] = kTbl[j];
}

// x <= ~0 expected (RBF arguments); valid for x in (-inf, 700].  NaN in -> NaN out.
__device__ __forceinline__ double exp_tbl(double x, const double* __restrict__ tbl) {
    const double kMagic = 6755399441055744.0;        // 1.5 * 2^52: rounds to nearest integer
    const double kInv = 92.33248261689366;           // 64 / ln 2
    const double kHi = 0x1.62e42fe000000p-7;         // ln2/64, low 24 bits zero: k * kHi is exact
    const double kLo = 0x1.f473de6af278fp-36;        // ln2/64 - kHi
    const double t = fma(x, kInv, kMagic);
    const int k = __double2loint(t);
    const double kd = t - kMagic;
    double r = fma(kd, -kHi, x);
    r = fma(kd, -kLo, r);
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = p * r;
    p = fma(p, r, r);
    const double T = tbl[k & (EXP_TBL - 1)];
    double res = fma(T, p, T);
    int e = k >> 6;
    if (e < -1020) {          // result near or below the normal range (x < ~-707): rare
        if (!(x >= -800.0)) return (x != x) ? x : 0.0;
        e += 1000;
        res = __hiloint2double(__double2hiint(res) + (e << 20), __double2loint(res));
        return res * 0x1p-1000;
    }
    return __hiloint2double(__double2hiint(res) + (e << 20), __double2loint(res));
}

// Four independent exponentials with the stages interleaved by hand (4 dependency chains in flight).
__device__ __forceinline__ void exp_tbl_x4(const double (&x)[4], double (&out)[4], const double* __restrict__ tbl) {
    const double kMagic = 6755399441055744.0, kInv = 92.33248261689366;
    const double kHi = 0x1.62e42fe000000p-7, kLo = 0x1.f473de6af278fp-36;
    double t[4], kd[4], r[4], p[4], T[4];
    int k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = fma(x[i], kInv, kMagic);
#pragma unroll
    for (int i = 0; i < 4; ++i) { k[i] = __double2loint(t[i]); kd[i] = t[i] - kMagic; }
#pragma unroll
    for (int i = 0; i < 4; ++i) T[i] = tbl[k[i] & (EXP_TBL - 1)];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = fma(kd[i], -kHi, x[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = fma(kd[i], -kLo, r[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(r[i], 1.0 / 120.0, 1.0 / 24.0);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(p[i], r[i], 1.0 / 6.0);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(p[i], r[i], 0.5);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = p[i] * r[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(p[i], r[i], r[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double res = fma(T[i], p[i], T[i]);
        int e = k[i] >> 6;
        if (e < -1020) {
            if (!(x[i] >= -800.0)) {
                out[i] = (x[i] != x[i]) ? x[i] : 0.0;
                continue;
            }
            e += 1000;
            res = __hiloint2double(__double2hiint(res) + (e << 20), __double2loint(res)) * 0x1p-1000;
        } else {
            res = __hiloint2double(__double2hiint(res) + (e << 20), __double2loint(res));
        }
        out[i] = res;
    }
}

}  // namespace socks
