// Table-driven fp64 exp for the fused predict kernel: exp(kmul * d2) = 2^e * T[j] * (1 + p(r)) for a
// launch-constant kmul < 0 folded into every constant (no multiply for the argument):
//   k = rint(d2 * kmul * 2048/ln2) = 2048 e + j,   r' = d2 - k * ln2/(2048 kmul)   (ONE fused multiply-add: the
//   constant is the correctly rounded double, so r' carries a relative argument error of 2^-53 |x|),
//   exp(kmul r') - 1 = r'(c1 + r'(c2 + r' c3)),  c_i = kmul^i / i!   (|kmul r'| <= ln2/4096: r^4/24 < 2^-54),
//   T[j] = 2^(j/2048) generated on the fly into 16 KB of shared memory.
// 7 FP64-pipe instructions per value (CUDA's exp: 14 DFMA + ~4 more).  Error: ~2 ulp + |x| 2^-53 relative, i.e.
// <= 2.5e-16 for |x| <= 1 and <= 6e-14 at |x| = 500 where the kernel value is 1e-217 and cannot influence the
// sums (checked offline against 60-digit arithmetic); results below 2^-1022 go through a two-step scaling so
// subnormals round like IEEE and x < -746 gives exactly 0, as numpy's exp does.
// The materialising Gram kernels keep CUDA's exp (1 ulp): rbf_kernel parity is judged at 1e-13.
#pragma once
#include <string.h>

#include "common.cuh"

namespace socks {

constexpr int EXP_TBL2 = 2048;

struct ExpScaled {
    double kinv;  // kmul * 2048 / ln2
    double c;     // ln2 / (2048 kmul), correctly rounded
    double c1, c2, c3;
    double kmul;
};

inline ExpScaled make_exp_scaled(double kmul) {
    ExpScaled p;
    const long double ln2 = 0.693147180559945309417232121458176568L;
    const long double k = kmul;
    p.kmul = kmul;
    p.kinv = (double)(k * 2048.0L / ln2);
    p.c = (double)(ln2 / (2048.0L * k));
    p.c1 = (double)k;
    p.c2 = (double)(k * k / 2.0L);
    p.c3 = (double)(k * k * k / 6.0L);
    return p;
}

// T[j] = 2^(j/2048) = 2^(hi/64) * 2^(lo/2048), hi = j >> 5, lo = j & 31: two 64/32-entry constant tables and
// one multiply per entry (relative error of an entry <= 1.5 ulp, part of the ~2 ulp budget above).
__device__ __forceinline__ void exp_table2_init(double* tbl, int tid, int nthreads) {
    static const double kHi[64] = {
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
        1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951
    };
    static const double kLo[32] = {
        1.0, 1.0003385080526823, 1.0006771306930664, 1.001015867959941,
        1.0013547198921082, 1.0016936865283832, 1.002032767907594, 1.0023719640685822,
        1.0027112750502025, 1.0030507008913223, 1.0033902416308227, 1.0037298973075977,
        1.004069667960554, 1.0044095536286128, 1.0047495543507072, 1.0050896701657839,
        1.0054299011128027, 1.005770247230737, 1.006110708558573, 1.00645128513531,
        1.0067919769999607, 1.0071327841915512, 1.0074737067491204, 1.0078147447117207,
        1.0081558981184175, 1.0084971670082898, 1.0088385514204294, 1.0091800513939415,
        1.0095216669679448, 1.0098633981815708, 1.0102052450739643, 1.0105472076842836
    };
    for (int j = tid; j < EXP_TBL2; j += nthreads) tbl[j] = kHi[j >> 5] * kLo[j & 31];
}

// out[i] = exp(P.kmul * d2[i]), four independent chains interleaved by hand.
__device__ __forceinline__ void exp_scaled_x4(const double (&d2)[4], double (&out)[4], const double* __restrict__ tbl,
                                              const ExpScaled& P) {
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: rounds to nearest integer
    double t[4], kd[4], r[4], p[4], T[4];
    int k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = fma(d2[i], P.kinv, kMagic);
#pragma unroll
    for (int i = 0; i < 4; ++i) { k[i] = __double2loint(t[i]); kd[i] = t[i] - kMagic; }
#pragma unroll
    for (int i = 0; i < 4; ++i) T[i] = tbl[k[i] & (EXP_TBL2 - 1)];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = fma(kd[i], -P.c, d2[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(r[i], P.c3, P.c2);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(p[i], r[i], P.c1);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = p[i] * r[i];
    // Fast path iff every k lies in [-1020*2048, 0].  The test uses the full 64-bit integer bits(t) - bits(magic)
    // (= k exactly while |k| < 2^51), not the 32-bit low word: for pairs more than ~1200 sigma apart
    // (|kmul d2| >= 2^31 ln2/2048 = 7.3e5) the low word wraps and would otherwise pass for a small exponent.
    // NaN / inf arguments land on the slow path as well.
    const long long kMagicBits = 0x4338000000000000LL;
    const long long kSpan = 1020LL * 2048LL;
    long long k64[4];
    unsigned long long umax = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        k64[i] = __double_as_longlong(t[i]) - kMagicBits;
        const unsigned long long u = (unsigned long long)(k64[i] + kSpan);
        umax = u > umax ? u : umax;
    }
    if (umax <= (unsigned long long)kSpan) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double res = fma(T[i], p[i], T[i]);
            out[i] = __hiloint2double(__double2hiint(res) + ((k[i] >> 11) << 20), __double2loint(res));
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double res = fma(T[i], p[i], T[i]);
        if (k64[i] < -kSpan || k64[i] > 0) {
            const double x = d2[i] * P.kmul;
            if (!(x >= -800.0)) {  // underflows to exactly 0 (as numpy's exp); NaN propagates
                out[i] = (x != x) ? x : 0.0;
                continue;
            }
            // -800 <= x: k64 >= -2.4e6 fits the low word; two-step scaling so subnormals round like IEEE
            const int e = (k[i] >> 11) + 1000;
            res = __hiloint2double(__double2hiint(res) + (e << 20), __double2loint(res)) * 0x1p-1000;
        } else {
            res = __hiloint2double(__double2hiint(res) + ((k[i] >> 11) << 20), __double2loint(res));
        }
        out[i] = res;
    }
}

}  // namespace socks
