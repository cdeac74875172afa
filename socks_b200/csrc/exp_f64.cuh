// Table-driven fp64 exp for the fused kernels: exp(x) = 2^e * T[j] * (1 + p(r)),
//   k = rint(x * 64/ln2) = 64 e + j,  r = x - k ln2/64 (Cody-Waite hi/lo),  |r| <= ln2/128,
//   p(r) = r + r^2/2 + ... + r^5/120  (truncation r^6/720 < 2^-54),  T[j] = 2^(j/64) in shared memory.
// 10 FP64-pipe instructions (CUDA's exp: 14 DFMA + ~4 more), max error ~2 ulp (2.2e-16 relative, checked
// offline against 60-digit arithmetic on 2e5 arguments in [-700, 0]); results below 2^-1022 go through a
// two-step scaling so subnormals round like IEEE and x < -746 gives exactly 0, as numpy's exp does.
// The materialising Gram kernels keep CUDA's exp (1 ulp): they are HBM-write bound anyway.
#pragma once
#include <string.h>

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

// ---- Scaled variant for the fused predict kernel: exp(kmul * d2) for a launch-constant kmul < 0, with kmul
// folded into every constant so that the multiply disappears: k = rint(d2 * kmul * 256/ln2) = 256 e + j,
// r' = d2 - k * ln2/(256 kmul) (hi/lo split computed on the host in extended precision), r = kmul r',
// exp(r) - 1 = r'(c1 + r'(c2 + r'(c3 + r' c4))) with c_i = kmul^i / i!  (|r| <= ln2/512: r^5/120 < 2^-54),
// T[j] = 2^(j/256).  9 FP64-pipe instructions per value.
constexpr int EXP_TBL2 = 256;

struct ExpScaled {
    double kinv;      // kmul * 256 / ln2
    double chi, clo;  // ln2 / (256 kmul) split: k * chi exact for |k| < 2^24
    double c1, c2, c3, c4;
    double kmul;
};

inline ExpScaled make_exp_scaled(double kmul) {
    ExpScaled p;
    const long double ln2 = 0.693147180559945309417232121458176568L;
    p.kmul = kmul;
    p.kinv = (double)((long double)kmul * 256.0L / ln2);
    const long double c = ln2 / (256.0L * (long double)kmul);
    double hi = (double)c;
    unsigned long long bits;
    memcpy(&bits, &hi, 8);
    bits &= ~((1ull << 24) - 1);
    memcpy(&hi, &bits, 8);
    p.chi = hi;
    p.clo = (double)(c - (long double)hi);
    const long double k = kmul;
    p.c1 = (double)k;
    p.c2 = (double)(k * k / 2.0L);
    p.c3 = (double)(k * k * k / 6.0L);
    p.c4 = (double)(k * k * k * k / 24.0L);
    return p;
}

__device__ __forceinline__ void exp_table2_init(double* tbl, int tid, int nthreads) {
    static const double kTbl2[EXP_TBL2] = {
        1.0, 1.0027112750502025, 1.0054299011128027, 1.0081558981184175,
        1.0108892860517005, 1.0136300849514894, 1.016378314910953, 1.019133996077738,
        1.0218971486541166, 1.0246677928971357, 1.0274459491187637, 1.030231637686041,
        1.0330248790212284, 1.0358256936019572, 1.0386341019613787, 1.041450124688316,
        1.0442737824274138, 1.0471050958792898, 1.0499440858006872, 1.0527907730046264,
        1.0556451783605572, 1.0585073227945128, 1.061377227289262, 1.0642549128844645,
        1.0671404006768237, 1.0700337118202419, 1.0729348675259756, 1.075843889062791,
        1.0787607977571199, 1.0816856149932152, 1.0846183622133092, 1.0875590609177697,
        1.0905077326652577, 1.0934643990728858, 1.0964290818163769, 1.099401802630222,
        1.102382583307841, 1.1053714457017412, 1.1083684117236787, 1.1113735033448175,
        1.1143867425958924, 1.1174081515673693, 1.1204377524096067, 1.12347556733302,
        1.1265216186082418, 1.129575928566288, 1.1326385195987192, 1.1357094141578055,
        1.1387886347566916, 1.1418762039695616, 1.1449721444318042, 1.148076478840179,
        1.1511892299529827, 1.154310420590216, 1.1574400736337511, 1.1605782120274988,
        1.1637248587775775, 1.1668800369524817, 1.1700437696832502, 1.1732160801636373,
        1.1763969916502812, 1.1795865274628758, 1.182784710984341, 1.1859915656609938,
        1.189207115002721, 1.1924313825831512, 1.1956643920398273, 1.1989061670743806,
        1.202156731452703, 1.2054161090051239, 1.2086843236265816, 1.2119613992768012,
        1.215247359980469, 1.2185422298274085, 1.2218460329727576, 1.2251587936371455,
        1.22848053610687, 1.2318112847340759, 1.2351510639369334, 1.2384998981998165,
        1.241857812073484, 1.245224830175258, 1.2486009771892048, 1.2519862778663162,
        1.255380757024691, 1.2587844395497165, 1.2621973503942507, 1.2656195145788063,
        1.2690509571917332, 1.2724917033894028, 1.275941778396392, 1.2794012075056693,
        1.2828700160787783, 1.2863482295460256, 1.2898358734066657, 1.2933329732290895,
        1.2968395546510096, 1.3003556433796506, 1.3038812651919358, 1.3074164459346773,
        1.3109612115247644, 1.3145155879493546, 1.318079601266064, 1.3216532776031575,
        1.3252366431597413, 1.3288297242059544, 1.3324325470831615, 1.3360451382041458,
        1.339667524053303, 1.3432997311868353, 1.3469417862329458, 1.3505937158920345,
        1.3542555469368927, 1.3579273062129011, 1.3616090206382248, 1.365300717204012,
        1.3690024229745905, 1.3727141650876684, 1.3764359707545302, 1.380167867260238,
        1.383909881963832, 1.387662042298529, 1.3914243757719262, 1.3951969099662003,
        1.3989796725383112, 1.4027726912202048, 1.4065759938190154, 1.4103896082172707,
        1.4142135623730951, 1.4180478843204152, 1.4218926021691656, 1.4257477441054942,
        1.42961333839197, 1.433489413367789, 1.4373759974489824, 1.4412731191286257,
        1.4451808069770467, 1.449099089642035, 1.4530279958490526, 1.4569675544014438,
        1.460917794180647, 1.4648787441464057, 1.4688504333369818, 1.4728328908693675,
        1.4768261459394993, 1.4808302278224719, 1.4848451658727524, 1.488870989524397,
        1.4929077282912648, 1.4969554117672355, 1.5010140696264256, 1.5050837316234065,
        1.5091644275934228, 1.5132561874526098, 1.5173590411982147, 1.5214730189088146,
        1.5255981507445384, 1.529734466947287, 1.533881997840956, 1.5380407738316568,
        1.5422108254079407, 1.5463921831410214, 1.550584877685, 1.5547889397770887,
        1.559004400237837, 1.5632312899713576, 1.567469639965553, 1.5717194812923414,
        1.5759808451078865, 1.5802537626528246, 1.5845382652524937, 1.588834384317164,
        1.593142151342267, 1.597461597908627, 1.6017927556826934, 1.606135656416771,
        1.6104903319492543, 1.6148568142048607, 1.6192351351948637, 1.6236253270173289,
        1.6280274218573478, 1.632441451987275, 1.6368674497669644, 1.6413054476440063,
        1.645755478153965, 1.6502175739206177, 1.6546917676561943, 1.6591780921616162,
        1.6636765803267364, 1.6681872651305825, 1.6727101796415966, 1.6772453570178785,
        1.681792830507429, 1.6863526334483934, 1.6909247992693053, 1.6955093614893326,
        1.7001063537185235, 1.7047158096580513, 1.709337763100463, 1.713972247929926,
        1.718619298122478, 1.723278947746274, 1.7279512309618377, 1.732636182022311,
        1.7373338352737062, 1.7420442251551564, 1.746767386199169, 1.7515033530318782,
        1.7562521603732995, 1.761013843037584, 1.7657884359332727, 1.7705759740635547,
        1.7753764925265212, 1.7801900265154245, 1.785016611318935, 1.789856282321401,
        1.7947090750031072, 1.7995750249405351, 1.804454167806624, 1.809346539371032,
        1.8142521755003989, 1.8191711121586085, 1.8241033854070534, 1.8290490314048973,
        1.8340080864093424, 1.8389805867758937, 1.843966568958626, 1.8489660695104508,
        1.8539791250833855, 1.8590057724288205, 1.864046048397789, 1.8690999899412386,
        1.8741676341103, 1.8792490180565602, 1.8843441790323345, 1.8894531543909392,
        1.8945759815869656, 1.8997126981765553, 1.9048633418176741, 1.9100279502703899,
        1.9152065613971474, 1.9203992131630474, 1.925605943636125, 1.930826790987627,
        1.9360617934922943, 1.9413109895286405, 1.9465744175792332, 1.9518521162309783,
        1.9571441241754002, 1.9624504802089273, 1.9677712232331759, 1.9731063922552343,
        1.978456026387951, 1.9838201648502194, 1.9891988469672663, 1.9945921121709402
    };
    for (int j = tid; j < EXP_TBL2; j += nthreads) tbl[j] = kTbl2[j];
}

// out[i] = exp(P.kmul * d2[i]), four independent chains interleaved by hand.
__device__ __forceinline__ void exp_scaled_x4(const double (&d2)[4], double (&out)[4], const double* __restrict__ tbl,
                                              const ExpScaled& P) {
    const double kMagic = 6755399441055744.0;
    double t[4], kd[4], r[4], p[4], T[4];
    int k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = fma(d2[i], P.kinv, kMagic);
#pragma unroll
    for (int i = 0; i < 4; ++i) { k[i] = __double2loint(t[i]); kd[i] = t[i] - kMagic; }
#pragma unroll
    for (int i = 0; i < 4; ++i) T[i] = tbl[k[i] & (EXP_TBL2 - 1)];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = fma(kd[i], -P.chi, d2[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = fma(kd[i], -P.clo, r[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(r[i], P.c4, P.c3);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(p[i], r[i], P.c2);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fma(p[i], r[i], P.c1);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = p[i] * r[i];
    // results near or below the normal range (x < ~-707) need the two-step scaling: test once for all four
    const int kmin = min(min(k[0], k[1]), min(k[2], k[3]));
    if ((kmin >> 8) >= -1020) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double res = fma(T[i], p[i], T[i]);
            out[i] = __hiloint2double(__double2hiint(res) + ((k[i] >> 8) << 20), __double2loint(res));
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double res = fma(T[i], p[i], T[i]);
        int e = k[i] >> 8;
        if (e < -1020) {
            const double x = d2[i] * P.kmul;
            if (!(x >= -800.0)) {
                out[i] = (x != x) ? x : 0.0;
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
