// Fused KernelSR predict (gym_socks/algorithms/reach/kernel_sr.py:329-342): K(X,pts) is never written.
//
// With the fitted value functions known, all T-1 recursion steps collapse into ONE contraction
//     acc[r][j] = sum_i coef[i][r] k(x_i, p_j),   coef[.][0] = w (normaliser), coef[.][r] = vf[t0+r] * w
// and out[t0+r-1][j] = step(p_j, acc[r][j] / acc[0][j]).  The contraction runs on the fp64 tensor cores
// (DMMA.8x8x4): each lane evaluates the kernel value of one (sample, point) pair in registers and uses it AS IS
// as the B fragment (lane l holds B[k = l%4][n = l/4]); the 16 coefficient rows are the A fragments.
//
// Two evaluations of k share that skeleton:
//  * factored (default): the ratio acc[r]/acc[0] is invariant under a per-column scale, and
//        k(x,p) = exp(-g|p'|^2) * exp(-g|x'|^2) * exp(2g p'.x'),   x' = x - c, p' = p - c  (c = centre of X's box).
//    exp(-g|x'|^2) is folded into the packed coefficients once per fit, exp(-g|p'|^2) cancels, and the inner
//    loop evaluates only exp(2g p'.x'): a d-term dot product in units of ln2/2048 (the bandwidth is folded into
//    the per-thread point coordinates), a 2048-entry table of 2^(j/2048) in shared memory and a degree-2
//    polynomial -- 8 FP64-pipe instructions per pair at d = 2 (DMUL DFMA | DADD DADD DADD | DFMA DMUL DFMA)
//    against 11 for the direct form.  Error: |f| <= 1/2 table step => cubic term < (ln2/4096)^3/6 = 8.1e-13
//    relative per kernel value, and both sums have positive weights (w = diag(inv G) > 0, vf >= 0), so the
//    ratio moves by < 2e-12 (the parity bar is 1e-8).  A column is only evaluated this way when every factor is
//    far from over/underflow (g|p'|^2, 2g|p'|R_x, gR_x^2 < 600) and its un-scaled normaliser is >= 1e-280, i.e.
//    when no term of the reference's own sum can have underflowed; every other column is recomputed by the
//    direct form below, so 0/0 -> NaN and exact zeros appear exactly where numpy puts them.
//  * direct: squared distance by differences + table exp with full IEEE range handling (exp_f64.cuh); used for
//    whole launches when the factored form is not admissible (g R_x^2 >= 600) and, one thread per column with
//    CUDA's exp, for the columns the factored pass flagged.
// Both sum the sample axis in fixed 2048-sample splits combined in order, and the choice between them depends
// only on the column and on the fitted state, so every output column is bit-identical however the test points
// are chunked or sharded.
//
// Work per pair: N*M*(3d+2+E+2T) algorithmic flops; 8*M*(d+T) bytes of I/O.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>

#include "exp_f64.cuh"
#include "umma.cuh"

namespace socks {

constexpr int SR_CP = 20;        // coef row stride (16 rows + pad; 20 mod 16 = 4 -> conflict-free A fragments)
constexpr int SP_TS = 64;        // samples per shared-memory stage
constexpr int SP_SPLIT = 2048;   // samples per CTA (blockIdx.y): FIXED, so the summation order does not depend on M
constexpr int SP_HDR = 16;       // header doubles of the packed state: [0] factored-ok, [1] R_x^2, [2..10) centre
constexpr double SP_RANGE = 600.0;

// Integer tensor-core engine (sr_predict_i8_kernel below)
constexpr int PI_KB = 32;            // samples per K-block (one tcgen05.mma kind::i8 K step)
constexpr int PI_SPLIT = 4096;       // samples per CTA = one exact int32 accumulation: 8 pairs * 255^2 * 4096 < 2^31
constexpr int PI_STAGES = 4;         // K-blocks in flight in tensor memory (stages of the kernel-value digits)
constexpr int PI_LSTAGES = 16;       // K-blocks in flight in shared memory (coefficient digits + samples): hides the L2 latency
constexpr int PI_TBL = 256;          // table of 2^(j/256) ...
constexpr int PI_REP = 16;           // ... replicated 16x: lane l reads copy l % 16, so a lookup never bank-conflicts
constexpr int PI_TW = 128;           // points per CTA = tensor-memory lanes
constexpr int PI_KDIG = 7;           // digits (bytes) of a kernel value: 51-bit fixed point, k = sum_s d_s 2^(-3 - 8 s)
constexpr int PI_GROUPS = 10;        // digit groups kept (weight of group g: 2^(-10 - 8 g)); the rest is < 2^-87 of the scale
constexpr double PI_TAU = 0.001953125;  // 2^-9: columns with a smaller normaliser (relative) take the direct form


__host__ __device__ inline int64_t sp_npad(int64_t n) { return (n + SP_TS - 1) / SP_TS * SP_TS; }
inline int sp_blocks(int T) { return T <= 1 ? 1 : (T - 1 + SOCKS_SR_ROWS - 2) / (SOCKS_SR_ROWS - 1); }

// ---- pack: fitted state -> centred samples + coefficient rows (runs once per fit) --------------------------
// hdr[2+k] = centre of the bounding box of X; hdr[1] = max_i |x_i - c|^2; hdr[0] = 1 if the factored form is
// admissible for this (X, bandwidth).  One CTA: N*d is a few 10^4 values.
__global__ void __launch_bounds__(1024) sr_center_kernel(const double* __restrict__ X, int64_t n, int d, double kmul,
                                                         double* __restrict__ hdr) {
    __shared__ double slo[32], shi[32], sc[SOCKS_MAX_DIM];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = 0; k < d; ++k) {
        double lo = INFINITY, hi = -INFINITY;
        for (int64_t i = tid; i < n; i += blockDim.x) {
            const double v = X[i * d + k];
            lo = fmin(lo, v);
            hi = fmax(hi, v);
        }
        for (int off = 16; off > 0; off >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
        }
        if (lane == 0) { slo[warp] = lo; shi[warp] = hi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fmin(lo, slo[w]); hi = fmax(hi, shi[w]); }
            sc[k] = 0.5 * lo + 0.5 * hi;
        }
        __syncthreads();
    }
    double r2 = 0.0;
    for (int64_t i = tid; i < n; i += blockDim.x) {
        double a = 0.0;
        for (int k = 0; k < d; ++k) {
            const double t = X[i * d + k] - sc[k];
            a = fma(t, t, a);
        }
        r2 = fmax(r2, a);
    }
    for (int off = 16; off > 0; off >>= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, off));
    if (lane == 0) shi[warp] = r2;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r2 = fmax(r2, shi[w]);
        const bool ok = d <= 8 && kmul < 0.0 && (-kmul) * r2 < SP_RANGE;  // false for NaN/inf samples
        hdr[0] = ok ? 1.0 : 0.0;
        hdr[1] = r2;
        for (int k = 0; k < 8; ++k) hdr[2 + k] = (k < d) ? sc[k] : 0.0;
    }
}

// coefE[b][i][0] = w_i, coefE[b][i][r] = vf[15 b + r][i] * w_i (direct form);  coefF = coefE * exp(kmul |x_i - c|^2)
// (factored form);  Xc[i] = x_i - c.  Rows beyond the block and samples beyond n are zero.
__global__ void sr_pack_kernel(const double* __restrict__ X, const double* __restrict__ w,
                               const double* __restrict__ vf, int64_t ldvf, int64_t n, int64_t n_pad, int d, int T,
                               int nblk, double kmul, const double* __restrict__ hdr, double* __restrict__ Xc,
                               double* __restrict__ coefE, double* __restrict__ coefF) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const bool live = i < n;
    double q2 = 0.0;
    for (int k = 0; k < d; ++k) {
        const double t = live ? X[i * d + k] - (k < 8 ? hdr[2 + k] : 0.0) : 0.0;
        Xc[i * d + k] = t;
        q2 = fma(t, t, q2);
    }
    const double wi = live ? w[i] : 0.0;
    const double e = (hdr[0] != 0.0 && live) ? exp(kmul * q2) : 0.0;
    for (int b = 0; b < nblk; ++b) {
        const int t0 = b * (SOCKS_SR_ROWS - 1);
        const int R = min(T - 1 - t0, SOCKS_SR_ROWS - 1) + 1;
        double* ce = coefE + ((int64_t)b * n_pad + i) * SR_CP;
        double* cf = coefF + ((int64_t)b * n_pad + i) * SR_CP;
        for (int r = 0; r < SR_CP; ++r) {
            double c = 0.0;
            if (live && r < R) c = (r == 0) ? wi : vf[(int64_t)(t0 + r) * ldvf + i] * wi;
            ce[r] = c;
            cf[r] = c * e;
        }
    }
}

// ---- operands of the integer tensor-core engine (once per fit) -----------------------------------------------
// cscale[b][r] = 2^e > max_i coefE[b][i][r] (1 for an all-zero row)
__global__ void __launch_bounds__(256) pi_rowscale_kernel(const double* __restrict__ coefE, int64_t n_pad,
                                                          double* __restrict__ cscale) {
    __shared__ double sm[256];
    const int b = blockIdx.x / 16, r = blockIdx.x % 16;
    const double* c = coefE + (int64_t)b * n_pad * SR_CP + r;
    double mx = 0.0;
    for (int64_t i = threadIdx.x; i < n_pad; i += 256) mx = fmax(mx, c[i * SR_CP]);
    sm[threadIdx.x] = mx;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + off]);
        __syncthreads();
    }
    if (threadIdx.x == 0) cscale[blockIdx.x] = (sm[0] > 0.0 && sm[0] < INFINITY) ? ldexp(1.0, ilogb(sm[0]) + 1) : 1.0;
}

// Core-matrix layout of an (R rows x 32 bytes) K-major operand tile (see umma.cuh): [r / 8][16-byte chunk c][r % 8][16 B].
__device__ __forceinline__ int pi_core_offset(int r, int c) { return (r >> 3) * 256 + c * 128 + (r & 7) * 16; }

// coefD[b][kb]: the B operand of one K-block, 128 rows x 32 bytes: row 16 t + r = digit t (byte 7 - t, unsigned) of
// rint(coefE[b][i][r] / cscale[b][r] * 2^63), column = sample i % 32.
// xsD[kb][i % 32] = (x'_i[0..d), cx_i) with cx_i = kmul |x'_i|^2 * 256/ln2: the sample's share of the exponent.
// tbl[j][rep] = 2^51 2^(j/256).
__global__ void pi_pack_kernel(const double* __restrict__ Xc, const double* __restrict__ coefE,
                               const double* __restrict__ cscale, int64_t n_pad, int d, int nblk, double kmul,
                               uint8_t* __restrict__ coefD, double* __restrict__ xsD, double* __restrict__ tbl) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < PI_TBL * PI_REP) tbl[i] = exp2((double)(i / PI_REP) / PI_TBL) * 0x1p51;  // pre-scaled: see the producers
    if (i >= n_pad) return;
    const int64_t kb = i / PI_KB;
    const int e = (int)(i % PI_KB);
    double q2 = 0.0;
    double* xs = xsD + i * (d + 1);
    for (int k = 0; k < d; ++k) {
        const double t = Xc[i * d + k];
        xs[k] = t;
        q2 = fma(t, t, q2);
    }
    xs[d] = kmul * q2 * (PI_TBL / 0.693147180559945309417232121458176568);
    for (int b = 0; b < nblk; ++b) {
        uint8_t* tile = coefD + ((int64_t)b * (n_pad / PI_KB) + kb) * 4096;
        for (int r = 0; r < 16; ++r) {
            const double c = coefE[((int64_t)b * n_pad + i) * SR_CP + r] / cscale[b * 16 + r];  // in [0, 1)
            const unsigned long long q = __double2ull_rn(c * 0x1p63);
#pragma unroll
            for (int t = 0; t < 8; ++t)
                tile[pi_core_offset(16 * t + r, e >> 4) + (e & 15)] = (uint8_t)(q >> (8 * (7 - t)));
        }
    }
}

template <int D>
__device__ __forceinline__ void sp_load_stage(double* __restrict__ sx, double* __restrict__ sc,
                                              const double* __restrict__ X, const double* __restrict__ coef,
                                              int64_t n_rows, int64_t s0, int tid, int nthreads) {
    // coefficient rows: SP_TS x 20 doubles in 16-byte chunks, contiguous in global memory
    const double* csrc = coef + s0 * SR_CP;
    for (int c = tid; c < SP_TS * SR_CP / 2; c += nthreads) cp_async16(sc + c * 2, csrc + c * 2);
    for (int e = tid; e < SP_TS * D; e += nthreads) {
        const int64_t i = s0 + e / D;
        sx[e] = (i < n_rows) ? __ldg(X + s0 * D + e) : 0.0;
    }
}

// accumulator rows r = lq (+8) of this split -> part[split][r][j], 16-byte stores
template <int Q>
__device__ __forceinline__ void sp_store_partials(const double (&acc)[Q][2][2], double* __restrict__ part, int64_t ldp,
                                                  int64_t pbase, int64_t m, int lq, int lr) {
    double* pbase_out = part + (int64_t)blockIdx.y * 16 * ldp;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int64_t j = pbase + q * 8 + 2 * lr;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double* dst = pbase_out + (int64_t)(lq + 8 * h) * ldp + j;
            if (j + 1 < m) {
                *reinterpret_cast<double2*>(dst) = make_double2(acc[q][h][0], acc[q][h][1]);
            } else if (j < m) {
                dst[0] = acc[q][h][0];
            }
        }
    }
}

// ---- per-column finish shared by the in-kernel reduction of the factored pass and the finish kernel of the direct pass
struct FinishArgs {
    int* counters;          // one arrival counter per point tile (zeroed before the launch)
    int* list;              // [0] = count, [1..] = flagged columns
    const double* tubes;
    double* out;
    int64_t ldout;
    double kmul;
    int problem, T, t0, R, write_last, build_list;
};

// sum[0] = normaliser, sum[r] = numerator of step t0 + r - 1.  After a factored pass the column only stands if no factor
// could have over/underflowed and its un-scaled normaliser is >= 1e-280 (see the header); otherwise it is queued for the
// direct form.  Also writes out[T-1] = 1_target (kernel_sr.py:57) when write_last.
template <int ROWS>
__device__ __forceinline__ void finish_column(const double (&sum)[ROWS], int64_t j, const double* __restrict__ p, int d,
                                              const double* __restrict__ hdr, const FinishArgs& fa) {
    if (fa.write_last)
        fa.out[(int64_t)(fa.T - 1) * fa.ldout + j] = in_box(p, fa.tubes + (int64_t)(fa.T - 1) * 4 * d, 1, d) ? 1.0 : 0.0;
    if (fa.R <= 1) return;
    const double den = sum[0];
    double q2 = 0.0;
    for (int k = 0; k < d; ++k) {
        const double t = p[k] - hdr[2 + k];
        q2 = fma(t, t, q2);
    }
    const double g = -fa.kmul;
    bool ok = (g * q2 < SP_RANGE) && (2.0 * g * sqrt(q2 * hdr[1]) < SP_RANGE);
    ok = ok && (den * exp(fa.kmul * q2) >= 1e-280) && (den < INFINITY);  // false for NaN
    if (!ok) {
        if (fa.build_list) fa.list[1 + atomicAdd(fa.list, 1)] = (int)j;
        return;
    }
#pragma unroll
    for (int r = 1; r < ROWS; ++r)
        if (r < fa.R) {
            const int t = fa.t0 + r - 1;
            fa.out[(int64_t)t * fa.ldout + j] = sr_step(p, fa.tubes + (int64_t)t * 4 * d, d, fa.problem, sum[r] / den);
        }
}

// ---- factored form ------------------------------------------------------------------------------------------
// CTA = WARPS warps x (8 Q) points x one 2048-sample split.  Lane (lq = lane/4, lr = lane%4) evaluates sample
// s = 4 ks + lr against the points {8 q + lq}: exactly the B-fragment layout of mma.m8n8k4.
// ps = p' * (2 g * 2048/ln2): the dot product s = ps . x' is the exponent in table steps; t = s + 1.5*2^52 rounds
// it (k = low word of t), f = s - k in [-1/2, 1/2], 2^(f/2048) = 1 + f (d1 + f d2), 2^(k/2048) = 2^(k>>11) * tbl[k & 2047].
template <int D, int WARPS, int MINB, int Q, bool HI, int U>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    sr_predict_fast_kernel(const double* __restrict__ Xc, const double* __restrict__ coefF, int64_t n_pad,
                           const double* __restrict__ pts, int64_t m, const double* __restrict__ hdr, double pscale,
                           int force_direct, double* __restrict__ part, FinishArgs fa) {
    if (force_direct || hdr[0] == 0.0) return;  // launch-uniform: the direct kernel handles this launch
    __shared__ __align__(16) double sx[2][SP_TS * D];
    __shared__ __align__(16) double sc[2][SP_TS * SR_CP];
    __shared__ double etbl[EXP_TBL2];
    __shared__ int s_last;
    constexpr int TW = WARPS * Q * 8;     // points per CTA
    constexpr int ROWS = HI ? 16 : 8;      // coefficient rows carried
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lq = lane >> 2, lr = lane & 3;
    // blockIdx.x = sample split (fastest: the CTAs of one point tile are scheduled together, so the tile's partial
    // sums are still in L2 when its last CTA reduces them), blockIdx.y = point tile
    const int split = blockIdx.x, nsplit = gridDim.x;
    const int64_t tile = blockIdx.y;
    const int64_t pbase = tile * TW + warp * (Q * 8);
    exp_table2_init(etbl, tid, WARPS * 32);  // visible after the first __syncthreads of the main loop

    double ps[Q][D];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int64_t j = min(pbase + q * 8 + lq, m - 1);
#pragma unroll
        for (int k = 0; k < D; ++k) ps[q][k] = (__ldg(pts + j * D + k) - hdr[2 + k]) * pscale;
    }
    double acc[Q][2][2];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int h = 0; h < 2; ++h) acc[q][h][0] = acc[q][h][1] = 0.0;

    const double kMagic = 6755399441055744.0;          // 1.5 * 2^52
    const double kD2 = 5.72744624533987700e-08;         // (ln2 / 2048)^2 / 2
    // ln2 / 2048, held in a register: as a literal the compiler re-materialises it (two moves) in every iteration,
    // and every non-FP64 instruction in this loop costs about one FP64-pipe cycle (profiles/r2_predict_*.txt)
    double kD1;
    asm("mov.f64 %0, 0d3F362E42FEFA39EF;" : "=d"(kD1));
    const char* etb = reinterpret_cast<const char*>(etbl);
    const int tile0 = split * (SP_SPLIT / SP_TS);
    const int tile1 = min((int)(n_pad / SP_TS), tile0 + SP_SPLIT / SP_TS);
    sp_load_stage<D>(sx[0], sc[0], Xc, coefF, n_pad, (int64_t)tile0 * SP_TS, tid, WARPS * 32);
    cp_async_commit();
    for (int tile = tile0; tile < tile1; ++tile) {
        const int cur = (tile - tile0) & 1;
        cp_async_wait<0>();
        __syncthreads();  // stage `cur` visible to all; everyone is done reading stage cur^1
        if (tile + 1 < tile1)
            sp_load_stage<D>(sx[cur ^ 1], sc[cur ^ 1], Xc, coefF, n_pad, (int64_t)(tile + 1) * SP_TS, tid, WARPS * 32);
        cp_async_commit();
        // this lane's sample row (lr) and coefficient row (lq) of the stage; the unrolled body addresses them with
        // immediate offsets and the pointers advance once per U sample quads
        const double* xp = sx[cur] + lr * D;
        const double* cp = sc[cur] + lr * SR_CP + lq;
#pragma unroll 1
        for (int kb = 0; kb < SP_TS / 4; kb += U) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                double xs[D];
#pragma unroll
                for (int k = 0; k < D; ++k) xs[k] = xp[u * 4 * D + k];
                const double a0 = cp[u * 4 * SR_CP];
                const double a1 = HI ? cp[u * 4 * SR_CP + 8] : 0.0;
                double sv[Q], t[Q], f[Q], pl[Q], tb[Q];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    double a = ps[q][0] * xs[0];
#pragma unroll
                    for (int k = 1; k < D; ++k) a = fma(ps[q][k], xs[k], a);
                    sv[q] = a;
                }
#pragma unroll
                for (int q = 0; q < Q; ++q) t[q] = sv[q] + kMagic;
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    tb[q] = *reinterpret_cast<const double*>(etb + ((__double2loint(t[q]) << 3) & ((EXP_TBL2 - 1) << 3)));
#pragma unroll
                // k back to double by I2F.F64 (conversion unit) rather than t - magic (one more FP64-pipe DADD): measured
                // 2 % faster, same bits (k is exact either way)
                for (int q = 0; q < Q; ++q) f[q] = sv[q] - __int2double_rn(__double2loint(t[q]));
#pragma unroll
                for (int q = 0; q < Q; ++q) pl[q] = fma(f[q], kD2, kD1);
#pragma unroll
                for (int q = 0; q < Q; ++q) pl[q] = pl[q] * f[q];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const double res = fma(tb[q], pl[q], tb[q]);
                    // 2^(k >> 11): high word += (k >> 11) << 20 = (k & ~2047) * 512, as LOP3 + IMAD (the compiler's
                    // own choice is shift + mask + add)
                    int hi2;
                    asm("{\n\t.reg .b32 km;\n\tand.b32 km, %1, 0xfffff800;\n\tmad.lo.s32 %0, km, 512, %2;\n\t}"
                        : "=r"(hi2)
                        : "r"(__double2loint(t[q])), "r"(__double2hiint(res)));
                    const double kv = __hiloint2double(hi2, __double2loint(res));
                    dmma884(acc[q][0][0], acc[q][0][1], a0, kv);
                    if (HI) dmma884(acc[q][1][0], acc[q][1][1], a1, kv);
                }
            }
            xp += U * 4 * D;
            cp += U * 4 * SR_CP;
        }
    }
    // ---- epilogue: this split's partial sums -> part[tile][split][ROWS][TW]; the LAST CTA of the tile to arrive sums the
    // splits in index order (deterministic, independent of arrival order), divides, applies the step and writes the
    // result rows.  Nothing but the result leaves the chip: the partials are read back from L2 and then discarded.
    double* tpart = part + (tile * nsplit + split) * (int64_t)(ROWS * TW);
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int h = 0; h < (HI ? 2 : 1); ++h)
            *reinterpret_cast<double2*>(tpart + (lq + 8 * h) * TW + warp * (Q * 8) + q * 8 + 2 * lr) =
                make_double2(acc[q][h][0], acc[q][h][1]);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(fa.counters + tile, 1) == nsplit - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const double* tbase = part + tile * nsplit * (int64_t)(ROWS * TW);
    for (int col = tid; col < TW; col += WARPS * 32) {
        const int64_t j = tile * TW + col;
        if (j >= m) continue;
        double sum[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) sum[r] = 0.0;
        for (int sp = 0; sp < nsplit; ++sp) {
            const double* src = tbase + (int64_t)sp * (ROWS * TW) + col;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) sum[r] += __ldcg(src + r * TW);
        }
        finish_column<ROWS>(sum, j, pts + j * D, D, hdr, fa);
    }
    __syncthreads();  // every thread has read its columns: the tile's partials are dead
    {
        const char* base = reinterpret_cast<const char*>(tbase);
        const int lines = nsplit * ROWS * TW * 8 / 128;
        for (int l = tid; l < lines; l += WARPS * 32)
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + (int64_t)l * 128) : "memory");
    }
}

// ---- integer tensor-core engine ----------------------------------------------------------------------------------
// Same contraction acc[r][j] = sum_i coef[i][r] k(x_i, p_j), but off the FP64 pipe: the kernel value k in (0, 1] is turned
// into a 51-bit fixed-point integer whose BYTES are unsigned 8-bit digits, the coefficients (>= 0, row-scaled) likewise
// (once per fit), and every digit product is an exact int8 x int8 -> int32 MMA (tcgen05.mma kind::i8, accumulators in
// tensor memory).  The FP64 pipe is left with the exponent and the exponential alone (10 instructions per pair), the 32
// DMMA cycles per 32 pairs of the fp64 kernel are gone.
//   * CTA = 128 points (TMEM lane = point) x one fixed 4096-sample split; 4 PW producer warps (warp w serves the points
//     of lane quarter w % 4 and the samples of slice w / 4 of every K-block: PW warps per scheduler hide the latencies of
//     the dependent FP64 chain), then one loader lane (bulk copies) and one MMA-issuing lane.
//   * Per K-block (32 samples) a producer thread evaluates its point against the 32 samples,
//         a = ps . x' + cx_i + cp   (exponent in units of ln2/256, expanded form of -g |x - p|^2),
//         2^(a/256) = 2^(k >> 8) * tbl[k & 255] * (1 + f (d1 + f (d2 + f d3))),   k = rint(a), f = a - k,
//     folds the power of two into the table value's exponent field and adds 2^52: the mantissa field of the sum IS the
//     51-bit fixed-point value q = rint(k 2^51) (no F2I/I2F -- 64-bit conversions run at one lane per clock and scheduler
//     -- and no variable shifts), transposes the bytes of 4 q's at a time (PRMT) and writes the 7 digit rows of its lane
//     straight into tensor memory (tcgen05.st): the A operand never touches shared memory.
//   * The B operand of the K-block stacks the 8 coefficient digits of the 16 rows: 128 "columns" n = 16 t + r.  The MMA of
//     kernel digit s targets D columns 16 s .. 16 s + 127, so the product (s, t) lands in group g = s + t at columns
//     16 g + r: 7 MMAs (128 x 128 x 32) per K-block accumulate all 56 digit pairs into 14 groups of 16 columns.
//   * Epilogue: groups 0..9 -> exact 64-bit integer merge -> fp64 -> x cscale[r] -> partial sums of this split; the last
//     CTA of the tile sums the splits in index order and finishes the column exactly like the fp64 kernel.  Integer sums
//     are exact, the split boundaries are fixed: every column is bit-identical however the points are chunked or sharded.
// MEASURED (configs[1], 250 000 points): 8.5 ms against 8.0 ms for the fp64 kernel above, so this engine is NOT the default
// (variant 6).  The contraction is indeed gone from the FP64 pipe (tensor pipe 23 % busy, FP64 29 %), but every
// instruction that goes to a 16-lane pipe (FP64, ALU, IMAD) holds the scheduler's dispatch port for two cycles.  The
// first version (62-bit values, variable shifts: ~35 such instructions per pair-warp = 70 cycles against the 60 of the
// fp64 kernel; ncu: alu 45 % + fp64 29 % + lsu 15 % + fma 8 % + adu 6 % = 103 % of the issue bandwidth) took 9.5 ms; the
// 51-bit extraction without shifts (~27 instructions) 8.5 ms.
// A column stands only if the expanded exponent is accurate for it (same range test as the factored fp64 form), the
// point is finite and its normaliser is >= 2^-9 of the scale (then the fixed-point error is < 3e-9 relative in the worst
// case, ~1e-13 typically); every other
// column is queued for the direct form, which reproduces numpy's underflow / 0/0 behaviour.
template <int D>
struct PiCfg {
    static constexpr int XS_BYTES = PI_KB * (D + 1) * 8;
    static constexpr int STAGE_BYTES = 4096 + XS_BYTES;  // coefficient digits + samples of one K-block
    static constexpr int TBL_BYTES = PI_TBL * PI_REP * 8;
    static constexpr int NBAR = 2 * PI_LSTAGES + 2 * PI_STAGES + 2;
    static constexpr int SMEM_BYTES = TBL_BYTES + PI_LSTAGES * STAGE_BYTES + NBAR * 8 + 16;
};

template <int D, int PW>  // PW producer warps per 32 points: each evaluates 32 / PW samples of a K-block
__global__ void __launch_bounds__(128 * PW + 64, 1)
    sr_predict_i8_kernel(const uint8_t* __restrict__ coefD, const double* __restrict__ xsD, const double* __restrict__ tblG,
                         const double* __restrict__ cscale, int nkb_total, const double* __restrict__ pts, int64_t m,
                         const double* __restrict__ hdr, double* __restrict__ part, FinishArgs fa) {
    if (hdr[0] == 0.0) return;  // launch-uniform: the direct kernels handle this launch
    using Cfg = PiCfg<D>;
    extern __shared__ __align__(1024) uint8_t pi_smem[];
    const double* tbl = reinterpret_cast<const double*>(pi_smem);
    uint8_t* stages = pi_smem + Cfg::TBL_BYTES;
    uint64_t* ld_full = reinterpret_cast<uint64_t*>(stages + PI_LSTAGES * Cfg::STAGE_BYTES);
    uint64_t* ld_empty = ld_full + PI_LSTAGES;
    uint64_t* a_full = ld_empty + PI_LSTAGES;
    uint64_t* empty = a_full + PI_STAGES;
    uint64_t* acc_full = empty + PI_STAGES;
    uint64_t* tbl_full = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tbl_full + 1);
    __shared__ int s_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int split = blockIdx.x, nsplit = gridDim.x;
    const int64_t tile = blockIdx.y;
    const int kb0 = split * (PI_SPLIT / PI_KB);
    const int nk = min(nkb_total - kb0, PI_SPLIT / PI_KB);
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    if (tid == 32) {
        for (int i = 0; i < PI_LSTAGES; ++i) {
            umma::mbar_init(ld_full + i, 1);
            umma::mbar_init(ld_empty + i, 1);
        }
        for (int i = 0; i < PI_STAGES; ++i) {
            umma::mbar_init(a_full + i, 128 * PW);
            umma::mbar_init(empty + i, 1);
        }
        umma::mbar_init(acc_full, 1);
        umma::mbar_init(tbl_full, 1);
        umma::mbar_fence_init();
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t A_COL = 256;  // kernel-value digits: stage st at columns 256 + 64 st, digit s at + 8 s

    if (warp == 4 * PW) {
        if (lane == 0) {
            umma::mbar_arrive_expect_tx(tbl_full, Cfg::TBL_BYTES);
            umma::bulk_g2s(pi_smem, tblG, Cfg::TBL_BYTES, tbl_full);
            for (int i = 0; i < nk; ++i) {
                const int st = i % PI_LSTAGES, round = i / PI_LSTAGES;
                if (round > 0) umma::mbar_wait(ld_empty + st, (uint32_t)(round - 1) & 1u);
                uint8_t* dst = stages + st * Cfg::STAGE_BYTES;
                umma::mbar_arrive_expect_tx(ld_full + st, Cfg::STAGE_BYTES);
                umma::bulk_g2s(dst, coefD + (int64_t)(kb0 + i) * 4096, 4096, ld_full + st);
                umma::bulk_g2s(dst + 4096, xsD + (int64_t)(kb0 + i) * (PI_KB * (D + 1)), Cfg::XS_BYTES, ld_full + st);
            }
        }
    } else if (warp == 4 * PW + 1) {
        if (lane == 0) {
            const uint32_t idesc = umma::idesc_i8(128, 128, 0, 0);  // unsigned digits
            for (int i = 0; i < nk; ++i) {
                const int st = i % PI_STAGES, round = i / PI_STAGES;
                const int ls = i % PI_LSTAGES, lround = i / PI_LSTAGES;
                umma::mbar_wait(ld_full + ls, (uint32_t)lround & 1u);
                umma::mbar_wait(a_full + st, (uint32_t)round & 1u);
                umma::tc_fence_after();
                const uint64_t db = umma::smem_desc_kmajor_noswizzle(umma::smem_u32(stages + ls * Cfg::STAGE_BYTES), 128, 256);
#pragma unroll
                for (int sd = 0; sd < PI_KDIG; ++sd)
                    umma::mma_i8_ts(tmem + (uint32_t)(16 * sd), tmem + A_COL + (uint32_t)(64 * st + 8 * sd), db, idesc, 1u);
                umma::mma_commit(empty + st);
                umma::mma_commit(ld_empty + ls);  // (the producers read the samples of this stage before they arrived)
            }
            umma::mma_commit(acc_full);
        }
    } else {
        // ---- producers: thread = (point, sample slice)
        constexpr int NS = PI_KB / PW, NW = NS / 4;  // samples / digit words of this warp per K-block
        const int quarter = warp & 3, sub = warp >> 2;
        const int col = quarter * 32 + lane;          // point of the tile = TMEM lane
        const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
        if (sub == 0) {  // zero the 15 accumulator groups of this lane (every MMA accumulates)
            const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int c = 0; c < 240; c += 8) umma::tmem_st8(trow + (uint32_t)c, z);
        }
        const int64_t j = min(tile * PI_TW + col, m - 1);
        const double L = -fa.kmul * (PI_TBL / 0.693147180559945309417232121458176568);  // g * 256/ln2
        double ps[D], cp = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const double t = __ldg(pts + j * D + k) - hdr[2 + k];
            ps[k] = 2.0 * L * t;
            cp = fma(-L * t, t, cp);
        }
        const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
        const double d1 = 2.70760617406228636e-03;  // ln2 / 256
        const double d2 = 3.66556559691664515e-06;  // d1^2 / 2
        const double d3 = 3.30823122780754510e-09;  // d1^3 / 6
        const char* tb_lane = reinterpret_cast<const char*>(tbl) + ((lane & 15) << 3);
        umma::mbar_wait(tbl_full, 0);
        for (int i = 0; i < nk; ++i) {
            const int st = i % PI_STAGES, round = i / PI_STAGES;
            const int ls = i % PI_LSTAGES, lround = i / PI_LSTAGES;
            umma::mbar_wait(ld_full + ls, (uint32_t)lround & 1u);
            const double* xs = reinterpret_cast<const double*>(stages + ls * Cfg::STAGE_BYTES + 4096) + sub * NS * (D + 1);
            uint32_t dig[PI_KDIG][NW];  // [digit][word = 4 samples]
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                uint32_t qh[4], ql[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const double* x = xs + (w * 4 + e) * (D + 1);
                    double a = x[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) a = fma(ps[k], x[k], a);
                    a += cp;
                    const double t = a + kMagic;
                    const int ki = __double2loint(t);
                    const double f = a - (t - kMagic);  // no I2F / F2I below: 64-bit conversions run at 1 lane/clk/SMSP
                    const double tb = *reinterpret_cast<const double*>(tb_lane + ((ki & (PI_TBL - 1)) << 7));  // 2^51 2^(j/256)
                    double pl = fma(f, d3, d2);
                    pl = fma(f, pl, d1);
                    pl = fma(f, pl, 1.0);
                    // the power of two goes into the table value's exponent field (values below 2^-1000 do not matter),
                    // then y = 2^52 + rint(k 2^51): the mantissa field IS the 51-bit fixed-point value -- no shifts
                    const double tbs = __hiloint2double(__double2hiint(tb) + (max(ki >> 8, -1000) << 20), __double2loint(tb));
                    const double y = fma(tbs, pl, 4503599627370496.0);
                    qh[e] = (uint32_t)__double2hiint(y) & 0x000FFFFFu;
                    ql[e] = (uint32_t)__double2loint(y);
                }
                // bytes of 4 q's -> one word per digit (4 x 4 byte transposes; byte 7 is always 0)
                {
                    const uint32_t a01 = __byte_perm(qh[0], qh[1], 0x5140), b01 = __byte_perm(qh[0], qh[1], 0x7362);
                    const uint32_t a23 = __byte_perm(qh[2], qh[3], 0x5140), b23 = __byte_perm(qh[2], qh[3], 0x7362);
                    dig[2][w] = __byte_perm(a01, a23, 0x5410);  // byte 4 of q
                    dig[1][w] = __byte_perm(a01, a23, 0x7632);  // byte 5
                    dig[0][w] = __byte_perm(b01, b23, 0x5410);  // byte 6 (top digit: 3 bits)
                }
                {
                    const uint32_t a01 = __byte_perm(ql[0], ql[1], 0x5140), b01 = __byte_perm(ql[0], ql[1], 0x7362);
                    const uint32_t a23 = __byte_perm(ql[2], ql[3], 0x5140), b23 = __byte_perm(ql[2], ql[3], 0x7362);
                    dig[6][w] = __byte_perm(a01, a23, 0x5410);  // byte 0
                    dig[5][w] = __byte_perm(a01, a23, 0x7632);
                    dig[4][w] = __byte_perm(b01, b23, 0x5410);
                    dig[3][w] = __byte_perm(b01, b23, 0x7632);  // byte 3
                }
            }
            if (round > 0) {
                umma::mbar_wait(empty + st, (uint32_t)(round - 1) & 1u);  // the MMAs that read this TMEM stage are done
                umma::tc_fence_after();
            }
#pragma unroll
            for (int sd = 0; sd < PI_KDIG; ++sd) {
                const uint32_t ta = trow + A_COL + (uint32_t)(64 * st + 8 * sd + sub * NW);
                if constexpr (NW == 8) umma::tmem_st8(ta, dig[sd]);
                else if constexpr (NW == 4) umma::tmem_st4(ta, dig[sd]);
                else umma::tmem_st2(ta, dig[sd]);
            }
            umma::tmem_st_wait();
            umma::tc_fence_before();
            umma::mbar_arrive(a_full + st);
        }
        // ---- epilogue of the split: groups -> fp64 partial sums
        if (sub == 0) {
        umma::mbar_wait(acc_full, 0);
        umma::tc_fence_after();
        double val[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) val[r] = 0.0;
#pragma unroll
        for (int g = PI_GROUPS - 2; g >= 0; g -= 2) {  // pairs (g, g + 1), smallest weights first
            uint32_t u0[16], u1[16];
            umma::tmem_ld16(trow + (uint32_t)(16 * g), u0);
            umma::tmem_ld16(trow + (uint32_t)(16 * (g + 1)), u1);
            umma::tmem_ld_wait();
            const double wgt = __longlong_as_double((long long)(1023 - 10 - 8 * (g + 1)) << 52);  // 2^(-10 - 8 (g + 1))
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const long long pr = ((long long)(int)u0[r] << 8) + (long long)(int)u1[r];  // < 2^40, exact
                val[r] = fma(__longlong_as_double(pr + 0x4338000000000000LL) - 0x1.8p52, wgt, val[r]);
            }
        }
        double* tpart = part + (tile * nsplit + split) * (int64_t)(16 * PI_TW);
#pragma unroll
        for (int r = 0; r < 16; ++r) tpart[r * PI_TW + col] = val[r] * cscale[r];
        }
    }
    umma::tc_fence_before();
    __threadfence();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
    if (tid == 0) s_last = (atomicAdd(fa.counters + tile, 1) == nsplit - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const double* tbase = part + tile * nsplit * (int64_t)(16 * PI_TW);
    if (tid < PI_TW) {
        const int64_t j = tile * PI_TW + tid;
        if (j < m) {
            double sum[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) sum[r] = 0.0;
            for (int sp = 0; sp < nsplit; ++sp) {
                const double* src = tbase + (int64_t)sp * (16 * PI_TW) + tid;
#pragma unroll
                for (int r = 0; r < 16; ++r) sum[r] += __ldcg(src + r * PI_TW);
            }
            const double* p = pts + j * D;
            if (fa.write_last)
                fa.out[(int64_t)(fa.T - 1) * fa.ldout + j] = in_box(p, fa.tubes + (int64_t)(fa.T - 1) * 4 * D, 1, D) ? 1.0 : 0.0;
            if (fa.R > 1) {
                double q2 = 0.0;
                bool finite = true;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double t = p[k] - hdr[2 + k];
                    q2 = fma(t, t, q2);
                    finite = finite && (fabs(p[k]) < INFINITY);
                }
                const double g = -fa.kmul;
                bool ok = finite && (g * q2 < SP_RANGE) && (2.0 * g * sqrt(q2 * hdr[1]) < SP_RANGE);
                ok = ok && (sum[0] >= PI_TAU * cscale[0]);
                if (!ok) {
                    if (fa.build_list) fa.list[1 + atomicAdd(fa.list, 1)] = (int)j;
                } else {
#pragma unroll
                    for (int r = 1; r < 16; ++r)
                        if (r < fa.R) {
                            const int t = fa.t0 + r - 1;
                            fa.out[(int64_t)t * fa.ldout + j] =
                                sr_step(p, fa.tubes + (int64_t)t * 4 * D, D, fa.problem, sum[r] / sum[0]);
                        }
                }
            }
        }
    }
    __syncthreads();
    {
        const char* base = reinterpret_cast<const char*>(tbase);
        const int lines = nsplit * 16 * PI_TW * 8 / 128;
        for (int l = tid; l < lines; l += 128 * PW + 64)
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + (int64_t)l * 128) : "memory");
    }
}

// ---- direct form (whole launch): distance by differences + table exp with IEEE range handling ---------------
template <int D, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    sr_predict_dmma_kernel(const double* __restrict__ X, const double* __restrict__ coef, int64_t n, int64_t n_pad,
                           const double* __restrict__ pts, int64_t m, const double* __restrict__ hdr, ExpScaled ex,
                           int rows, int force_direct, double* __restrict__ part, int64_t ldp) {
    if (!force_direct && hdr[0] != 0.0) return;  // launch-uniform: the factored kernel handled this launch
    constexpr int Q = 4;
    __shared__ __align__(16) double sx[2][SP_TS * D];
    __shared__ __align__(16) double sc[2][SP_TS * SR_CP];
    __shared__ double etbl[EXP_TBL2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lq = lane >> 2, lr = lane & 3;
    const int64_t pbase = (int64_t)blockIdx.x * (WARPS * Q * 8) + warp * (Q * 8);
    exp_table2_init(etbl, tid, WARPS * 32);

    double p[Q][D];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int64_t j = min(pbase + q * 8 + lq, m - 1);
#pragma unroll
        for (int k = 0; k < D; ++k) p[q][k] = __ldg(pts + j * D + k);
    }
    double acc[Q][2][2];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int h = 0; h < 2; ++h) acc[q][h][0] = acc[q][h][1] = 0.0;

    const bool hi_rows = rows > 8;
    const int tile0 = blockIdx.y * (SP_SPLIT / SP_TS);
    const int tile1 = min((int)(n_pad / SP_TS), tile0 + SP_SPLIT / SP_TS);
    sp_load_stage<D>(sx[0], sc[0], X, coef, n, (int64_t)tile0 * SP_TS, tid, WARPS * 32);
    cp_async_commit();
    for (int tile = tile0; tile < tile1; ++tile) {
        const int cur = (tile - tile0) & 1;
        cp_async_wait<0>();
        __syncthreads();
        if (tile + 1 < tile1)
            sp_load_stage<D>(sx[cur ^ 1], sc[cur ^ 1], X, coef, n, (int64_t)(tile + 1) * SP_TS, tid, WARPS * 32);
        cp_async_commit();
        const double* x_s = sx[cur];
        const double* c_s = sc[cur];
#pragma unroll 2
        for (int ks = 0; ks < SP_TS / 4; ++ks) {
            const int s = ks * 4 + lr;
            double xs[D];
#pragma unroll
            for (int k = 0; k < D; ++k) xs[k] = x_s[s * D + k];
            const double a0 = c_s[s * SR_CP + lq];
            const double a1 = c_s[s * SR_CP + lq + 8];
            double arg[Q], kv[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                double d2 = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double t = p[q][k] - xs[k];
                    d2 = fma(t, t, d2);
                }
                arg[q] = d2;
            }
            exp_scaled_x4(arg, kv, etbl, ex);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                dmma884(acc[q][0][0], acc[q][0][1], a0, kv[q]);
                if (hi_rows) dmma884(acc[q][1][0], acc[q][1][1], a1, kv[q]);
            }
        }
    }
    sp_store_partials<Q>(acc, part, ldp, pbase, m, lq, lr);
}

// ---- finish of the DIRECT pass: sum the splits in order, divide by the normaliser row, apply the FHT/THT step;
// also out[T-1] = 1_target(pts) (kernel_sr.py:57) when write_last.  (The factored pass finishes inside its own kernel.)
__global__ void sr_predict_finish_kernel(const double* __restrict__ part, int64_t ldp, int splits, int64_t m, int R,
                                         int t0, const double* __restrict__ pts, int d,
                                         const double* __restrict__ tubes, int problem, int T,
                                         const double* __restrict__ hdr, int force_direct, int write_last,
                                         double* __restrict__ out, int64_t ldout) {
    if (!force_direct && hdr[0] != 0.0) return;  // launch-uniform: the factored kernel already wrote the result
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double* p = pts + j * d;
    if (write_last)
        out[(int64_t)(T - 1) * ldout + j] = in_box(p, tubes + (int64_t)(T - 1) * 4 * d, 1, d) ? 1.0 : 0.0;
    if (R <= 1) return;
    double den = 0.0;
    for (int s = 0; s < splits; ++s) den += part[(int64_t)s * 16 * ldp + j];
    for (int r = 1; r < R; ++r) {
        double num = 0.0;
        for (int s = 0; s < splits; ++s) num += part[((int64_t)s * 16 + r) * ldp + j];
        const int t = t0 + r - 1;
        out[(int64_t)t * ldout + j] = sr_step(p, tubes + (int64_t)t * 4 * d, d, problem, num / den);
    }
}

// ---- direct form, one thread per listed column (list == NULL: all columns): CUDA's exp, plain FMAs ---------
// Runs over the columns the factored pass flagged, and as the whole-launch path for d > 8 and for kmul >= 0.
__global__ void __launch_bounds__(128)
    sr_predict_cols_kernel(const double* __restrict__ X, const double* __restrict__ coefE, int64_t n, int64_t n_pad,
                           int d, const double* __restrict__ pts, int64_t m, const int* __restrict__ list, double kmul,
                           const double* __restrict__ tubes, int problem, int T, int nblk, double* __restrict__ out,
                           int64_t ldout) {
    const int64_t count = list ? (int64_t)list[0] : m;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < count; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = list ? (int64_t)list[1 + g] : g;
        const double* p = pts + j * d;
        for (int b = 0; b < nblk; ++b) {
            const int t0 = b * (SOCKS_SR_ROWS - 1);
            const int R = min(T - 1 - t0, SOCKS_SR_ROWS - 1) + 1;
            const double* coef = coefE + (int64_t)b * n_pad * SR_CP;
            double acc[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) acc[r] = 0.0;
            for (int64_t i = 0; i < n; ++i) {
                double d2 = 0.0;
                for (int k = 0; k < d; ++k) {
                    const double t = p[k] - __ldg(X + i * d + k);
                    d2 = fma(t, t, d2);
                }
                const double kv = exp(d2 * kmul);
                const double* c = coef + i * SR_CP;
#pragma unroll
                for (int r = 0; r < 16; ++r) acc[r] = fma(__ldg(c + r), kv, acc[r]);
            }
#pragma unroll
            for (int r = 1; r < 16; ++r)
                if (r < R) {
                    const int t = t0 + r - 1;
                    out[(int64_t)t * ldout + j] = sr_step(p, tubes + (int64_t)t * 4 * d, d, problem, acc[r] / acc[0]);
                }
        }
    }
}

// variant 1: FMA-pipe contraction, one point per thread, samples staged in shared memory (cross-check of the
// tensor-core kernels with CUDA's exp; also what the roofline comparison in profiles/ calls "FMA variant").
template <int D>
__global__ void __launch_bounds__(128)
    sr_predict_fma_kernel(const double* __restrict__ X, const double* __restrict__ coef, int64_t n, int64_t n_pad,
                          const double* __restrict__ pts, int64_t m, double neg_gamma,
                          const double* __restrict__ tubes, int problem, int t0, int R, double* __restrict__ out,
                          int64_t ldout) {
    __shared__ __align__(16) double sx[2][SP_TS * D];
    __shared__ __align__(16) double sc[2][SP_TS * SR_CP];
    const int tid = threadIdx.x;
    const int64_t j = (int64_t)blockIdx.x * 128 + tid;
    const int64_t jc = min(j, m - 1);
    double p[D];
#pragma unroll
    for (int k = 0; k < D; ++k) p[k] = __ldg(pts + jc * D + k);
    double acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = 0.0;

    const int ntiles = (int)(n_pad / SP_TS);
    sp_load_stage<D>(sx[0], sc[0], X, coef, n, 0, tid, 128);
    cp_async_commit();
    for (int tile = 0; tile < ntiles; ++tile) {
        const int cur = tile & 1;
        cp_async_wait<0>();
        __syncthreads();
        if (tile + 1 < ntiles)
            sp_load_stage<D>(sx[cur ^ 1], sc[cur ^ 1], X, coef, n, (int64_t)(tile + 1) * SP_TS, tid, 128);
        cp_async_commit();
        const double* x_s = sx[cur];
        const double* c_s = sc[cur];
#pragma unroll 2
        for (int s = 0; s < SP_TS; ++s) {
            double d2 = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const double t = p[k] - x_s[s * D + k];
                d2 = fma(t, t, d2);
            }
            const double kv = exp(d2 * neg_gamma);
            const double2* c2 = reinterpret_cast<const double2*>(c_s + s * SR_CP);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double2 c = c2[r];
                acc[2 * r] = fma(c.x, kv, acc[2 * r]);
                acc[2 * r + 1] = fma(c.y, kv, acc[2 * r + 1]);
            }
        }
    }
    if (j < m) {
#pragma unroll
        for (int r = 1; r < 16; ++r)
            if (r < R) {
                const int t = t0 + r - 1;
                out[(int64_t)t * ldout + j] =
                    sr_step(pts + j * D, tubes + (int64_t)t * 4 * D, D, problem, acc[r] / acc[0]);
            }
    }
}

__global__ void sr_last_row_kernel(const double* __restrict__ pts, int64_t m, int d, const double* __restrict__ tubes,
                                   int T, double* __restrict__ out, int64_t ldout) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    out[(int64_t)(T - 1) * ldout + j] = in_box(pts + j * d, tubes + (int64_t)(T - 1) * 4 * d, 1, d) ? 1.0 : 0.0;
}

struct PackView {
    const double *hdr, *Xc, *coefE, *coefF;
    // integer tensor-core engine: table, row scales, samples per K-block, coefficient digits per block and K-block
    const double *tbl, *cscale, *xsD;
    const uint8_t* coefD;
    int64_t n_pad;
    int nblk;
};

static inline int64_t pack_doubles(int64_t n, int d, int T) {
    const int64_t n_pad = sp_npad(n), nblk = sp_blocks(T);
    const int64_t fp64_part = SP_HDR + n_pad * d + 2 * nblk * n_pad * SR_CP;
    const int64_t i8_part = PI_TBL * PI_REP + nblk * 16 + n_pad * (d + 1) + nblk * (n_pad / PI_KB) * (4096 / 8);
    return fp64_part + i8_part + 64;
}
static inline PackView pack_view(const double* pack, int64_t n, int d, int T) {
    PackView v;
    v.n_pad = sp_npad(n);
    v.nblk = sp_blocks(T);
    v.hdr = pack;
    v.Xc = pack + SP_HDR;
    v.coefE = v.Xc + v.n_pad * d;
    v.coefF = v.coefE + (int64_t)v.nblk * v.n_pad * SR_CP;
    v.tbl = v.coefF + (int64_t)v.nblk * v.n_pad * SR_CP;  // every offset is a multiple of 2 doubles: 16-byte aligned
    v.cscale = v.tbl + PI_TBL * PI_REP;
    v.xsD = v.cscale + (int64_t)v.nblk * 16;
    v.coefD = reinterpret_cast<const uint8_t*>(v.xsD + v.n_pad * (d + 1));
    return v;
}
static inline int predict_splits(int64_t n) { return (int)ceil_div(sp_npad(n), SP_SPLIT); }
static inline int64_t predict_part_doubles(int64_t n, int64_t m) {
    return (int64_t)predict_splits(n) * 16 * round_up(m, 256);  // whole point tiles: [tile][split][rows][tile width]
}
static inline int64_t predict_counter_ints(int64_t m) { return ceil_div(m, 64) + 2; }

// CTA shape per launch: fewer warps per CTA when there are few point tiles, so that every SM still receives
// >= ~24 (tile, split) items and the last, partially filled wave is short.  The shape changes which warp evaluates a
// column, never the order of its sum.
static inline bool small_tiles(int64_t m, int splits, int pts_per_cta) {
    return ceil_div(m, pts_per_cta) * splits < 148 * 24;
}

static int g_pi_pw = 4;  // producer warps per lane quarter (2 or 4); SOCKS_PREDICT_I8_PW overrides (measurement)

template <int D>
static int launch_i8(const PackView& v, const double* pts, int64_t m, int b, double* part, const FinishArgs& fa,
                     cudaStream_t st) {
    static std::atomic<bool> done[SOCKS_MAX_DEVICES];  // dynamic-smem opt-in is per device
    constexpr int smem = 160 * 1024;                    // > half an SM: one CTA per SM (each allocates all of tensor memory)
    static_assert(PiCfg<D>::SMEM_BYTES <= smem, "stage ring");
    int dev = 0;
    SOCKS_CUDA(cudaGetDevice(&dev));
    const bool tracked = dev >= 0 && dev < SOCKS_MAX_DEVICES;
    if (!tracked || !done[dev].load(std::memory_order_acquire)) {
        SOCKS_CUDA(cudaFuncSetAttribute((sr_predict_i8_kernel<D, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        SOCKS_CUDA(cudaFuncSetAttribute((sr_predict_i8_kernel<D, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (tracked) done[dev].store(true, std::memory_order_release);
    }
    static const int pw_env = [] {
        const char* e = getenv("SOCKS_PREDICT_I8_PW");
        return e ? atoi(e) : 0;
    }();
    if (pw_env == 2 || pw_env == 4) g_pi_pw = pw_env;
    const int nkb = (int)(v.n_pad / PI_KB);
    const int splits = (int)ceil_div(v.n_pad, PI_SPLIT);
    if (g_pi_pw == 2)
        sr_predict_i8_kernel<D, 2><<<dim3((unsigned)splits, (unsigned)ceil_div(m, PI_TW)), 128 * 2 + 64, smem, st>>>(
            v.coefD + (int64_t)b * nkb * 4096, v.xsD, v.tbl, v.cscale + b * 16, nkb, pts, m, v.hdr, part, fa);
    else
        sr_predict_i8_kernel<D, 4><<<dim3((unsigned)splits, (unsigned)ceil_div(m, PI_TW)), 128 * 4 + 64, smem, st>>>(
            v.coefD + (int64_t)b * nkb * 4096, v.xsD, v.tbl, v.cscale + b * 16, nkb, pts, m, v.hdr, part, fa);
    return SOCKS_OK;
}

template <int D>
static void launch_tensor(const PackView& v, const double* X, int64_t n, const double* pts, int64_t m, double kmul,
                          int b, int R, int force_direct, int shape, double* part, int64_t ldp, int splits,
                          const FinishArgs& fa, cudaStream_t st) {
    constexpr int Q = (D <= 4) ? 4 : 2;
    const double* cf = v.coefF + (int64_t)b * v.n_pad * SR_CP;
    const double* ce = v.coefE + (int64_t)b * v.n_pad * SR_CP;
    const double pscale = -2.0 * kmul * (2048.0 / 0.693147180559945309417232121458176568);
    const bool small = shape == 2 || (shape == 0 && small_tiles(m, splits, 8 * Q * 8));
#define SOCKS_FAST_LAUNCH(W, MB, HI, U)                                                                              \
    sr_predict_fast_kernel<D, W, MB, Q, HI, U><<<dim3((unsigned)splits, (unsigned)ceil_div(m, W * Q * 8)), W * 32, 0, st>>>( \
        v.Xc, cf, v.n_pad, pts, m, v.hdr, pscale, force_direct, part, fa)
    // 8 warps x 2 CTAs/SM (unroll 2) and 4 warps x 4 CTAs/SM (unroll 4): best of the shape sweep in
    // profiles/r2_predict_shape_sweep.json (all shapes within 4 %)
    if (shape == 3) {  // integer tensor-core engine
        if (!force_direct) launch_i8<D>(v, pts, m, b, part, fa, st);
    } else if (small) {
        if (R > 8) SOCKS_FAST_LAUNCH(4, 4, true, 4); else SOCKS_FAST_LAUNCH(4, 4, false, 4);
    } else {
        if (R > 8) SOCKS_FAST_LAUNCH(8, 2, true, 2); else SOCKS_FAST_LAUNCH(8, 2, false, 2);
    }
#undef SOCKS_FAST_LAUNCH
    const ExpScaled ex = make_exp_scaled(kmul);
    // 2 CTAs/SM (<= 128 registers): at 3 the direct form spilled 60-464 bytes per thread (d = 2..8)
    sr_predict_dmma_kernel<D, 8, 2><<<dim3((unsigned)ceil_div(m, 256), (unsigned)splits), 256, 0, st>>>(
        X, ce, n, v.n_pad, pts, m, v.hdr, ex, R, force_direct, part, ldp);
}

}  // namespace socks

using namespace socks;

extern "C" int64_t socks_sr_pack_bytes(int64_t n, int d, int T) { return pack_doubles(n, d, T) * 8; }

extern "C" int socks_sr_pack(const double* X, const double* w, const double* vf, int64_t ldvf, int64_t n, int d,
                             double scale, int divide, int T, double* pack, socks_stream_t stream) {
    SOCKS_CHECK_ARG(X != nullptr, 1);
    SOCKS_CHECK_ARG(w != nullptr, 2);
    SOCKS_CHECK_ARG(vf != nullptr && ldvf >= n, 3);
    SOCKS_CHECK_ARG(n >= 1, 5);
    SOCKS_CHECK_ARG(d >= 1 && d <= SOCKS_MAX_DIM, 6);
    SOCKS_CHECK_ARG(scale == scale && (!divide || scale != 0.0), 7);
    SOCKS_CHECK_ARG(T >= 1, 9);
    SOCKS_CHECK_ARG(pack != nullptr && (reinterpret_cast<uintptr_t>(pack) & 15) == 0, 10);
    const double kmul = divide ? 1.0 / scale : scale;
    cudaStream_t st = (cudaStream_t)stream;
    const PackView v = pack_view(pack, n, d, T);
    sr_center_kernel<<<1, 1024, 0, st>>>(X, n, d, kmul, pack);
    SOCKS_CHECK_LAUNCH();
    sr_pack_kernel<<<(unsigned)ceil_div(v.n_pad, 128), 128, 0, st>>>(X, w, vf, ldvf, n, v.n_pad, d, T, v.nblk, kmul,
                                                                     pack, const_cast<double*>(v.Xc),
                                                                     const_cast<double*>(v.coefE),
                                                                     const_cast<double*>(v.coefF));
    SOCKS_CHECK_LAUNCH();
    if (d <= 8) {  // operands of the integer tensor-core engine
        pi_rowscale_kernel<<<(unsigned)(v.nblk * 16), 256, 0, st>>>(v.coefE, v.n_pad, const_cast<double*>(v.cscale));
        pi_pack_kernel<<<(unsigned)ceil_div(std::max<int64_t>(v.n_pad, PI_TBL * PI_REP), 128), 128, 0, st>>>(
            v.Xc, v.coefE, v.cscale, v.n_pad, d, v.nblk, kmul, const_cast<uint8_t*>(v.coefD), const_cast<double*>(v.xsD),
            const_cast<double*>(v.tbl));
        SOCKS_CHECK_LAUNCH();
    }
    return SOCKS_OK;
}

extern "C" int64_t socks_sr_predict_packed_work_bytes(int64_t n, int64_t m) {
    return (predict_part_doubles(n, m) + (predict_counter_ints(m) + m + 4) / 2 + 4 + 16) * 8;  // partials, counters, list
}

extern "C" int socks_sr_predict_packed(const double* pack, const double* X, int64_t n, int d, double scale,
                                       int divide, const double* pts, int64_t m, const double* tubes, int problem,
                                       int T, double* out, int64_t ldout, double* work, int variant,
                                       socks_stream_t stream) {
    SOCKS_CHECK_ARG(pack != nullptr && (reinterpret_cast<uintptr_t>(pack) & 15) == 0, 1);
    SOCKS_CHECK_ARG(X != nullptr, 2);
    SOCKS_CHECK_ARG(n >= 1, 3);
    SOCKS_CHECK_ARG(d >= 1 && d <= SOCKS_MAX_DIM, 4);
    SOCKS_CHECK_ARG(scale == scale && (!divide || scale != 0.0), 5);
    SOCKS_CHECK_ARG(m >= 0 && m < (int64_t)1 << 31, 8);
    SOCKS_CHECK_ARG(tubes != nullptr, 9);
    SOCKS_CHECK_ARG(problem == SOCKS_PROBLEM_THT || problem == SOCKS_PROBLEM_FHT, 10);
    SOCKS_CHECK_ARG(T >= 1, 11);
    SOCKS_CHECK_ARG(ldout >= m, 13);
    SOCKS_CHECK_ARG(work != nullptr && (reinterpret_cast<uintptr_t>(work) & 15) == 0, 14);
    // variant: 0 auto (factored form on the fp64 tensor cores, CTA shape by size) | 1 FMA-pipe kernel | 2 direct form on
    // the fp64 tensor cores | 3 / 4: auto with the CTA shape forced to 8 / 4 warps (measurement only) | 5: same as 0 |
    // 6: integer tensor-core engine (tcgen05.mma kind::i8; correct to 3e-14 of variant 0 but 6 % slower: see DESIGN 4.5)
    SOCKS_CHECK_ARG(variant >= 0 && variant <= 6, 15);
    if (m == 0) return SOCKS_OK;
    SOCKS_CHECK_ARG(pts != nullptr && out != nullptr, 7);
    const double kmul = divide ? 1.0 / scale : scale;
    cudaStream_t st = (cudaStream_t)stream;
    const PackView v = pack_view(pack, n, d, T);
    const unsigned gm = (unsigned)ceil_div(m, 256);
    if (T == 1) {
        sr_last_row_kernel<<<gm, 256, 0, st>>>(pts, m, d, tubes, T, out, ldout);
        SOCKS_CHECK_LAUNCH();
        return SOCKS_OK;
    }
    if (d > 8 || !(kmul < 0.0)) {  // no templated tensor-core kernel: one thread per column, CUDA's exp
        sr_last_row_kernel<<<gm, 256, 0, st>>>(pts, m, d, tubes, T, out, ldout);
        sr_predict_cols_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, st>>>(X, v.coefE, n, v.n_pad, d, pts, m, nullptr,
                                                                          kmul, tubes, problem, T, v.nblk, out, ldout);
        SOCKS_CHECK_LAUNCH();
        return SOCKS_OK;
    }
    // partial sums, 128-byte aligned (whole L2 lines are discarded after the in-kernel reduction)
    double* part = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(work) + 127) & ~(uintptr_t)127);
    int* counters = reinterpret_cast<int*>(part + predict_part_doubles(n, m));  // one per point tile, then the list
    const int64_t nctr = predict_counter_ints(m);
    int* list = counters + nctr;  // list[0] = number of flagged columns
    const int64_t ldp = round_up(m, 2);
    const int splits = predict_splits(n);
    const int force_direct = (variant == 2);
    const int shape = variant == 3 ? 1 : (variant == 4 ? 2 : (variant == 6 ? 3 : 0));
    SOCKS_CHECK_ARG(ceil_div(m, 64) <= 65535, 8);  // point tiles ride on gridDim.y
    for (int b = 0; b < v.nblk; ++b) {
        const int t0 = b * (SOCKS_SR_ROWS - 1);
        const int R = ((T - 1 - t0 < SOCKS_SR_ROWS - 1) ? (T - 1 - t0) : (SOCKS_SR_ROWS - 1)) + 1;
        if (variant == 1) {
            const double* ce = v.coefE + (int64_t)b * v.n_pad * SR_CP;
            if (b == 0) sr_last_row_kernel<<<gm, 256, 0, st>>>(pts, m, d, tubes, T, out, ldout);
#define SOCKS_FMA_CASE(DD)                                                                                         \
    case DD:                                                                                                       \
        sr_predict_fma_kernel<DD><<<(unsigned)ceil_div(m, 128), 128, 0, st>>>(X, ce, n, v.n_pad, pts, m, kmul, tubes, \
                                                                             problem, t0, R, out, ldout);          \
        break;
            switch (d) {
                SOCKS_FMA_CASE(1) SOCKS_FMA_CASE(2) SOCKS_FMA_CASE(3) SOCKS_FMA_CASE(4)
                SOCKS_FMA_CASE(5) SOCKS_FMA_CASE(6) SOCKS_FMA_CASE(7) SOCKS_FMA_CASE(8)
            }
#undef SOCKS_FMA_CASE
            SOCKS_CHECK_LAUNCH();
            continue;
        }
        // arrival counters of this launch (and, once, the flagged-column count)
        SOCKS_CUDA(cudaMemsetAsync(counters, 0, sizeof(int) * (size_t)(nctr + (b == 0 ? 1 : 0)), st));
        const FinishArgs fa{counters, list, tubes, out, ldout, kmul, problem, T, t0, R, b == 0, b == 0};
#define SOCKS_PRED_CASE(DD)                                                                                    \
    case DD:                                                                                                   \
        launch_tensor<DD>(v, X, n, pts, m, kmul, b, R, force_direct, shape, part, ldp, splits, fa, st);        \
        break;
        switch (d) {
            SOCKS_PRED_CASE(1) SOCKS_PRED_CASE(2) SOCKS_PRED_CASE(3) SOCKS_PRED_CASE(4)
            SOCKS_PRED_CASE(5) SOCKS_PRED_CASE(6) SOCKS_PRED_CASE(7) SOCKS_PRED_CASE(8)
        }
#undef SOCKS_PRED_CASE
        SOCKS_CHECK_LAUNCH();
        sr_predict_finish_kernel<<<gm, 256, 0, st>>>(part, ldp, splits, m, R, t0, pts, d, tubes, problem, T, v.hdr,
                                                     force_direct, b == 0, out, ldout);
        SOCKS_CHECK_LAUNCH();
    }
    if (variant != 1 && !force_direct) {
        const unsigned gc = (unsigned)std::min<int64_t>(ceil_div(m, 128), 148 * 8);
        sr_predict_cols_kernel<<<gc, 128, 0, st>>>(X, v.coefE, n, v.n_pad, d, pts, m, list, kmul, tubes, problem, T,
                                                   v.nblk, out, ldout);
        SOCKS_CHECK_LAUNCH();
    }
    return SOCKS_OK;
}

// Pack + predict in one call (the pack lives at the head of `work`): the form a caller without fitted-state
// caching uses.  KernelSR.predict packs once per fit and calls socks_sr_predict_packed.
extern "C" int64_t socks_sr_predict_work_bytes(int64_t n, int64_t m) {
    // d and T are not arguments of this (round-1) helper: size the pack for the largest supported d and one block
    // per 15 steps up to T = 256
    return socks_sr_pack_bytes(n, SOCKS_MAX_DIM, 256) + socks_sr_predict_packed_work_bytes(n, m);
}

extern "C" int socks_sr_predict(const double* X, const double* w, const double* vf, int64_t ldvf, int64_t n, int d,
                                double scale, int divide, const double* pts, int64_t m, const double* tubes,
                                int problem, int T, double* out, int64_t ldout, double* work, int variant,
                                socks_stream_t stream) {
    SOCKS_CHECK_ARG(X != nullptr, 1);
    SOCKS_CHECK_ARG(w != nullptr, 2);
    SOCKS_CHECK_ARG(vf != nullptr && ldvf >= n, 3);
    SOCKS_CHECK_ARG(n >= 1, 5);
    SOCKS_CHECK_ARG(d >= 1 && d <= SOCKS_MAX_DIM, 6);
    SOCKS_CHECK_ARG(scale == scale && (!divide || scale != 0.0), 7);
    SOCKS_CHECK_ARG(m >= 0, 10);
    SOCKS_CHECK_ARG(tubes != nullptr, 11);
    SOCKS_CHECK_ARG(problem == SOCKS_PROBLEM_THT || problem == SOCKS_PROBLEM_FHT, 12);
    SOCKS_CHECK_ARG(T >= 1 && T <= 256, 13);
    SOCKS_CHECK_ARG(ldout >= m, 15);
    SOCKS_CHECK_ARG(work != nullptr && (reinterpret_cast<uintptr_t>(work) & 15) == 0, 16);
    SOCKS_CHECK_ARG(variant >= 0 && variant <= 6, 17);
    if (m == 0) return SOCKS_OK;
    SOCKS_CHECK_ARG(pts != nullptr && out != nullptr, 9);
    int rc = socks_sr_pack(X, w, vf, ldvf, n, d, scale, divide, T, work, stream);
    if (rc) return rc;
    double* rest = work + round_up(pack_doubles(n, d, T), 2);
    rc = socks_sr_predict_packed(work, X, n, d, scale, divide, pts, m, tubes, problem, T, out, ldout, rest, variant,
                                 stream);
    return rc;
}
