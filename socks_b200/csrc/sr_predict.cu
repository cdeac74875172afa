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
#include <cmath>

#include "exp_f64.cuh"

namespace socks {

constexpr int SR_CP = 20;        // coef row stride (16 rows + pad; 20 mod 16 = 4 -> conflict-free A fragments)
constexpr int SP_TS = 64;        // samples per shared-memory stage
constexpr int SP_SPLIT = 2048;   // samples per CTA (blockIdx.y): FIXED, so the summation order does not depend on M
constexpr int SP_HDR = 16;       // header doubles of the packed state: [0] factored-ok, [1] R_x^2, [2..10) centre
constexpr double SP_RANGE = 600.0;


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
    int64_t n_pad;
    int nblk;
};

static inline int64_t pack_doubles(int64_t n, int d, int T) {
    const int64_t n_pad = sp_npad(n);
    return SP_HDR + n_pad * d + 2 * (int64_t)sp_blocks(T) * n_pad * SR_CP + 64;
}
static inline PackView pack_view(const double* pack, int64_t n, int d, int T) {
    PackView v;
    v.n_pad = sp_npad(n);
    v.nblk = sp_blocks(T);
    v.hdr = pack;
    v.Xc = pack + SP_HDR;
    v.coefE = v.Xc + v.n_pad * d;
    v.coefF = v.coefE + (int64_t)v.nblk * v.n_pad * SR_CP;
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
    if (small) {
        if (R > 8) SOCKS_FAST_LAUNCH(4, 4, true, 4); else SOCKS_FAST_LAUNCH(4, 4, false, 4);
    } else {
        if (R > 8) SOCKS_FAST_LAUNCH(8, 2, true, 2); else SOCKS_FAST_LAUNCH(8, 2, false, 2);
    }
#undef SOCKS_FAST_LAUNCH
    const ExpScaled ex = make_exp_scaled(kmul);
    sr_predict_dmma_kernel<D, 8, 3><<<dim3((unsigned)ceil_div(m, 256), (unsigned)splits), 256, 0, st>>>(
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
    // variant: 0 auto | 1 FMA-pipe kernel | 2 direct form on the tensor cores | 3 / 4: auto with the CTA shape forced
    // to 8 / 4 warps (measurement only)
    SOCKS_CHECK_ARG(variant >= 0 && variant <= 4, 15);
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
    const int shape = variant == 3 ? 1 : (variant == 4 ? 2 : 0);
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
    SOCKS_CHECK_ARG(variant >= 0 && variant <= 4, 17);
    if (m == 0) return SOCKS_OK;
    SOCKS_CHECK_ARG(pts != nullptr && out != nullptr, 9);
    int rc = socks_sr_pack(X, w, vf, ldvf, n, d, scale, divide, T, work, stream);
    if (rc) return rc;
    double* rest = work + round_up(pack_doubles(n, d, T), 2);
    rc = socks_sr_predict_packed(work, X, n, d, scale, divide, pts, m, tubes, problem, T, out, ldout, rest, variant,
                                 stream);
    return rc;
}
