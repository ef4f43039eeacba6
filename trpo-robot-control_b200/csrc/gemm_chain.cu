// gemm_chain.cu -- generic (any NumLayers / any widths) path of the Fisher-vector product and policy gradient.
//
// The per-sample loops of the reference (TRPO_FVP.c:771-924, TRPO_Update.c:254-371) are restructured as a chain of
// batched FP64 GEMMs over a chunk of samples whose activations stay L2-resident:
//   forward  (per layer i)   [Y | RY]_{i+1} = act( [Y_i,1] * [W_i;B_i] ),  RX = RY_i*W_i + [Y_i,1]*[VW_i;VB_i]   (:783-836)
//   backward (per layer i)   G_{i-1} = (G_i * W_{i-1}^T) .* f'(Y_{i-1})                                            (:857-900)
//   outer    (per layer i)   [RGW_{i-1};RGB_{i-1}] += [Y_{i-1},1]^T * G_i   (split-K over sample slices)            (:890-921)
// The flat parameter layout already stores [W_i;B_i] as one (L_i+1) x L_{i+1} row-major matrix, so the bias is the
// "ones" row of the augmented operand and no packing is needed.
//
// All three are one tiled GEMM on the FP64 tensor pipe (DMMA.8x8x4): CTA tile 128 x 64, operands staged by cp.async
// (16-byte when the leading dimension is even, 8-byte otherwise; zero-filled at the ragged edges, so odd sizes such as
// 15, 17, 376 need no padding in HBM) into a double-buffered shared tile whose row strides (20 / 36 / 68 doubles) make
// every fragment read bank-conflict-free. The forward kernel carries two accumulator sets (x and R{x}) and issues the
// three products of the R-op per fragment pair: 512 threads, 32 x 16 warp tiles, k-step 16 (122 registers, 16 warps per
// SM). The single-product kernels (backward, outer) use 256 threads, 32 x 32 warp tiles, k-step 32 and run two CTAs
// per SM (126-128 registers). tanh runs branch-free on 8 values in lock step (dmma_common.cuh).
// Determinism: every output element has exactly one owner thread per (slice); slices are summed in fixed order.
#include <stdint.h>
#include <stdlib.h>

#include "trpo_internal.cuh"
#include "dmma_common.cuh"

namespace {

constexpr int BM = 128, BN = 64, NT = 256;
constexpr int BK_DUAL = 16, BK_SINGLE = 32;             // k-step: the dual (forward R-op) kernel holds 4 operand tiles per stage
constexpr int RSB = 68;                                 // shared row strides: % 16 == 4 -> conflict-free fragment reads
template <int BK> struct Tile {
    static constexpr int RSA = BK + 4;                  // 20 / 36
    static constexpr int A = BM * RSA, B = BK * RSB;    // doubles per operand tile
};

__device__ __forceinline__ double act_apply(char a, double x) {
    switch (a) {
        case 't': return tanh(x);
        case 'o': return 0.1 * x;
        case 's': return 1.0 / (1.0 + exp(-x));
        default:  return x;
    }
}
__device__ __forceinline__ double act_deriv(char a, double y) {   // f'(x) expressed through y = f(x)
    switch (a) {
        case 't': return 1.0 - y * y;
        case 'o': return 0.1;
        case 's': return y * (1.0 - y);
        default:  return 1.0;
    }
}

// ---- tile loaders (all 256 threads; 8-byte cp.async, src-size 0 = zero fill) --------------------------------------
// A tile from row-major activations X[rows x ld]: As[m][k] = X[(m0+m)*ld + k0+k]; column k == ld is the augmented
// "ones" column (value ones_val) when aug is set.
template <int BK, int NTH, int TM = BM>
__device__ __forceinline__ void load_a_rowmajor(double *As, const double *X, int rows, int ld, int m0, int k0,
                                                bool aug, double ones_val, int tid) {
    constexpr int RSA = Tile<BK>::RSA;
    if (X != nullptr && (ld & 1) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {
        // even leading dimension: 16-byte copies (half the LDGSTS instructions); k0 and RSA are even
#pragma unroll
        for (int it = 0; it < TM * BK / 2 / NTH; ++it) {
            const int idx = tid + it * NTH, m = idx / (BK / 2), k = (idx % (BK / 2)) * 2;
            const int gm = m0 + m, gk = k0 + k;
            double *dst = &As[m * RSA + k];
            if (aug && gk == ld && gm < rows) { dst[0] = ones_val; dst[1] = 0.0; }
            else {
                const bool in = gm < rows && gk < ld;      // ld even: gk < ld implies gk + 1 < ld
                cp_async16(dst, in ? &X[(size_t)gm * ld + gk] : X, in ? 16 : 0);
            }
        }
        return;
    }
#pragma unroll
    for (int it = 0; it < TM * BK / NTH; ++it) {
        const int idx = tid + it * NTH, m = idx / BK, k = idx % BK;
        const int gm = m0 + m, gk = k0 + k;
        const bool in = X != nullptr && gm < rows && gk < ld;
        if (aug && gk == ld && gm < rows) As[m * RSA + k] = ones_val;
        else cp_async8(&As[m * RSA + k], in ? &X[(size_t)gm * ld + gk] : X, in ? 8 : 0);
    }
}
// A tile for the outer product: As[m][k] = Yprev[(s0+k)*M0 + m0+m] for m < M0, 1.0 for m == M0 (bias-gradient row)
template <int BK>
__device__ __forceinline__ void load_a_transposed(double *As, const double *Y, int s_end, int M0, int m0, int s0, int tid) {
    // A warp copies an 8 (m) x 4 (k) patch per step: 64-byte global segments along m, and shared addresses
    // m*RSA + k that fall on 16 distinct banks per half warp (lanes along m only would be an 8-way conflict: RSA % 16 == 4).
    constexpr int RSA = Tile<BK>::RSA;
    const int lane = tid & 31, w = tid >> 5, mm = lane & 7, kk = lane >> 3;
#pragma unroll
    for (int it = 0; it < BM * BK / NT; ++it) {
        const int blk = it * 8 + w, m = (blk & 15) * 8 + mm, k = (blk >> 4) * 4 + kk;
        const int gm = m0 + m, gs = s0 + k;
        const bool in = Y != nullptr && gm < M0 && gs < s_end;
        if (gm == M0 && gs < s_end) As[m * RSA + k] = 1.0;
        else cp_async8(&As[m * RSA + k], in ? &Y[(size_t)gs * M0 + gm] : Y, in ? 8 : 0);
    }
}
// The same operand in its NATURAL orientation: An[k][m] = Yprev[(s0+k)*M0 + m0+m] (1.0 for m == M0), row stride RSN.
// No transposition on the way in: 16-byte copies along m (8 per thread and stage instead of 16 element copies), and the
// m8n8k4 A fragment (lane (g,t) holds A[m=g][k=t]) reads An[(4q+t)*RSN + m] -- per half warp t*RSN + g covers 16
// distinct banks because RSN % 16 == 4.
constexpr int RSN = BM + 4;
template <int BK>
__device__ __forceinline__ void load_a_natural(double *An, const double *Y, int s_end, int M0, int m0, int s0, int tid) {
    if (Y != nullptr && (M0 & 1) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0) {
#pragma unroll
        for (int it = 0; it < BK * BM / 2 / NT; ++it) {
            const int idx = tid + it * NT, k = idx >> 6, m = (idx & 63) * 2;
            const int gm = m0 + m, gs = s0 + k;
            if (gm == M0) {                                  // bias-gradient column of the augmented matrix
                An[k * RSN + m] = gs < s_end ? 1.0 : 0.0;
                An[k * RSN + m + 1] = 0.0;
            } else {
                const bool in = gm < M0 && gs < s_end;
                cp_async16(&An[k * RSN + m], in ? &Y[(size_t)gs * M0 + gm] : Y, in ? 16 : 0);
            }
        }
        return;
    }
#pragma unroll 2
    for (int it = 0; it < BK * BM / NT; ++it) {
        const int idx = tid + it * NT, k = idx >> 7, m = idx & 127;
        const int gm = m0 + m, gs = s0 + k;
        const bool in = Y != nullptr && gm < M0 && gs < s_end;
        if (gm == M0) An[k * RSN + m] = gs < s_end ? 1.0 : 0.0;
        else cp_async8(&An[k * RSN + m], in ? &Y[(size_t)gs * M0 + gm] : Y, in ? 8 : 0);
    }
}
// B tile from a row-major matrix M[kdim x N]: Bs[k][n] = M[(k0+k)*N + n0+n]
template <int BK, int NTH>
__device__ __forceinline__ void load_b_rowmajor(double *Bs, const double *M, int kdim, int N, int k0, int n0, int tid) {
    if ((N & 1) == 0 && (reinterpret_cast<uintptr_t>(M) & 15) == 0) {
#pragma unroll
        for (int it = 0; it < BK * BN / 2 / NTH; ++it) {
            const int idx = tid + it * NTH, k = idx >> 5, n = (idx & 31) * 2;
            const int gk = k0 + k, gn = n0 + n;
            const bool in = gk < kdim && gn < N;
            cp_async16(&Bs[k * RSB + n], in ? &M[(size_t)gk * N + gn] : M, in ? 16 : 0);
        }
        return;
    }
#pragma unroll
    for (int it = 0; it < BK * BN / NTH; ++it) {
        const int idx = tid + it * NTH, k = idx >> 6, n = idx & 63;
        const int gk = k0 + k, gn = n0 + n;
        const bool in = gk < kdim && gn < N;
        cp_async8(&Bs[k * RSB + n], in ? &M[(size_t)gk * N + gn] : M, in ? 8 : 0);
    }
}
// B tile from W[N x Kd] row-major used transposed: Bs[k][n] = W[(n0+n)*Kd + k0+k]
template <int BK>
__device__ __forceinline__ void load_b_transposed(double *Bs, const double *W, int Kd, int N, int k0, int n0, int tid) {
    // 8 (k) x 4 (n) patch per warp and step: 64-byte global segments along k, shared addresses k*RSB + n at most 2-way
    // conflicting (lanes along k only: 8-way, RSB % 16 == 4)
    const int lane = tid & 31, w = tid >> 5, kk = lane & 7, nn = lane >> 3;
    constexpr int KB = BK / 8;                              // k-patches per tile row
#pragma unroll
    for (int it = 0; it < BK * BN / NT; ++it) {
        const int blk = it * 8 + w, k = (blk % KB) * 8 + kk, n = (blk / KB) * 4 + nn;
        const int gk = k0 + k, gn = n0 + n;
        const bool in = gk < Kd && gn < N;
        cp_async8(&Bs[k * RSB + n], in ? &W[(size_t)gn * Kd + gk] : W, in ? 8 : 0);
    }
}

// one k-step (16) of a warp's 32 x 32 sub-tile: acc += A*B [, racc += RA*B + A*VB]
template <bool DUAL, bool HAS_RA, int BK, int NJ, bool ANAT = false>
__device__ __forceinline__ void mma_stage(double (&acc)[4][NJ][2], double (&racc)[4][NJ][2], const double *As,
                                          const double *RAs, const double *Bs, const double *VBs, int wm, int wn, int g, int t) {
    constexpr int RSA = Tile<BK>::RSA;
#pragma unroll
    for (int q = 0; q < BK / 4; ++q) {
        double a[4], ra[4], b[NJ], vb[NJ];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = ANAT ? As[(4 * q + t) * RSN + 32 * wm + 8 * i + g] : As[(32 * wm + 8 * i + g) * RSA + 4 * q + t];
            if (DUAL && HAS_RA) ra[i] = RAs[(32 * wm + 8 * i + g) * RSA + 4 * q + t];
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            b[j] = Bs[(4 * q + t) * RSB + 8 * NJ * wn + 8 * j + g];
            if (DUAL) vb[j] = VBs[(4 * q + t) * RSB + 8 * NJ * wn + 8 * j + g];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (DUAL && HAS_RA) dmma(racc[i][j], ra[i], b[j]);
                dmma(acc[i][j], a[i], b[j]);
                if (DUAL) dmma(racc[i][j], a[i], vb[j]);
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// forward: DUAL = also propagate R{} (FVP); HAS_RA = the incoming R{y} is non-zero (false for layer 0).
// The dual (R-op) kernel carries two accumulator sets; to keep 16 warps per SM it runs 512 threads with 32 x 16 warp tiles
// (64 accumulator registers per thread) instead of 256 threads with 32 x 32 tiles (128 registers, 8 warps per SM).
// TM = rows per CTA tile. The dual kernel exists as 128 rows x 512 threads (one CTA per SM) and as 64 rows x 256 threads (two
// CTAs per SM, same 32 x 16 warp tiles): with two independent CTAs the first global loads and the tanh / store epilogue of
// one overlap the main loop of the other, and a block barrier only stops 8 warps.
template <bool DUAL, bool HAS_RA, int TM = BM>
__global__ void __launch_bounds__(DUAL ? (TM == BM ? 512 : 256) : NT, (DUAL && TM == BM) ? 1 : 2) k_chain_fwd(const double *__restrict__ Yin, const double *__restrict__ RYin,
                                                     const double *__restrict__ W, const double *__restrict__ VW,
                                                     int rows, int Kd, int N, char act,
                                                     double *__restrict__ Yout, double *__restrict__ RYout,
                                                     double *__restrict__ Gout, const double *__restrict__ inv_var,
                                                     const int *__restrict__ done) {
    if (done && *done) return;
    extern __shared__ __align__(16) double smem[];
    constexpr int BK = DUAL ? BK_DUAL : BK_SINGLE;
    constexpr int NTH = DUAL ? (TM == BM ? 512 : 256) : NT, WN = DUAL ? 4 : 2, NJ = BN / 8 / WN;     // warps along n, n-tiles per warp
    static_assert(NTH / 32 / WN * 32 == TM, "warp rows must cover the tile");
    constexpr int A_TILE = TM * Tile<BK>::RSA, B_TILE = Tile<BK>::B;
    constexpr int STAGE = A_TILE * (DUAL && HAS_RA ? 2 : 1) + B_TILE * (DUAL ? 2 : 1);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3, wm = w / WN, wn = w % WN;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * BN;     // the n-tiles of one row block run together: its A tile is read from L2 once
    double acc[4][NJ][2] = {}, racc[4][NJ][2] = {};
    __shared__ double exp2_tab[64];
    load_exp2_table(exp2_tab);                          // visible after the first barrier of the k loop
    // augmented contraction length Kd + 1 (bias row) -- unless Kd is a whole number of k-steps: then the bias row would cost
    // a k-step of its own (17 instead of 16 for a 256-wide layer), so the accumulators start from [B ; VB] instead
    const bool bias_init = (Kd % BK) == 0;
    const int nk = bias_init ? Kd / BK : (Kd + 1 + BK - 1) / BK;
    if (bias_init) {
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int gn = n0 + 8 * NJ * wn + 8 * j + 2 * t + r;
                const double bw = gn < N ? W[(size_t)Kd * N + gn] : 0.0;
                const double bv = (DUAL && gn < N) ? VW[(size_t)Kd * N + gn] : 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc[i][j][r] = bw; racc[i][j][r] = bv; }
            }
    }
    auto stage_ptrs = [&](int st, double *&As, double *&RAs, double *&Bs, double *&VBs) {
        double *p = smem + st * STAGE;
        As = p; p += A_TILE;
        RAs = p; if (DUAL && HAS_RA) p += A_TILE;
        Bs = p; p += B_TILE;
        VBs = p;
    };
    auto load = [&](int st, int k0) {
        double *As, *RAs, *Bs, *VBs;
        stage_ptrs(st, As, RAs, Bs, VBs);
        load_a_rowmajor<BK, NTH, TM>(As, Yin, rows, Kd, m0, k0, true, 1.0, tid);
        load_b_rowmajor<BK, NTH>(Bs, W, Kd + 1, N, k0, n0, tid);
        if (DUAL) {
            if (HAS_RA) load_a_rowmajor<BK, NTH, TM>(RAs, RYin, rows, Kd, m0, k0, false, 0.0, tid);
            load_b_rowmajor<BK, NTH>(VBs, VW, Kd + 1, N, k0, n0, tid);
        }
        cp_async_commit();
    };
    // 3-stage cp.async pipeline, one block barrier per k-step: the barrier that publishes stage `it` also guarantees that
    // everybody is done with stage it-1, whose buffer the load issued right after it overwrites
    // interior k-steps of an aligned, full-width tile take a branch-free loader (16-byte copies, no bounds on k, rows past the
    // end zero-filled by src-size 0); only the last k-step(s) need the general one. The general loaders cost 11 % of this
    // kernel's stall samples in index arithmetic and bound / augmented-column branches (profiles/r01_summary.md).
    const bool fast_ok = (Kd & 1) == 0 && (N & 1) == 0 && n0 + BN <= N && (reinterpret_cast<uintptr_t>(Yin) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                         (!DUAL || ((reinterpret_cast<uintptr_t>(VW) & 15) == 0 &&
                                    (!HAS_RA || (reinterpret_cast<uintptr_t>(RYin) & 15) == 0)));
    auto load_fast = [&](int st, int k0) {                  // requires k0 + BK <= Kd
        double *As, *RAs, *Bs, *VBs;
        stage_ptrs(st, As, RAs, Bs, VBs);
        constexpr int RSA = Tile<BK>::RSA, CH = BK / 2;     // 16-byte chunks per tile row
#pragma unroll
        for (int it2 = 0; it2 < TM * CH / NTH; ++it2) {
            const int idx = tid + it2 * NTH, m = idx / CH, k = (idx % CH) * 2, gm = m0 + m;
            const int bytes = gm < rows ? 16 : 0;
            const size_t off = (size_t)(gm < rows ? gm : rows - 1) * Kd + k0 + k;
            cp_async16(&As[m * RSA + k], Yin + off, bytes);
            if (DUAL && HAS_RA) cp_async16(&RAs[m * RSA + k], RYin + off, bytes);
        }
#pragma unroll
        for (int it2 = 0; it2 < BK * (BN / 2) / NTH; ++it2) {
            const int idx = tid + it2 * NTH, k = idx / (BN / 2), n = (idx % (BN / 2)) * 2;
            const size_t off = (size_t)(k0 + k) * N + n0 + n;
            cp_async16(&Bs[k * RSB + n], W + off, 16);
            if (DUAL) cp_async16(&VBs[k * RSB + n], VW + off, 16);
        }
        cp_async_commit();
    };
    auto load_any = [&](int st, int k0) { if (fast_ok && k0 + BK <= Kd) load_fast(st, k0); else load(st, k0); };
    constexpr int NS = DUAL ? 3 : 2;
#pragma unroll
    for (int s0 = 0; s0 < NS - 1; ++s0) { if (s0 < nk) load_any(s0, s0 * BK); else cp_async_commit(); }
    for (int it = 0; it < nk; ++it) {
        cp_async_wait_group<NS - 2>();
        __syncthreads();
        if (it + NS - 1 < nk) load_any((it + NS - 1) % NS, (it + NS - 1) * BK); else cp_async_commit();
        double *As, *RAs, *Bs, *VBs;
        stage_ptrs(it % NS, As, RAs, Bs, VBs);
        mma_stage<DUAL, HAS_RA, BK, NJ>(acc, racc, As, RAs, Bs, VBs, wm, wn, g, t);
    }
    // epilogue: activation, R{y} = R{x} f'(x), last layer: R-gradient seed RG_K = Ry_K / sigma^2 * f' (TRPO_FVP.c:852-882)
#pragma unroll
    for (int i0 = 0; i0 < 4; i0 += 8 / (2 * NJ)) {
        // 8 values per lock-step tanh: (8 / (2 NJ)) row blocks x NJ tiles x 2 columns
        constexpr int RB = 8 / (2 * NJ);
        double xv[8], dv[8];
#pragma unroll
        for (int ii = 0; ii < RB; ++ii)
#pragma unroll
            for (int j = 0; j < NJ; ++j) { xv[(ii * NJ + j) * 2] = acc[i0 + ii][j][0]; xv[(ii * NJ + j) * 2 + 1] = acc[i0 + ii][j][1]; }
        if (act == 't') tanh_vec<8>(xv, dv, exp2_tab);
        else {
#pragma unroll
            for (int e = 0; e < 8; ++e) { xv[e] = act_apply(act, xv[e]); dv[e] = act_deriv(act, xv[e]); }
        }
#pragma unroll
        for (int ii = 0; ii < RB; ++ii) {
            const int i = i0 + ii, gm = m0 + 32 * wm + 8 * i + g;
            if (gm >= rows) continue;
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int gn = n0 + 8 * NJ * wn + 8 * j + 2 * t + r;
                    if (gn >= N) continue;
                    const double y = xv[(ii * NJ + j) * 2 + r], d = dv[(ii * NJ + j) * 2 + r];
                    if (Yout) Yout[(size_t)gm * N + gn] = y;
                    if (DUAL) {
                        const double ry = racc[i][j][r] * d;
                        if (RYout) RYout[(size_t)gm * N + gn] = ry;
                        if (Gout) Gout[(size_t)gm * N + gn] = ry * inv_var[gn] * d;
                    }
                }
        }
    }
}

// backward: Gout[s][n] = f'(Yprev[s][n]) * sum_k Gin[s][k] * W[n][k]       (W is [N x Kd] row-major)
__global__ void __launch_bounds__(NT, 2) k_chain_bwd(const double *__restrict__ Gin, const double *__restrict__ W,
                                                     const double *__restrict__ Yprev, int rows, int Kd, int N, char act_prev,
                                                     double *__restrict__ Gout, const int *__restrict__ done) {
    if (done && *done) return;
    extern __shared__ __align__(16) double smem[];
    constexpr int BK = BK_SINGLE, A_TILE = Tile<BK>::A, B_TILE = Tile<BK>::B;
    constexpr int STAGE = A_TILE + B_TILE;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3, wm = w >> 1, wn = w & 1;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;     // the n-tiles of one row block run together: its A tile is read from L2 once
    double acc[4][4][2] = {}, dummy[4][4][2];
    const int nk = (Kd + BK - 1) / BK;
    auto load = [&](int st, int k0) {
        double *As = smem + st * STAGE, *Bs = As + A_TILE;
        load_a_rowmajor<BK, NT>(As, Gin, rows, Kd, m0, k0, false, 0.0, tid);
        load_b_transposed<BK>(Bs, W, Kd, N, k0, n0, tid);
        cp_async_commit();
    };
    load(0, 0);
    for (int it = 0; it < nk; ++it) {          // one barrier per k-step: it publishes stage `it` and frees the other buffer
        cp_async_wait_group<0>();
        __syncthreads();
        if (it + 1 < nk) load((it + 1) & 1, (it + 1) * BK);
        const double *As = smem + (it & 1) * STAGE, *Bs = As + A_TILE;
        mma_stage<false, false, BK, 4>(acc, dummy, As, nullptr, Bs, nullptr, wm, wn, g, t);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + 32 * wm + 8 * i + g;
        if (gm >= rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int gn = n0 + 32 * wn + 8 * j + 2 * t + r;
                if (gn >= N) continue;
                Gout[(size_t)gm * N + gn] = acc[i][j][r] * act_deriv(act_prev, Yprev[(size_t)gm * N + gn]);
            }
    }
}

// outer: out[slice][m*N + n] (+)= sum_{s in slice} [Yprev,1][s][m] * G[s][n],  m in [0, M0]  (row M0 = bias gradient)
// bias_colsum: the grid's m-tiles cover rows [0, M0) only (M0 a multiple of the 32-row warp tile: a 257th row would cost a
// whole extra 128-row CTA tile, a 65th one a third warp row) and the bias-gradient row = column sums of G is formed by the m-tile-0 CTAs from the B tiles they
// stage anyway (fixed order: 4 k-phases per column, then the phases).
__global__ void __launch_bounds__(NT, 2) k_chain_outer(const double *__restrict__ Yprev, const double *__restrict__ G,
                                                       int rows, int M0, int N, int per_slice, int tiles_n,
                                                       double *__restrict__ partial, int P, int out_off, int accumulate,
                                                       int bias_colsum, const int *__restrict__ done) {
    if (done && *done) return;
    extern __shared__ __align__(16) double smem[];
    constexpr int BK = BK_SINGLE, A_TILE = BK * RSN, B_TILE = Tile<BK>::B;       // A in its natural orientation
    constexpr int STAGE = A_TILE + B_TILE;
    static_assert(A_TILE <= Tile<BK>::A && (A_TILE & 1) == 0, "natural A tile must fit the single-product stage and keep B 16-byte aligned");
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3, wm = w >> 1, wn = w & 1;
    const int m0 = (blockIdx.x / tiles_n) * BM, n0 = (blockIdx.x % tiles_n) * BN;
    const int slice = blockIdx.y;
    const int s0 = slice * per_slice;
    const int s1 = min(rows, s0 + per_slice);
    double acc[4][4][2] = {}, dummy[4][4][2];
    const bool colsum = bias_colsum && m0 == 0;
    double bsum = 0.0;
    // warps whose 32 x 32 sub-tile lies entirely outside the matrix (64-wide layers in a 128-row tile, a 17-column action
    // layer in a 64-column tile) issue no DMMAs; they still help with the copies
    const bool active = m0 + 32 * wm < M0 + (bias_colsum ? 0 : 1) && n0 + 32 * wn < N;
    const int nk = s1 > s0 ? (s1 - s0 + BK - 1) / BK : 0;
    auto load = [&](int st, int ks) {
        double *As = smem + st * STAGE, *Bs = As + A_TILE;
        load_a_natural<BK>(As, Yprev, s1, M0, m0, ks, tid);
        load_b_rowmajor<BK, NT>(Bs, G + (size_t)ks * N, s1 - ks, N, 0, n0, tid);
        cp_async_commit();
    };
    if (nk) load(0, s0);
    for (int it = 0; it < nk; ++it) {
        cp_async_wait_group<0>();
        __syncthreads();
        if (it + 1 < nk) load((it + 1) & 1, s0 + (it + 1) * BK);
        const double *As = smem + (it & 1) * STAGE, *Bs = As + A_TILE;
        if (active) mma_stage<false, false, BK, 4, true>(acc, dummy, As, nullptr, Bs, nullptr, wm, wn, g, t);
        if (colsum) {
#pragma unroll
            for (int kk = 0; kk < BK / 4; ++kk) bsum += Bs[(4 * kk + (tid >> 6)) * RSB + (tid & 63)];
        }
    }
    double *out = partial + (size_t)slice * P + out_off;
    if (colsum) {
        __syncthreads();                                   // everybody is done with the stage buffers
        smem[tid] = bsum;
        __syncthreads();
        const int gn = n0 + tid;
        if (tid < BN && gn < N) {
            const double sum = ((smem[tid] + smem[tid + 64]) + smem[tid + 128]) + smem[tid + 192];
            const size_t o = (size_t)M0 * N + gn;
            out[o] = accumulate ? out[o] + sum : sum;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + 32 * wm + 8 * i + g;
        if (gm >= M0 + (bias_colsum ? 0 : 1)) continue;      // with bias_colsum the row M0 belongs to the column sums above
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int gn = n0 + 32 * wn + 8 * j + 2 * t + r;
                if (gn >= N) continue;
                const size_t o = (size_t)gm * N + gn;
                out[o] = accumulate ? out[o] + acc[i][j][r] : acc[i][j][r];
            }
    }
}


inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------------------------
// tail: the LAST weight layer of the FVP when the action dimension is small (A <= 24) and the output activation is
// linear / 0.1x (its pre-activation does not enter the FVP). One kernel does what forward(K-1) + backward(K) would do in two
// 128 x 64-tile GEMM launches that are 4x too wide for a 17-column layer:
//   Rx_K = Ry_{K-1} W + [y_{K-1},1] [VW;VB]      (TRPO_FVP.c:795-803)       k loop over H+1, 128 rows x 8*NTA columns per CTA
//   RG_K = Rx_K f'^2 / sigma^2                   (:809-823, :852-854, :869-882)
//   RG_{K-1} = (RG_K W^T) .* f'(y_{K-1})         (:890-899)                 contraction over the 8*NTA padded actions
// A warp owns 16 rows; RG_K stays in its accumulator registers and is used directly as the A fragment of the second
// product (contraction-index permutation, as in fvp_fused.cu), with W^T staged once per CTA in shared memory at a row
// stride of HP + 2 doubles (2*stride % 16 == 4: conflict-free fragment reads).
// TMT rows per CTA, 16 per warp. 64-row CTAs of 4 warps (two per SM, 104 KB of shared memory each at H = 256) let one CTA's
// W^T staging, first loads and store epilogue overlap the other's DMMA phases; the 128-row shape (one CTA per SM) ran the phases
// back to back: 35 % DMMA-pipe activity, long-scoreboard stalls on top (profiles/r02_summary.md).
template <int NTA, int TMT>
__global__ void __launch_bounds__(2 * TMT, TMT == 64 ? 2 : 1) k_chain_tail(const double *__restrict__ Y, const double *__restrict__ RY,
                                                      const double *__restrict__ W, const double *__restrict__ VW,
                                                      int rows, int H, int A, char act_prev, double d3,
                                                      const double *__restrict__ inv_var,
                                                      double *__restrict__ GK, double *__restrict__ Gprev,
                                                      const int *__restrict__ done) {
    if (done && *done) return;
    extern __shared__ __align__(16) double smem[];
    constexpr int BK = 16, RSA = Tile<BK>::RSA, AP = 8 * NTA, RSBT = AP + 4;
    constexpr int NTT = 2 * TMT;
    constexpr int A_TILE = TMT * RSA, B_TILE = BK * RSBT, STAGE = 2 * A_TILE + 2 * B_TILE;
    const int HP = (H + 7) & ~7, RST = HP + 2;
    constexpr int NS = 3;                                       // cp.async ring: two k-steps of look-ahead cover a DRAM round trip
    double *WT = smem;                                          // [AP][RST], staged over the ring once the forward loop is done
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.x * TMT;

    double rxa[2][NTA][2] = {}, rxb[2][NTA][2] = {};
    const int nk = (H + 1 + BK - 1) / BK;
    auto load = [&](int st, int k0) {
        double *As = smem + st * STAGE, *RAs = As + A_TILE, *Bs = RAs + A_TILE, *VBs = Bs + B_TILE;
        load_a_rowmajor<BK, NTT, TMT>(As, Y, rows, H, m0, k0, true, 1.0, tid);
        load_a_rowmajor<BK, NTT, TMT>(RAs, RY, rows, H, m0, k0, false, 0.0, tid);
        for (int idx = tid; idx < BK * AP; idx += NTT) {
            const int k = idx / AP, n = idx % AP, gk = k0 + k;
            const bool in = gk <= H && n < A;
            cp_async8(&Bs[k * RSBT + n], in ? &W[(size_t)gk * A + n] : W, in ? 8 : 0);
            cp_async8(&VBs[k * RSBT + n], in ? &VW[(size_t)gk * A + n] : VW, in ? 8 : 0);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s0 = 0; s0 < NS - 1; ++s0) { if (s0 < nk) load(s0, s0 * BK); else cp_async_commit(); }
    for (int it = 0; it < nk; ++it) {
        cp_async_wait_group<NS - 2>();
        __syncthreads();                                        // publishes stage `it`, frees the buffer of step it - 1
        if (it + NS - 1 < nk) load((it + NS - 1) % NS, (it + NS - 1) * BK); else cp_async_commit();
        const double *As = smem + (it % NS) * STAGE, *RAs = As + A_TILE, *Bs = RAs + A_TILE, *VBs = Bs + B_TILE;
#pragma unroll
        for (int q = 0; q < BK / 4; ++q) {
            double a[2], ra[2], b[NTA], vb[NTA];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                a[i] = As[(16 * w + 8 * i + g) * RSA + 4 * q + t];
                ra[i] = RAs[(16 * w + 8 * i + g) * RSA + 4 * q + t];
            }
#pragma unroll
            for (int j = 0; j < NTA; ++j) { b[j] = Bs[(4 * q + t) * RSBT + 8 * j + g]; vb[j] = VBs[(4 * q + t) * RSBT + 8 * j + g]; }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < NTA; ++j) { dmma(rxa[i][j], ra[i], b[j]); dmma(rxb[i][j], a[i], vb[j]); }
        }
    }
    cp_async_wait_group<0>();
    __syncthreads();                                            // the ring is idle: W^T takes its place
    // W^T, zero padded: WT[k][j] = W[j][k]
    for (int idx = tid; idx < AP * RST; idx += NTT) WT[idx] = 0.0;
    __syncthreads();
    for (int idx = tid; idx < H * A; idx += NTT) { const int j = idx / A, k = idx % A; WT[k * RST + j] = W[idx]; }
    __syncthreads();
    // RG_K in accumulator layout; rows past the end of the chunk contribute nothing
    double gk[2][NTA][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int gm = m0 + 16 * w + 8 * i + g;
#pragma unroll
        for (int j = 0; j < NTA; ++j)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int col = 8 * j + 2 * t + r;
                const double v = (gm < rows && col < A) ? (rxa[i][j][r] + rxb[i][j][r]) * d3 * inv_var[col] * d3 : 0.0;
                gk[i][j][r] = v;
                if (gm < rows && col < A) GK[(size_t)gm * A + col] = v;
            }
    }
    // RG_{K-1} = (RG_K W^T) .* f'(y_{K-1}), 32 columns at a time. The y_{K-1} values the epilogue multiplies with are fetched
    // BEFORE the block's DMMAs (16 registers of double2): issued after them, every f' waited a full global-load latency
    // (long-scoreboard stalls on the 1 - y^2 DFMAs were the kernel's top stall, profiles/r02_summary.md).
    const bool pair_ok = (H & 1) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0 && (reinterpret_cast<uintptr_t>(Gprev) & 15) == 0;
    for (int n0 = 0; n0 < HP; n0 += 32) {
        double2 yv[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int gm = m0 + 16 * w + 8 * i + g;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = n0 + 8 * j + 2 * t;
                yv[i][j] = make_double2(0.0, 0.0);
                if (gm < rows && col < H) {
                    const size_t o = (size_t)gm * H + col;
                    if (pair_ok) yv[i][j] = *reinterpret_cast<const double2 *>(&Y[o]);
                    else { yv[i][j].x = Y[o]; if (col + 1 < H) yv[i][j].y = Y[o + 1]; }
                }
            }
        }
        double acc[2][4][2] = {};
#pragma unroll
        for (int b = 0; b < NTA; ++b)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const double *wrow = WT + (8 * b + 2 * t + r) * RST + n0 + g;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double bf = (n0 + 8 * j < HP) ? wrow[8 * j] : 0.0;
#pragma unroll
                    for (int i = 0; i < 2; ++i) dmma(acc[i][j], gk[i][b][r], bf);
                }
            }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int gm = m0 + 16 * w + 8 * i + g;
            if (gm >= rows) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = n0 + 8 * j + 2 * t;
                if (col >= H) continue;
                const size_t o = (size_t)gm * H + col;
                const double g0 = acc[i][j][0] * act_deriv(act_prev, yv[i][j].x), g1 = acc[i][j][1] * act_deriv(act_prev, yv[i][j].y);
                if (pair_ok) *reinterpret_cast<double2 *>(&Gprev[o]) = make_double2(g0, g1);
                else { Gprev[o] = g0; if (col + 1 < H) Gprev[o + 1] = g1; }
            }
        }
    }
}

template <int NTA, int TMT = BM>
size_t tail_smem_bytes(int H) {
    constexpr int BK = 16, AP = 8 * NTA;
    const int HP = (H + 7) & ~7;
    const size_t ring = 3 * (size_t)(2 * TMT * Tile<BK>::RSA + 2 * BK * (AP + 4)), wt = (size_t)AP * (HP + 2);     // W^T reuses the ring
    return sizeof(double) * (ring > wt ? ring : wt);
}

template <int NTA>
int launch_tail(const double *Y, const double *RY, const double *W, const double *VW, int rows, int H, int A, char act_prev,
                double d3, const double *inv_var, double *GK, double *Gprev, const int *done, cudaStream_t st) {
    // two 64-row CTAs per SM when both fit (113 KB each); TRPO_CHAIN_TAIL_TM=128 keeps the one-CTA shape
    static const bool want_half = !(getenv("TRPO_CHAIN_TAIL_TM") && atoi(getenv("TRPO_CHAIN_TAIL_TM")) == 128);
    if (want_half && tail_smem_bytes<NTA, 64>(H) <= 113 * 1024) {
        const size_t bytes = tail_smem_bytes<NTA, 64>(H);
        if (cudaFuncSetAttribute(k_chain_tail<NTA, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return -1;
        k_chain_tail<NTA, 64><<<cdiv(rows, 64), 128, bytes, st>>>(Y, RY, W, VW, rows, H, A, act_prev, d3, inv_var, GK, Gprev, done);
        return 0;
    }
    const size_t bytes = tail_smem_bytes<NTA>(H);
    // size depends on the layer width and the attribute is per device: set it on every launch (host-side only, ~1 us)
    if (cudaFuncSetAttribute(k_chain_tail<NTA, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return -1;
    k_chain_tail<NTA, BM><<<cdiv(rows, BM), NT, bytes, st>>>(Y, RY, W, VW, rows, H, A, act_prev, d3, inv_var, GK, Gprev, done);
    return 0;
}

constexpr size_t SMEM_FWD_DUAL = sizeof(double) * 3 * (2 * Tile<BK_DUAL>::A + 2 * Tile<BK_DUAL>::B);
constexpr size_t SMEM_FWD_L0   = sizeof(double) * 3 * (Tile<BK_DUAL>::A + 2 * Tile<BK_DUAL>::B);
constexpr size_t SMEM_SINGLE   = sizeof(double) * 2 * (Tile<BK_SINGLE>::A + Tile<BK_SINGLE>::B);
constexpr int TM_HALF = 64;
constexpr size_t SMEM_FWD_DUAL_H = sizeof(double) * 3 * (2 * TM_HALF * Tile<BK_DUAL>::RSA + 2 * Tile<BK_DUAL>::B);
constexpr size_t SMEM_FWD_L0_H   = sizeof(double) * 3 * (TM_HALF * Tile<BK_DUAL>::RSA + 2 * Tile<BK_DUAL>::B);

bool configure_kernels() {
    static DeviceOnce once;
    if (!once.pending()) return true;
    bool r = true;
    r = r && cudaFuncSetAttribute(k_chain_fwd<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FWD_DUAL) == cudaSuccess;
    r = r && cudaFuncSetAttribute(k_chain_fwd<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FWD_L0) == cudaSuccess;
    r = r && cudaFuncSetAttribute(k_chain_fwd<true, true, TM_HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FWD_DUAL_H) == cudaSuccess;
    r = r && cudaFuncSetAttribute(k_chain_fwd<true, false, TM_HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FWD_L0_H) == cudaSuccess;
    r = r && cudaFuncSetAttribute(k_chain_fwd<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_SINGLE) == cudaSuccess;
    r = r && cudaFuncSetAttribute(k_chain_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_SINGLE) == cudaSuccess;
    r = r && cudaFuncSetAttribute(k_chain_outer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_SINGLE) == cudaSuccess;
    if (r) once.mark();
    return r;
}

// policy-gradient seed (TRPO_Update.c:297-301) followed by f'(y_K) (:310-324): G[s][j] and GL[s][j]
__global__ void k_pg_seed(const double *__restrict__ mean, const double *__restrict__ action, const double *__restrict__ adv,
                          const double *__restrict__ logstd, const double *__restrict__ YK, char actK,
                          int rows, int A, double *__restrict__ G, double *__restrict__ GL) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * A) return;
    const int s = idx / A, j = idx % A;
    const double sd = exp(logstd[j]);
    const double t = (action[idx] - mean[idx]) / sd;
    double g = adv[s] * t / sd;
    g *= act_deriv(actK, YK[idx]);
    G[idx] = g;
    GL[idx] = adv[s] * (t * t - 1.0);
}


// split-K slices of one outer-product GEMM: about two CTAs per SM in total, never more than the partial rows available.
// Wide layers get few slices (each CTA then amortises its 128 x 64 partial-tile read-modify-write over many samples);
// rows of the partial buffer a layer does not use stay zero from the allocation-time memset.
inline int layer_slices(int tiles, int max_slices) {
    int ns = 296 / tiles;                                // floor: whole waves of 148 CTAs
    if (ns > max_slices) ns = max_slices;
    return ns < 1 ? 1 : ns;
}

}  // namespace

size_t chain_scratch_bytes(const NetDesc &net, int chunk, int nslices) {
    size_t maxL = 0, sumL = 0;
    for (int i = 1; i <= net.K; ++i) { sumL += net.L[i]; if ((size_t)net.L[i] > maxL) maxL = net.L[i]; }
    // + the row-permuted weight / direction copies of the TMA-fed forward kernel (2 doubles of slack keep them 16-byte aligned)
    return sizeof(double) * ((size_t)chunk * (sumL + 4 * maxL) + (size_t)nslices * net.P + 2 + 2 * chain_tma_perm_offset(net, net.K) +
                             chain_tma_tail_doubles(net.K >= 1 ? net.L[net.K - 1] : 0));
}

int chain_accumulate(const NetDesc &net, const ChainScratch &sc, ChainMode mode,
                     const double *d_theta, const double *d_v, const double *d_inv_var,
                     const double *d_obs, const double *d_mean, const double *d_action, const double *d_adv,
                     size_t nsamples, double *d_zsum, const int *d_done, const P2PComm *p2p, cudaStream_t st,
                     long long *launches) {
    const int K = net.K, A = net.L[K];
    const bool fvp = (mode == CHAIN_FVP);
    if (!configure_kernels()) return -1;
    int chunk_idx = 0;
    // While the batch is still crossing PCIe (first FVP after a piecewise set_batch) the loop walks it piece by piece: the copy
    // is the slower side (3 GB at 55 GB/s = 54 ms against 46 ms of arithmetic at Humanoid size), so the FVP ends one chunk's
    // worth of arithmetic after the last byte lands -- with 300 k-row chunks that tail was 14 + 4 ms, with 89 k-row pieces 1 - 4 ms.
    // Same sums, grouped differently: the streamed FVP agrees with a resident one to rounding (1e-15), not bitwise.
    const size_t step = (sc.piece_events && sc.piece_rows && sc.piece_rows < (size_t)sc.chunk) ? sc.piece_rows : (size_t)sc.chunk;
    // TMA-fed forward layers read the weights and the direction from row-permuted copies (gemm_chain_tma.cu), rebuilt per FVP
    bool perm_ready[TRPO_MAX_LAYERS] = {}, tail_ready = false;
    for (size_t c0 = 0; c0 < nsamples; c0 += step, ++chunk_idx) {
        const int rows = (int)((nsamples - c0 < step) ? nsamples - c0 : step);
        const int accumulate = chunk_idx > 0;
        if (sc.piece_events && sc.piece_rows) {          // host-to-device copy still in flight: wait for this chunk's rows only
            for (size_t pi = c0 / sc.piece_rows; pi * sc.piece_rows < c0 + rows && pi < (size_t)sc.n_pieces; ++pi)
                if (cudaStreamWaitEvent(st, sc.piece_events[pi], 0) != cudaSuccess) return -1;
        }
        // the last layer of a narrow-action FVP goes through the fused tail kernel (forward(K-1) + seed + backward(K))
        const bool tail = fvp && K >= 2 && A <= 24 && (net.ac[K] == 'l' || net.ac[K] == 'o') &&
                          tail_smem_bytes<3>(net.L[K - 1]) <= 200 * 1024;
        int ldgK = A;                                          // row stride of G_K (padded to even by the TMA-fed tail when A >= 16 is odd)
        // ---- forward ----
        for (int i = 0; i < (tail ? K - 1 : K); ++i) {
            const double *Yin = (i == 0) ? d_obs + c0 * net.L[0] : sc.Y[i];
            const bool last = (i == K - 1);
            dim3 grid(cdiv(net.L[i + 1], BN), cdiv(rows, BM));
            if (fvp) {
                const double *RYin = (i == 0) ? nullptr : sc.RY[i & 1];
                const bool needY = !last || net.ac[K] == 't' || net.ac[K] == 's';
                // two 64-row CTAs per SM by default; TRPO_CHAIN_FWD_TM=128 selects the one-CTA-per-SM 128-row variant
                static const bool half_tiles = !(getenv("TRPO_CHAIN_FWD_TM") && atoi(getenv("TRPO_CHAIN_FWD_TM")) == 128);
                dim3 grid_h(cdiv(net.L[i + 1], BN), cdiv(rows, TM_HALF));
                double *Yo = needY ? sc.Y[i + 1] : nullptr, *RYo = last ? nullptr : sc.RY[(i + 1) & 1], *Go = last ? sc.G[K & 1] : nullptr;
                const double *Wl = d_theta + net.w_off[i], *VWl = d_v + net.w_off[i];
                if (sc.wperm && chain_tma_fwd_eligible(Yin, RYin, Yo, RYo, Go, net.L[i], net.L[i + 1])) {
                    double *Wp = sc.wperm + chain_tma_perm_offset(net, i), *Vp = sc.vperm + chain_tma_perm_offset(net, i);
                    if (!perm_ready[i]) {
                        chain_tma_permute_rows(Wl, Wp, net.L[i], net.L[i + 1], st);
                        chain_tma_permute_rows(VWl, Vp, net.L[i], net.L[i + 1], st);
                        *launches += 2;
                        perm_ready[i] = true;
                    }
                    if (chain_tma_fwd(Yin, RYin, Wp, Vp, Wl, VWl, rows, net.L[i], net.L[i + 1], net.ac[i + 1], Yo, RYo, Go, d_inv_var, d_done, st))
                        return -1;
                } else
                if (i == 0 && half_tiles)
                    k_chain_fwd<true, false, TM_HALF><<<grid_h, 256, SMEM_FWD_L0_H, st>>>(Yin, nullptr, Wl, VWl, rows, net.L[i], net.L[i + 1],
                                                                                     net.ac[i + 1], Yo, RYo, Go, d_inv_var, d_done);
                else if (i == 0)
                    k_chain_fwd<true, false><<<grid, 512, SMEM_FWD_L0, st>>>(Yin, nullptr, Wl, VWl, rows, net.L[i], net.L[i + 1],
                                                                              net.ac[i + 1], Yo, RYo, Go, d_inv_var, d_done);
                else if (half_tiles)
                    k_chain_fwd<true, true, TM_HALF><<<grid_h, 256, SMEM_FWD_DUAL_H, st>>>(Yin, RYin, Wl, VWl, rows, net.L[i], net.L[i + 1],
                                                                                      net.ac[i + 1], Yo, RYo, Go, d_inv_var, d_done);
                else
                    k_chain_fwd<true, true><<<grid, 512, SMEM_FWD_DUAL, st>>>(Yin, RYin, Wl, VWl, rows, net.L[i], net.L[i + 1],
                                                                               net.ac[i + 1], Yo, RYo, Go, d_inv_var, d_done);
            } else {
                k_chain_fwd<false, false><<<grid, NT, SMEM_SINGLE, st>>>(Yin, nullptr, d_theta + net.w_off[i], nullptr, rows,
                                                       net.L[i], net.L[i + 1], net.ac[i + 1], sc.Y[i + 1], nullptr,
                                                       nullptr, nullptr, d_done);
            }
            ++*launches;
        }
        if (!fvp) {
            // seed G_K and the per-sample LogStd gradient (kept in RY[0] as a [rows x A] matrix)
            const int n = rows * A;
            k_pg_seed<<<cdiv(n, 256), 256, 0, st>>>(d_mean + c0 * A, d_action + c0 * A, d_adv + c0, d_theta + net.logstd_off,
                                                    sc.Y[K], net.ac[K], rows, A, sc.G[K & 1], sc.RY[0]);
            ++*launches;
            const int tl = cdiv(1, BM) * cdiv(A, BN), nsl = layer_slices(tl, sc.nslices);
            dim3 g1(tl, nsl);
            k_chain_outer<<<g1, NT, SMEM_SINGLE, st>>>(nullptr, sc.RY[0], rows, 0, A, cdiv(cdiv(rows, nsl), BK_SINGLE) * BK_SINGLE, cdiv(A, BN),
                                             sc.partial, net.P, net.logstd_off, accumulate, 0, d_done);
            ++*launches;
        }
        if (tail) {
            const int H = net.L[K - 1];
            const double d3 = net.ac[K] == 'o' ? 0.1 : 1.0;
            const double *W = d_theta + net.w_off[K - 1], *VW = d_v + net.w_off[K - 1];
            int rc;
            if (sc.tailw && chain_tma_tail_eligible(sc.Y[K - 1], sc.RY[(K - 1) & 1], sc.G[(K - 1) & 1], H, A)) {
                if (!tail_ready) { chain_tma_tail_prepare(W, VW, sc.tailw, H, A, st); *launches += 3; tail_ready = true; }
                if (A >= 16) ldgK = (A + 1) & ~1;               // even row stride: RG_K becomes a TMA operand of the layer's outer product
                rc = chain_tma_tail(sc.Y[K - 1], sc.RY[(K - 1) & 1], VW, sc.tailw, rows, H, A, net.ac[K - 1], d3, d_inv_var, sc.G[K & 1],
                                    ldgK, sc.G[(K - 1) & 1], d_done, st);
            } else
            if (A <= 8) rc = launch_tail<1>(sc.Y[K - 1], sc.RY[(K - 1) & 1], W, VW, rows, H, A, net.ac[K - 1], d3, d_inv_var, sc.G[K & 1], sc.G[(K - 1) & 1], d_done, st);
            else if (A <= 16) rc = launch_tail<2>(sc.Y[K - 1], sc.RY[(K - 1) & 1], W, VW, rows, H, A, net.ac[K - 1], d3, d_inv_var, sc.G[K & 1], sc.G[(K - 1) & 1], d_done, st);
            else rc = launch_tail<3>(sc.Y[K - 1], sc.RY[(K - 1) & 1], W, VW, rows, H, A, net.ac[K - 1], d3, d_inv_var, sc.G[K & 1], sc.G[(K - 1) & 1], d_done, st);
            if (rc) return -1;
            ++*launches;
        }
        // ---- backward + outer products ----
        for (int i = K; i >= 1; --i) {
            const double *Yprev = (i == 1) ? d_obs + c0 * net.L[0] : sc.Y[i - 1];
            const int M0 = net.L[i - 1], N = net.L[i];
            // the bias row would open a 32-row warp tile (or a whole 128-row CTA tile) of its own: column sums instead
            const int bias_colsum = (M0 % 32) == 0;
            const int tiles_n = cdiv(N, BN);
            const int ldg = (i == K) ? ldgK : N;               // row stride of G_i
            if (chain_tma_outer_eligible(Yprev, sc.G[i & 1], M0, N, ldg)) {
                // operands by TMA (gemm_chain_tma.cu); the bias-gradient row always comes from the column sums there
                const int ns = layer_slices(chain_tma_tiles_m(M0) * tiles_n, sc.nslices);
                if (chain_tma_outer(Yprev, sc.G[i & 1], ldg, rows, M0, N, cdiv(cdiv(rows, ns), BK_SINGLE) * BK_SINGLE, tiles_n, ns, sc.partial,
                                    net.P, net.w_off[i - 1], accumulate, d_done, st)) return -1;
            } else {
                const int tiles_m = bias_colsum ? cdiv(M0, BM) : cdiv(M0 + 1, BM);
                const int ns = layer_slices(tiles_m * tiles_n, sc.nslices);
                dim3 go(tiles_m * tiles_n, ns);
                k_chain_outer<<<go, NT, SMEM_SINGLE, st>>>(Yprev, sc.G[i & 1], rows, M0, N, cdiv(cdiv(rows, ns), BK_SINGLE) * BK_SINGLE, tiles_n,
                                                 sc.partial, net.P, net.w_off[i - 1], accumulate, bias_colsum, d_done);
            }
            ++*launches;
            if (i > 1 && !(tail && i == K)) {
                const double *Wb = d_theta + net.w_off[i - 1];
                if (chain_tma_bwd_eligible(sc.G[i & 1], Wb, sc.Y[i - 1], sc.G[(i - 1) & 1], N, M0)) {
                    if (chain_tma_bwd(sc.G[i & 1], Wb, sc.Y[i - 1], rows, N, M0, net.ac[i - 1], sc.G[(i - 1) & 1], d_done, st)) return -1;
                } else {
                    dim3 gb(cdiv(M0, BN), cdiv(rows, BM));
                    k_chain_bwd<<<gb, NT, SMEM_SINGLE, st>>>(sc.G[i & 1], Wb, sc.Y[i - 1], rows, N, M0,
                                                   net.ac[i - 1], sc.G[(i - 1) & 1], d_done);
                }
                ++*launches;
            }
        }
    }
    launch_reduce_partials(sc.partial, sc.nslices, net.P, d_zsum, d_done, p2p, st, launches);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int chain_forward(const NetDesc &net, const ChainScratch &sc, const double *d_theta, const double *d_obs,
                  size_t nsamples, double *d_mean_out, cudaStream_t st, long long *launches) {
    const int K = net.K, A = net.L[K];
    if (!configure_kernels()) return -1;
    for (size_t c0 = 0; c0 < nsamples; c0 += sc.chunk) {
        const int rows = (int)((nsamples - c0 < (size_t)sc.chunk) ? nsamples - c0 : sc.chunk);
        for (int i = 0; i < K; ++i) {
            const double *Yin = (i == 0) ? d_obs + c0 * net.L[0] : sc.Y[i];
            double *Yout = (i == K - 1) ? d_mean_out + c0 * A : sc.Y[i + 1];
            dim3 grid(cdiv(net.L[i + 1], BN), cdiv(rows, BM));
            k_chain_fwd<false, false><<<grid, NT, SMEM_SINGLE, st>>>(Yin, nullptr, d_theta + net.w_off[i], nullptr, rows, net.L[i],
                                                   net.L[i + 1], net.ac[i + 1], Yout, nullptr, nullptr, nullptr, nullptr);
            ++*launches;
        }
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
