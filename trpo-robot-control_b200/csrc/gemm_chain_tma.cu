// gemm_chain_tma.cu -- TMA-fed variants of the FP64 GEMM-chain kernels (sm_100a): forward R-op layer (k_fwd_tma), backward GEMM
// (k_bwd_tma), split-K outer product (k_outer_tma) and the fused tail of a narrow action layer (k_tail_tma). Same arithmetic as
// the cp.async kernels of gemm_chain.cu (TRPO_FVP.c:771-924 batched over the samples), which remain for odd widths.
//
// The cp.async loaders of gemm_chain.cu cost the DMMA main loop 8 - 10 % on their own (12 LDGSTS per thread and k-step plus their
// index arithmetic; tools/dmma_gemm_loop.cu: fragments + DMMA 37.1 TFLOP/s, + block barrier 36.5, + cp.async double buffer 33.0,
// tiles by TMA instead 35.4). Here one thread issues a handful of cp.async.bulk.tensor.2d copies per k-step, the tiles land in
// SWIZZLE_128B boxes of 16 FP64 columns (128-byte rows) and complete on an mbarrier; out-of-range rows / columns are zero-filled by
// the copy engine, so the kernels carry no bounds logic in the main loop.
//
// Fragment reads go through the swizzle, byte (row, col) of a box = row*128 + (((col >> 1) ^ (row & 7)) << 4) + (col & 1)*8, with
// the contraction index PERMUTED so that the 16 lanes of a half warp hit 16 distinct 8-byte bank pairs (both operands of a DMMA use
// the same permutation, the sum over k does not care; tests/test_host_logic.py restates the addressing lane by lane):
//   operand whose box ROWS are k (outer product: Y[s][m], G[s][n]):  lane t of step q takes row 8*(q/2) + 2t + (q&1)
//   operand whose box COLUMNS are k (backward: G[s][k], W[n][k]; forward / tail activations): lane t of step q takes column
//   16*(q/4) + 2*(q&3) + 8*(t/2) + (t&1); the forward weights are then read from a copy with rows permuted inside groups of 16
// Results agree with the cp.async kernels to rounding (different order inside a k-step), not bitwise.
#include <cuda.h>
#include <stdint.h>
#include <stdlib.h>

#include "trpo_internal.cuh"
#include "dmma_common.cuh"
#include "tma_common.cuh"

namespace {
using namespace tma;

constexpr int BM = 128, BN = 64, BK = 32, NT = 256;
constexpr int BOX_K_BYTES = BK * 128;                     // a [32 k-rows x 16 columns] box (outer product operands)
constexpr int STAGE_BYTES = (BM / 16 + BN / 16) * BOX_K_BYTES;      // 48 KB
constexpr int SMEM_BYTES = 2 * STAGE_BYTES + 1024;        // + slack to align the stages to the 1 KB swizzle atom

__device__ __forceinline__ double act_deriv(char a, double y) {   // f'(x) expressed through y = f(x)
    switch (a) {
        case 't': return 1.0 - y * y;
        case 'o': return 0.1;
        case 's': return y * (1.0 - y);
        default:  return 1.0;
    }
}
__device__ __forceinline__ unsigned char *align_1k(unsigned char *p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap *map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" :: "l"(map), "r"(c0), "r"(c1) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// outer: out[slice][m*N + n] (+)= sum_{s in slice} Yprev[s][m] * G[s][n] for m < M0; row M0 (bias gradient) = column sums of G,
// formed by the m-tile-0 CTAs from the B boxes they stage anyway (TRPO_FVP.c:890-899 summed over the samples, :903-921).
__global__ void __launch_bounds__(NT, 2) k_outer_tma(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapG,
                                                     int rows, int M0, int N, int per_slice, int tiles_n,
                                                     double *__restrict__ partial, int P, int out_off, int accumulate,
                                                     const int *__restrict__ done) {
    if (done && *done) return;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full[2];
    unsigned char *smem = align_1k(smem_dyn);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3, wm = w >> 1, wn = w & 1;
    const int m0 = (blockIdx.x / tiles_n) * BM, n0 = (blockIdx.x % tiles_n) * BN;
    const int slice = blockIdx.y, s0 = slice * per_slice, s1 = min(rows, s0 + per_slice);
    const int nk = s1 > s0 ? (s1 - s0 + BK - 1) / BK : 0;        // per_slice is a multiple of BK: a k-step never straddles two slices
    if (tid == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double acc[4][4][2] = {};
    const bool colsum = m0 == 0;
    double bsum = 0.0;
    const bool active = m0 + 32 * wm < M0 && n0 + 32 * wn < N;   // sub-tiles outside the matrix issue no DMMAs
    int off[2][2];                                        // lane offsets through the swizzle: [q & 1][8-column half of a box]
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int ib = 0; ib < 2; ++ib) off[c][ib] = (2 * t + c) * 128 + (((4 * ib + (g >> 1)) ^ (2 * t + c)) << 4) + (g & 1) * 8;
    auto issue = [&](int st, int ks) {
        unsigned char *base = smem + st * STAGE_BYTES;
        mbar_expect_tx(&full[st], STAGE_BYTES);
#pragma unroll
        for (int mb = 0; mb < BM / 16; ++mb) tma_load_2d(base + mb * BOX_K_BYTES, &mapY, &full[st], m0 + 16 * mb, ks);
#pragma unroll
        for (int nb = 0; nb < BN / 16; ++nb) tma_load_2d(base + (BM / 16 + nb) * BOX_K_BYTES, &mapG, &full[st], n0 + 16 * nb, ks);
    };
    if (tid == 0 && nk) issue(0, s0);
    // column sums: thread (column n, k-phase) adds its 8 rows of every k-step; box byte of (k, n) as above
    const int cn = tid & 63, cph = tid >> 6;
    const unsigned char *cs_base = smem + (BM / 16 + (cn >> 4)) * BOX_K_BYTES + cph * 1024 + ((cn & 15) & 1) * 8;
    for (int it = 0; it < nk; ++it) {
        __syncthreads();                                   // everybody is done with the other stage
        if (tid == 0 && it + 1 < nk) issue((it + 1) & 1, s0 + (it + 1) * BK);
        mbar_wait(&full[it & 1], (it >> 1) & 1);
        const unsigned char *A = smem + (it & 1) * STAGE_BYTES + 2 * wm * BOX_K_BYTES;
        const unsigned char *B = smem + (it & 1) * STAGE_BYTES + (BM / 16 + 2 * wn) * BOX_K_BYTES;
        if (active) {
#pragma unroll
            for (int q = 0; q < BK / 4; ++q) {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const double *>(A + (i >> 1) * BOX_K_BYTES + (q >> 1) * 1024 + off[q & 1][i & 1]);
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double *>(B + (j >> 1) * BOX_K_BYTES + (q >> 1) * 1024 + off[q & 1][j & 1]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
            }
        }
        if (colsum) {
            const unsigned char *cb = cs_base + (it & 1) * STAGE_BYTES;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) bsum += *reinterpret_cast<const double *>(cb + kk * 128 + ((((cn & 15) >> 1) ^ kk) << 4));
        }
    }
    double *out = partial + (size_t)slice * P + out_off;
    if (colsum) {
        __syncthreads();                                   // everybody is done with the stage buffers
        double *red = reinterpret_cast<double *>(smem);
        red[tid] = bsum;
        __syncthreads();
        const int gn = n0 + tid;
        if (tid < BN && gn < N) {
            const double sum = ((red[tid] + red[tid + 64]) + red[tid + 128]) + red[tid + 192];
            const size_t o = (size_t)M0 * N + gn;
            out[o] = accumulate ? out[o] + sum : sum;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + 32 * wm + 8 * i + g;
        if (gm >= M0) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {                  // scalar: slice * P + out_off need not be even
                const int gn = n0 + 32 * wn + 8 * j + 2 * t + r;
                if (gn >= N) continue;
                const size_t o = (size_t)gm * N + gn;
                out[o] = accumulate ? out[o] + acc[i][j][r] : acc[i][j][r];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// backward: Gout[s][n] = f'(Yprev[s][n]) * sum_k Gin[s][k] * W[n][k]   (W is [N x Kd] row-major: already K-major for this product)
// Boxes: Gin [128 rows x 16 k], W [64 rows x 16 k], two of each per 32-wide k-step; after the main loop the Yprev tile
// [128 rows x 64 columns] arrives in the then idle stage memory by TMA as well (its L2 prefetch is issued at kernel start).
constexpr int BWD_A_BOX = BM * 128, BWD_B_BOX = BN * 128;         // 16 KB, 8 KB
static_assert(2 * BWD_A_BOX + 2 * BWD_B_BOX == STAGE_BYTES, "backward stage = outer stage");
__global__ void __launch_bounds__(NT, 2) k_bwd_tma(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapW,
                                                   const __grid_constant__ CUtensorMap mapY, int rows, int Kd, int N, char act_prev,
                                                   double *__restrict__ Gout, const int *__restrict__ done) {
    if (done && *done) return;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full[3];
    unsigned char *smem = align_1k(smem_dyn);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3, wm = w >> 1, wn = w & 1;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;     // the n-tiles of one row block run together: its A boxes are read from L2 once
    const int nk = (Kd + BK - 1) / BK;
    if (tid == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&full[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int nb = 0; nb < BN / 16; ++nb) tma_prefetch_l2_2d(&mapY, n0 + 16 * nb, m0);
    }
    __syncthreads();
    double acc[4][4][2] = {};
    int off[4];                                           // lane offsets through the swizzle for c0 = q & 3
#pragma unroll
    for (int c0 = 0; c0 < 4; ++c0) off[c0] = g * 128 + (((c0 + 4 * (t >> 1)) ^ g) << 4) + (t & 1) * 8;
    auto issue = [&](int st, int k0) {
        unsigned char *base = smem + st * STAGE_BYTES;
        mbar_expect_tx(&full[st], STAGE_BYTES);
        tma_load_2d(base, &mapG, &full[st], k0, m0);
        tma_load_2d(base + BWD_A_BOX, &mapG, &full[st], k0 + 16, m0);
        tma_load_2d(base + 2 * BWD_A_BOX, &mapW, &full[st], k0, n0);
        tma_load_2d(base + 2 * BWD_A_BOX + BWD_B_BOX, &mapW, &full[st], k0 + 16, n0);
    };
    if (tid == 0) issue(0, 0);
    for (int it = 0; it < nk; ++it) {
        __syncthreads();
        if (tid == 0 && it + 1 < nk) issue((it + 1) & 1, (it + 1) * BK);
        mbar_wait(&full[it & 1], (it >> 1) & 1);
        const unsigned char *A = smem + (it & 1) * STAGE_BYTES + 32 * wm * 128;
        const unsigned char *B = smem + (it & 1) * STAGE_BYTES + 2 * BWD_A_BOX + 32 * wn * 128;
#pragma unroll
        for (int q = 0; q < BK / 4; ++q) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const double *>(A + (q >> 2) * BWD_A_BOX + i * 1024 + off[q & 3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double *>(B + (q >> 2) * BWD_B_BOX + j * 1024 + off[q & 3]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
        }
    }
    // the Yprev tile: 4 boxes of [128 rows x 16 columns] over both (now idle) stages
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&full[2], 4 * BWD_A_BOX);
#pragma unroll
        for (int nb = 0; nb < BN / 16; ++nb) tma_load_2d(smem + nb * BWD_A_BOX, &mapY, &full[2], n0 + 16 * nb, m0);
    }
    mbar_wait(&full[2], 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = 32 * wm + 8 * i + g, gm = m0 + r;
        if (gm >= rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + 32 * wn + 8 * j + 2 * t;        // N is even
            if (gn >= N) continue;
            // columns 8(j&1) + 2t, +1 of box 2wn + (j>>1): one 16-byte chunk, index 4(j&1) + t, swizzled by the row
            const double2 y = *reinterpret_cast<const double2 *>(smem + (2 * wn + (j >> 1)) * BWD_A_BOX + r * 128 + (((4 * (j & 1) + t) ^ g) << 4));
            *reinterpret_cast<double2 *>(&Gout[(size_t)gm * N + gn]) =
                make_double2(acc[i][j][0] * act_deriv(act_prev, y.x), acc[i][j][1] * act_deriv(act_prev, y.y));
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// forward R-op layer (TRPO_FVP.c:795-834): [Y | RY] = f([Yin | RYin] W, RYin W + Yin VW) with both biases as the accumulators'
// start values. 64 x 64 CTA tile, 8 warps of 32 x 16, k-step 16, two CTAs per SM (as k_chain_fwd<.., .., 64>).
// Yin / RYin boxes are [64 rows x 16 k] (k along the 128-byte box row). The weights' box rows ARE k, and the k order the
// activations' layout dictates (lane t of step c0 holds k = 2 c0 + 8 (t/2) + (t&1)) would put lanes t and t + 2 on rows 8 apart --
// the same swizzle phase, a 2-way conflict. So the weights are staged from a copy whose rows are permuted inside every group of
// 16 (k_permute_rows16: bit 0 <-> bit 1, bit 2 <-> bit 3 of the row index): lane t then reads row 8 (c0/2) + 2t + (c0&1), four
// distinct swizzle phases. The copy is rebuilt per FVP for W and the direction (2 P elements, microseconds).
constexpr int FWD_TM = 64, FWD_BK = 16;
constexpr int FWD_A_BOX = FWD_TM * 128;                   // 8 KB
constexpr int FWD_B_BOX = FWD_BK * 128;                   // 2 KB: [16 k-rows x 16 columns]
template <bool HAS_RA> struct FwdCfg {
    static constexpr int STAGE = FWD_A_BOX * (HAS_RA ? 2 : 1) + 2 * (BN / 16) * FWD_B_BOX;       // 32 KB / 24 KB
    static constexpr int NS = HAS_RA ? 3 : 4;
    static constexpr int SMEM = NS * STAGE + 1024;
};

__global__ void k_permute_rows16(const double *__restrict__ W, double *__restrict__ Wp, int Kd, int N, int Kpad) {
    const size_t total = (size_t)Kpad * N;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(idx / N), n = (int)(idx % N), rho = row & 15;
        const int kp = ((rho & 1) << 1) | ((rho & 2) >> 1) | ((rho & 4) << 1) | ((rho & 8) >> 1);
        const int k = (row & ~15) + kp;
        Wp[idx] = k < Kd ? W[(size_t)k * N + n] : 0.0;
    }
}

template <bool HAS_RA>
__global__ void __launch_bounds__(NT, 2) k_fwd_tma(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapRY,
                                                   const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapV,
                                                   const double *__restrict__ W, const double *__restrict__ VW,
                                                   int rows, int Kd, int N, char act,
                                                   double *__restrict__ Yout, double *__restrict__ RYout, double *__restrict__ Gout,
                                                   const double *__restrict__ inv_var, const int *__restrict__ done) {
    if (done && *done) return;
    using C = FwdCfg<HAS_RA>;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full[C::NS];
    __shared__ double exp2_tab[64];
    unsigned char *smem = align_1k(smem_dyn);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3, wm = w >> 2, wn = w & 3;
    const int m0 = blockIdx.y * FWD_TM, n0 = blockIdx.x * BN;   // the n-tiles of one row block run together: its A boxes are read from L2 once
    const int nk = (Kd + FWD_BK - 1) / FWD_BK;
    load_exp2_table(exp2_tab);                            // visible after the first barrier
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < C::NS; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int st, int k0) {
        unsigned char *base = smem + st * C::STAGE;
        mbar_expect_tx(&full[st], C::STAGE);
        tma_load_2d(base, &mapY, &full[st], k0, m0);
        if (HAS_RA) tma_load_2d(base + FWD_A_BOX, &mapRY, &full[st], k0, m0);
        unsigned char *bw = base + (HAS_RA ? 2 : 1) * FWD_A_BOX;
#pragma unroll
        for (int nb = 0; nb < BN / 16; ++nb) {
            tma_load_2d(bw + nb * FWD_B_BOX, &mapW, &full[st], n0 + 16 * nb, k0);
            tma_load_2d(bw + (BN / 16 + nb) * FWD_B_BOX, &mapV, &full[st], n0 + 16 * nb, k0);
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < C::NS - 1; ++s) if (s < nk) issue(s, s * FWD_BK);
    }
    // both biases are the accumulators' start values (the contraction runs over Kd, no augmented row)
    double acc[4][2][2], racc[4][2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int gn = n0 + 16 * wn + 8 * j + 2 * t + r;
            const double bw = gn < N ? W[(size_t)Kd * N + gn] : 0.0, bv = gn < N ? VW[(size_t)Kd * N + gn] : 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc[i][j][r] = bw; racc[i][j][r] = bv; }
        }
    const bool active = n0 + 16 * wn < N;                 // a warp whose 16 columns lie outside the matrix issues no DMMAs
    int offA[4], offB[2][2];
#pragma unroll
    for (int c0 = 0; c0 < 4; ++c0) offA[c0] = (32 * wm + g) * 128 + (((c0 + 4 * (t >> 1)) ^ g) << 4) + (t & 1) * 8;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < 2; ++j) offB[c][j] = wn * FWD_B_BOX + (2 * t + c) * 128 + (((4 * j + (g >> 1)) ^ (2 * t + c)) << 4) + (g & 1) * 8;
    for (int it = 0; it < nk; ++it) {
        __syncthreads();                                   // everybody is done with the stage of step it - 1
        if (tid == 0 && it + C::NS - 1 < nk) issue((it + C::NS - 1) % C::NS, (it + C::NS - 1) * FWD_BK);
        mbar_wait(&full[it % C::NS], (it / C::NS) & 1);
        if (!active) continue;
        const unsigned char *A = smem + (it % C::NS) * C::STAGE, *RA = A + FWD_A_BOX;
        const unsigned char *B = A + (HAS_RA ? 2 : 1) * FWD_A_BOX, *VB = B + (BN / 16) * FWD_B_BOX;
#pragma unroll
        for (int q = 0; q < FWD_BK / 4; ++q) {
            double a[4], ra[4], b[2], vb[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = *reinterpret_cast<const double *>(A + i * 1024 + offA[q]);
                if (HAS_RA) ra[i] = *reinterpret_cast<const double *>(RA + i * 1024 + offA[q]);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                b[j] = *reinterpret_cast<const double *>(B + (q >> 1) * 1024 + offB[q & 1][j]);
                vb[j] = *reinterpret_cast<const double *>(VB + (q >> 1) * 1024 + offB[q & 1][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (HAS_RA) dmma(racc[i][j], ra[i], b[j]);
                    dmma(acc[i][j], a[i], b[j]);
                    dmma(racc[i][j], a[i], vb[j]);
                }
        }
    }
    if (!active) return;
    // epilogue: activation, R{y} = R{x} f'(x), last layer: R-gradient seed RG_K = Ry_K / sigma^2 * f' (TRPO_FVP.c:852-882)
#pragma unroll
    for (int i0 = 0; i0 < 4; i0 += 2) {
        double xv[8], dv[8];                               // 8 values per lock-step tanh: 2 row blocks x 2 tiles x 2 columns
#pragma unroll
        for (int ii = 0; ii < 2; ++ii)
#pragma unroll
            for (int j = 0; j < 2; ++j) { xv[(ii * 2 + j) * 2] = acc[i0 + ii][j][0]; xv[(ii * 2 + j) * 2 + 1] = acc[i0 + ii][j][1]; }
        if (act == 't') tanh_vec<8>(xv, dv, exp2_tab);
        else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const double x = xv[e];
                switch (act) {
                    case 'o': xv[e] = 0.1 * x; break;
                    case 's': xv[e] = 1.0 / (1.0 + exp(-x)); break;
                    default: break;
                }
                dv[e] = act_deriv(act, xv[e]);
            }
        }
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {
            const int i = i0 + ii, gm = m0 + 32 * wm + 8 * i + g;
            if (gm >= rows) continue;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int gn = n0 + 16 * wn + 8 * j + 2 * t;    // N is even: both columns of the pair are inside or outside
                if (gn >= N) continue;
                const double y0 = xv[(ii * 2 + j) * 2], y1 = xv[(ii * 2 + j) * 2 + 1], d0 = dv[(ii * 2 + j) * 2], d1 = dv[(ii * 2 + j) * 2 + 1];
                const size_t o = (size_t)gm * N + gn;
                if (Yout) *reinterpret_cast<double2 *>(&Yout[o]) = make_double2(y0, y1);
                const double r0 = racc[i][j][0] * d0, r1 = racc[i][j][1] * d1;
                if (RYout) *reinterpret_cast<double2 *>(&RYout[o]) = make_double2(r0, r1);
                if (Gout) *reinterpret_cast<double2 *>(&Gout[o]) = make_double2(r0 * inv_var[gn] * d0, r1 * inv_var[gn + 1] * d1);
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// tail: the last weight layer of the FVP for a narrow action layer (A <= 24, linear / 0.1x output) -- what k_chain_tail does, with
// the operands of the forward product by TMA. That kernel issued 14 LDGSTS per thread for 48 DMMAs per warp and k-step (the big
// GEMMs: 12 for 128), so its forward loop was bound by the copy instructions, not by the tensor pipe.
//   Rx_K = Ry_{K-1} W + y_{K-1} VW + VB (TRPO_FVP.c:795-803): y / Ry boxes [64 rows x 16 k]; W / VW from row-permuted copies
//        padded to 32 columns (two boxes of [16 k-rows x 16]); VB is the start value of the second accumulator set
//   RG_K = Rx_K f'^2 / sigma^2 (:809-823, :852-854, :869-882);  RG_{K-1} = (RG_K W^T) .* f'(y_{K-1}) (:890-899)
// W^T (zero padded [8 NTA][HP + 2], built once per FVP by k_tail_wt_image) arrives as ONE bulk copy over the idle ring.
constexpr int TAIL_TM = 64, TAIL_NT = 128, TAIL_NS = 3;
constexpr int TAIL_STAGE = 2 * FWD_A_BOX + 4 * FWD_B_BOX;         // y, Ry, 2 W boxes, 2 VW boxes: 24 KB

__global__ void k_permute_rows16_pad(const double *__restrict__ W, double *__restrict__ Wp, int Kd, int Nin, int Nout, int Kpad) {
    const size_t total = (size_t)Kpad * Nout;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(idx / Nout), n = (int)(idx % Nout), rho = row & 15;
        const int kp = ((rho & 1) << 1) | ((rho & 2) >> 1) | ((rho & 4) << 1) | ((rho & 8) >> 1);
        const int k = (row & ~15) + kp;
        Wp[idx] = (k < Kd && n < Nin) ? W[(size_t)k * Nin + n] : 0.0;
    }
}
// WT[k][j] = W[j][k] for k < A, j < H, zero elsewhere; row stride RST
__global__ void k_tail_wt_image(const double *__restrict__ W, double *__restrict__ WT, int H, int A, int AP, int RST) {
    const int total = AP * RST;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int k = idx / RST, j = idx % RST;
        WT[idx] = (k < A && j < H) ? W[(size_t)j * A + k] : 0.0;
    }
}

template <int NTA>
__global__ void __launch_bounds__(TAIL_NT, 3) k_tail_tma(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapRY,
                                                         const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapV,
                                                         const double *__restrict__ Y, const double *__restrict__ VW,
                                                         const double *__restrict__ WTg, int rows, int H, int A, char act_prev, double d3,
                                                         const double *__restrict__ inv_var,
                                                         double *__restrict__ GK, int ldgk, double *__restrict__ Gprev,
                                                         const int *__restrict__ done) {
    if (done && *done) return;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full[TAIL_NS + 1];
    unsigned char *smem = align_1k(smem_dyn);
    constexpr int AP = 8 * NTA;
    const int HP = (H + 7) & ~7, RST = HP + 2;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.x * TAIL_TM;
    const int nk = (H + FWD_BK - 1) / FWD_BK;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s <= TAIL_NS; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int st, int k0) {
        unsigned char *base = smem + st * TAIL_STAGE;
        mbar_expect_tx(&full[st], TAIL_STAGE);
        tma_load_2d(base, &mapY, &full[st], k0, m0);
        tma_load_2d(base + FWD_A_BOX, &mapRY, &full[st], k0, m0);
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            tma_load_2d(base + 2 * FWD_A_BOX + nb * FWD_B_BOX, &mapW, &full[st], 16 * nb, k0);
            tma_load_2d(base + 2 * FWD_A_BOX + (2 + nb) * FWD_B_BOX, &mapV, &full[st], 16 * nb, k0);
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TAIL_NS - 1; ++s) if (s < nk) issue(s, s * FWD_BK);
    }
    double rxa[2][NTA][2] = {}, rxb[2][NTA][2];
#pragma unroll
    for (int j = 0; j < NTA; ++j)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int col = 8 * j + 2 * t + r;
            const double vb = col < A ? VW[(size_t)H * A + col] : 0.0;         // VB: the bias row of the direction
            rxb[0][j][r] = vb; rxb[1][j][r] = vb;
        }
    int offA[4], offB[2][2];
#pragma unroll
    for (int c0 = 0; c0 < 4; ++c0) offA[c0] = (16 * w + g) * 128 + (((c0 + 4 * (t >> 1)) ^ g) << 4) + (t & 1) * 8;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int jb = 0; jb < 2; ++jb) offB[c][jb] = (2 * t + c) * 128 + (((4 * jb + (g >> 1)) ^ (2 * t + c)) << 4) + (g & 1) * 8;
    for (int it = 0; it < nk; ++it) {
        __syncthreads();                                   // everybody is done with the stage of step it - 1
        if (tid == 0 && it + TAIL_NS - 1 < nk) issue((it + TAIL_NS - 1) % TAIL_NS, (it + TAIL_NS - 1) * FWD_BK);
        mbar_wait(&full[it % TAIL_NS], (it / TAIL_NS) & 1);
        const unsigned char *As = smem + (it % TAIL_NS) * TAIL_STAGE, *RAs = As + FWD_A_BOX;
        const unsigned char *Bs = As + 2 * FWD_A_BOX, *VBs = Bs + 2 * FWD_B_BOX;
#pragma unroll
        for (int q = 0; q < FWD_BK / 4; ++q) {
            double a[2], ra[2], b[NTA], vb[NTA];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                a[i] = *reinterpret_cast<const double *>(As + i * 1024 + offA[q]);
                ra[i] = *reinterpret_cast<const double *>(RAs + i * 1024 + offA[q]);
            }
#pragma unroll
            for (int j = 0; j < NTA; ++j) {
                b[j] = *reinterpret_cast<const double *>(Bs + (j >> 1) * FWD_B_BOX + (q >> 1) * 1024 + offB[q & 1][j & 1]);
                vb[j] = *reinterpret_cast<const double *>(VBs + (j >> 1) * FWD_B_BOX + (q >> 1) * 1024 + offB[q & 1][j & 1]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < NTA; ++j) { dmma(rxa[i][j], ra[i], b[j]); dmma(rxb[i][j], a[i], vb[j]); }
        }
    }
    __syncthreads();                                            // the ring is idle: W^T takes its place, one bulk copy
    const double *WT = reinterpret_cast<const double *>(smem);
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)(AP * RST * sizeof(double));
        mbar_expect_tx(&full[TAIL_NS], bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(smem)), "l"(WTg), "r"(bytes), "r"(smem_u32(&full[TAIL_NS])) : "memory");
    }
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
                if (gm < rows && col < ldgk) GK[(size_t)gm * ldgk + col] = v;      // column A of an odd-width RG_K is padding: zero
            }
    }
    mbar_wait(&full[TAIL_NS], 0);
    // RG_{K-1} = (RG_K W^T) .* f'(y_{K-1}), 32 columns at a time; the y values are fetched before the block's DMMAs
    for (int n0 = 0; n0 < HP; n0 += 32) {
        double2 yv[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int gm = m0 + 16 * w + 8 * i + g;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = n0 + 8 * j + 2 * t;
                yv[i][j] = (gm < rows && col < H) ? *reinterpret_cast<const double2 *>(&Y[(size_t)gm * H + col]) : make_double2(0.0, 0.0);
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
                if (col >= H) continue;                    // H is even: the pair is inside or outside
                *reinterpret_cast<double2 *>(&Gprev[(size_t)gm * H + col]) =
                    make_double2(acc[i][j][0] * act_deriv(act_prev, yv[i][j].x), acc[i][j][1] * act_deriv(act_prev, yv[i][j].y));
            }
        }
    }
}

// row-major FP64 matrix [nrows x ncols] (contiguous rows), boxes of 16 columns x box_rows rows, SWIZZLE_128B, zero fill
bool make_map(CUtensorMap *m, const double *base, size_t nrows, int ncols, int box_rows, int ld = 0) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)ncols, (cuuint64_t)nrows};
    const cuuint64_t strides[1] = {(cuuint64_t)(ld ? ld : ncols) * sizeof(double)};      // ld: row stride in doubles (even), default dense
    const cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool configure() {
    static DeviceOnce once;
    if (!once.pending()) return true;
    bool r = cudaFuncSetAttribute(k_outer_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess &&
             cudaFuncSetAttribute(k_bwd_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess &&
             cudaFuncSetAttribute(k_fwd_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdCfg<true>::SMEM) == cudaSuccess &&
             cudaFuncSetAttribute(k_fwd_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdCfg<false>::SMEM) == cudaSuccess;
    if (r) once.mark();
    return r;
}
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// TRPO_NO_CHAIN_TMA=1 keeps the cp.async kernels (A/B timing, tests of both)
bool chain_tma_enabled() {
    const char *e = getenv("TRPO_NO_CHAIN_TMA");
    return !(e && atoi(e)) && tma::encode_fn() != nullptr;
}
// tensor-map rows are 16-byte multiples: even widths, 16-byte aligned bases
// ldg: row stride of G in doubles (the TMA-fed tail writes a 17-column RG_K with stride 18 so that it qualifies)
bool chain_tma_outer_eligible(const double *Yprev, const double *G, int M0, int N, int ldg) {
    return chain_tma_enabled() && (M0 & 1) == 0 && (ldg & 1) == 0 && M0 >= 16 && N >= 16 && aligned16(Yprev) && aligned16(G);
}
bool chain_tma_bwd_eligible(const double *Gin, const double *W, const double *Yprev, const double *Gout, int Kd, int N) {
    return chain_tma_enabled() && (Kd & 1) == 0 && (N & 1) == 0 && Kd >= 16 && N >= 16 && aligned16(Gin) && aligned16(W) &&
           aligned16(Yprev) && aligned16(Gout);
}
int chain_tma_tiles_m(int M0) { return (M0 + BM - 1) / BM; }

int chain_tma_outer(const double *Yprev, const double *G, int ldg, int rows, int M0, int N, int per_slice, int tiles_n, int nslices,
                    double *partial, int P, int out_off, int accumulate, const int *done, cudaStream_t st) {
    CUtensorMap mY, mG;
    if (!configure() || !make_map(&mY, Yprev, (size_t)rows, M0, BK) || !make_map(&mG, G, (size_t)rows, N, BK, ldg)) return -1;
    dim3 grid(chain_tma_tiles_m(M0) * tiles_n, nslices);
    k_outer_tma<<<grid, NT, SMEM_BYTES, st>>>(mY, mG, rows, M0, N, per_slice, tiles_n, partial, P, out_off, accumulate, done);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int chain_tma_bwd(const double *Gin, const double *W, const double *Yprev, int rows, int Kd, int N, char act_prev, double *Gout,
                  const int *done, cudaStream_t st) {
    CUtensorMap mG, mW, mY;
    if (!configure() || !make_map(&mG, Gin, (size_t)rows, Kd, BM) || !make_map(&mW, W, (size_t)N, Kd, BN) ||
        !make_map(&mY, Yprev, (size_t)rows, N, BM)) return -1;
    dim3 grid((N + BN - 1) / BN, (rows + BM - 1) / BM);
    k_bwd_tma<<<grid, NT, SMEM_BYTES, st>>>(mG, mW, mY, rows, Kd, N, act_prev, Gout, done);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---- forward ----
bool chain_tma_fwd_eligible(const double *Yin, const double *RYin, const double *Yout, const double *RYout, const double *Gout, int Kd, int N) {
    static const bool off = getenv("TRPO_NO_CHAIN_TMA_FWD") && atoi(getenv("TRPO_NO_CHAIN_TMA_FWD"));
    return !off && chain_tma_enabled() && (Kd & 1) == 0 && (N & 1) == 0 && Kd >= 16 && N >= 16 && aligned16(Yin) && aligned16(RYin) &&
           aligned16(Yout) && aligned16(RYout) && aligned16(Gout);
}
size_t chain_tma_perm_doubles(int Kd, int N) { return ((size_t)((Kd + 15) / 16 * 16) * N + 1) & ~(size_t)1; }
size_t chain_tma_perm_offset(const NetDesc &net, int layer) {
    size_t o = 0;
    for (int i = 0; i < layer && i < net.K; ++i) o += chain_tma_perm_doubles(net.L[i], net.L[i + 1]);
    return o;
}
void chain_tma_permute_rows(const double *W, double *Wp, int Kd, int N, cudaStream_t st) {
    const int Kpad = (Kd + 15) / 16 * 16;
    const size_t total = (size_t)Kpad * N;
    k_permute_rows16<<<(int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184), 256, 0, st>>>(W, Wp, Kd, N, Kpad);
}
// Wp / Vp: the row-permuted copies of W[Kd x N] / VW[Kd x N] (chain_tma_permute_rows); W / VW: the originals (bias rows)
int chain_tma_fwd(const double *Yin, const double *RYin, const double *Wp, const double *Vp, const double *W, const double *VW,
                  int rows, int Kd, int N, char act, double *Yout, double *RYout, double *Gout, const double *inv_var,
                  const int *done, cudaStream_t st) {
    const int Kpad = (Kd + 15) / 16 * 16;
    CUtensorMap mY, mRY, mW, mV;
    if (!configure() || !make_map(&mY, Yin, (size_t)rows, Kd, FWD_TM) || !make_map(&mW, Wp, (size_t)Kpad, N, FWD_BK) ||
        !make_map(&mV, Vp, (size_t)Kpad, N, FWD_BK)) return -1;
    if (RYin) { if (!make_map(&mRY, RYin, (size_t)rows, Kd, FWD_TM)) return -1; }
    else mRY = mY;
    dim3 grid((N + BN - 1) / BN, (rows + FWD_TM - 1) / FWD_TM);
    if (RYin) k_fwd_tma<true><<<grid, NT, FwdCfg<true>::SMEM, st>>>(mY, mRY, mW, mV, W, VW, rows, Kd, N, act, Yout, RYout, Gout, inv_var, done);
    else k_fwd_tma<false><<<grid, NT, FwdCfg<false>::SMEM, st>>>(mY, mRY, mW, mV, W, VW, rows, Kd, N, act, Yout, RYout, Gout, inv_var, done);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---- tail ----
// scratch behind the forward layers' permuted copies: W', VW' [ceil16(H) x 32] and the W^T image [24 x (HP + 2)]
size_t chain_tma_tail_doubles(int H) {
    const size_t Hpad = (H + 15) / 16 * 16, HP = (H + 7) & ~7;
    return 2 * Hpad * 32 + 24 * (HP + 2);
}
bool chain_tma_tail_eligible(const double *Y, const double *RY, const double *Gprev, int H, int A) {
    static const bool off = getenv("TRPO_NO_CHAIN_TMA_TAIL") && atoi(getenv("TRPO_NO_CHAIN_TMA_TAIL"));
    const size_t HP = (H + 7) & ~7;
    return !off && chain_tma_enabled() && (H & 1) == 0 && H >= 16 && A <= 24 && aligned16(Y) && aligned16(RY) && aligned16(Gprev) &&
           24 * (HP + 2) * sizeof(double) <= (size_t)TAIL_NS * TAIL_STAGE;
}
void chain_tma_tail_prepare(const double *W, const double *VW, double *scratch, int H, int A, cudaStream_t st) {
    const int Hpad = (H + 15) / 16 * 16, HP = (H + 7) & ~7, RST = HP + 2, AP = A <= 8 ? 8 : A <= 16 ? 16 : 24;
    double *Wp = scratch, *Vp = scratch + (size_t)Hpad * 32, *WT = Vp + (size_t)Hpad * 32;
    k_permute_rows16_pad<<<(Hpad * 32 + 255) / 256, 256, 0, st>>>(W, Wp, H, A, 32, Hpad);
    k_permute_rows16_pad<<<(Hpad * 32 + 255) / 256, 256, 0, st>>>(VW, Vp, H, A, 32, Hpad);
    k_tail_wt_image<<<(AP * RST + 255) / 256, 256, 0, st>>>(W, WT, H, A, AP, RST);
}
int chain_tma_tail(const double *Y, const double *RY, const double *VW, const double *scratch, int rows, int H, int A, char act_prev,
                   double d3, const double *inv_var, double *GK, int ldgk, double *Gprev, const int *done, cudaStream_t st) {
    const int Hpad = (H + 15) / 16 * 16;
    const double *Wp = scratch, *Vp = scratch + (size_t)Hpad * 32, *WT = Vp + (size_t)Hpad * 32;
    CUtensorMap mY, mRY, mW, mV;
    if (!make_map(&mY, Y, (size_t)rows, H, TAIL_TM) || !make_map(&mRY, RY, (size_t)rows, H, TAIL_TM) ||
        !make_map(&mW, Wp, (size_t)Hpad, 32, FWD_BK) || !make_map(&mV, Vp, (size_t)Hpad, 32, FWD_BK)) return -1;
    constexpr int SMEM = TAIL_NS * TAIL_STAGE + 1024;
    static DeviceOnce once;
    if (once.pending()) {
        if (cudaFuncSetAttribute(k_tail_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(k_tail_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(k_tail_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return -1;
        once.mark();
    }
    const int grid = (rows + TAIL_TM - 1) / TAIL_TM;
    if (A <= 8) k_tail_tma<1><<<grid, TAIL_NT, SMEM, st>>>(mY, mRY, mW, mV, Y, VW, WT, rows, H, A, act_prev, d3, inv_var, GK, ldgk, Gprev, done);
    else if (A <= 16) k_tail_tma<2><<<grid, TAIL_NT, SMEM, st>>>(mY, mRY, mW, mV, Y, VW, WT, rows, H, A, act_prev, d3, inv_var, GK, ldgk, Gprev, done);
    else k_tail_tma<3><<<grid, TAIL_NT, SMEM, st>>>(mY, mRY, mW, mV, Y, VW, WT, rows, H, A, act_prev, d3, inv_var, GK, ldgk, Gprev, done);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
