// gemm_chain_f32.cu -- optional FP32 mode of the Fisher-vector product (stated tolerance ~1e-4 relative).
//
// Same GEMM chain as gemm_chain.cu (forward R-op dual GEMM, backward GEMM, split-K outer product), but activations,
// weights and observations are FP32 and the products run on the tensor cores as 3xTF32: each FP32 operand is split
// into a TF32 head and a TF32 tail (x = hi + lo) and a product is issued as lo*hi + hi*lo + hi*hi with FP32 accumulation,
// which keeps ~FP32 accuracy (a single TF32 product would lose 13 mantissa bits and miss the 1e-4 target).
// Measured on this box (profiles/tf32_peak_r01.txt): mma.sync.m16n8k8.tf32 (SASS HMMA.1688.F32.TF32) 278 TFLOP/s, i.e.
// 93 TFLOP/s effective for 3xTF32 against 37.1 for the FP64 tensor pipe and 71.7 for plain FFMA.
// CTA tile 128 x 64, k-step 32; the dual forward kernel runs 16 warps of 32 x 16, the single-product kernels 8 warps of
// 32 x 32 at two CTAs per SM (as in gemm_chain.cu); 4-byte cp.async with zero fill, shared row strides 36 / 72 floats
// (conflict-free fragment reads). Per-slice partial sums are FP32; the fixed-order slice
// reduction and everything downstream (CG) are FP64.
#include <stdint.h>
#include <stdlib.h>

#include "trpo_internal.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 32, NT = 256;
constexpr int RSA = BK + 4, RSB = BN + 8;               // 36: g*4+t distinct banks; 72: t*8+g distinct banks
constexpr int A_TILE = BM * RSA, B_TILE = BK * RSB;     // floats per operand tile

__device__ __forceinline__ void cp_async4(float *dst_smem, const float *src, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" :: "r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(float *dst_smem, const float *src, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

// x = hi + lo with hi = x truncated to TF32's 10 mantissa bits (one LOP3) and lo = x - hi (exact in FP32; the tensor core
// ignores its low 13 mantissa bits, an error of <= 2^-20 |x|). cvt.rna.tf32.f32 has no native SASS form on sm_100a -- it
// expands to ~6 instructions each (FSETP/FMUL/FFMA/LOP3) and the kernel is instruction-issue bound; round-to-nearest
// splits bought ~2 bits that the stated 1e-4 tolerance does not need.
__device__ __forceinline__ void split_tf32(float x, unsigned &hi, unsigned &lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// c += a*b with 3xTF32 error compensation (small terms first)
__device__ __forceinline__ void mma3(float (&c)[4], const unsigned (&ahi)[4], const unsigned (&alo)[4],
                                     const unsigned (&bhi)[2], const unsigned (&blo)[2]) {
    mma_tf32(c, alo, bhi);
    mma_tf32(c, ahi, blo);
    mma_tf32(c, ahi, bhi);
}

__device__ __forceinline__ float act_apply(char a, float x) {
    switch (a) {
        case 't': return tanhf(x);
        case 'o': return 0.1f * x;
        case 's': return 1.0f / (1.0f + expf(-x));
        default:  return x;
    }
}
__device__ __forceinline__ float act_deriv(char a, float y) {
    switch (a) {
        case 't': return 1.0f - y * y;
        case 'o': return 0.1f;
        case 's': return y * (1.0f - y);
        default:  return 1.0f;
    }
}

template <int NTH, int TM = BM>
__device__ __forceinline__ void load_a_rowmajor(float *As, const float *X, int rows, int ld, int m0, int k0,
                                                bool aug, float ones_val, int tid) {
    if (X != nullptr && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {
        // leading dimension a multiple of 4 floats: 16-byte copies, a quarter of the copy instructions and index arithmetic
        // (the kernel is instruction-issue bound: the 3xTF32 splits already cost ~2 ALU instructions per fragment element)
#pragma unroll
        for (int it = 0; it < TM * BK / 4 / NTH; ++it) {
            const int idx = tid + it * NTH, m = idx / (BK / 4), k = (idx % (BK / 4)) * 4;
            const int gm = m0 + m, gk = k0 + k;
            float *dst = &As[m * RSA + k];
            if (aug && gk == ld && gm < rows) { dst[0] = ones_val; dst[1] = dst[2] = dst[3] = 0.0f; }
            else {
                const bool in = gm < rows && gk < ld;      // ld % 4 == 0: gk < ld implies gk + 3 < ld
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
        else cp_async4(&As[m * RSA + k], in ? &X[(size_t)gm * ld + gk] : X, in ? 4 : 0);
    }
}
// The outer product's A operand in its natural orientation (see gemm_chain.cu): An[k][m] = Yprev[(s0+k)*M0 + m0+m], 1 for
// m == M0. Row stride RSN % 32 == 8: the m16n8k8 A fragment (rows g / g+8, columns t / t+4) reads (t)*RSN + g -> 32 banks.
constexpr int RSN = BM + 8;
__device__ __forceinline__ void load_a_natural(float *An, const float *Y, int s_end, int M0, int m0, int s0, int tid) {
    if (Y != nullptr && (M0 & 3) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0) {
#pragma unroll
        for (int it = 0; it < BK * BM / 4 / NT; ++it) {
            const int idx = tid + it * NT, k = idx >> 5, m = (idx & 31) * 4;
            const int gm = m0 + m, gs = s0 + k;
            float *dst = &An[k * RSN + m];
            if (gm == M0) { dst[0] = gs < s_end ? 1.0f : 0.0f; dst[1] = dst[2] = dst[3] = 0.0f; }
            else {
                const bool in = gm < M0 && gs < s_end;
                cp_async16(dst, in ? &Y[(size_t)gs * M0 + gm] : Y, in ? 16 : 0);
            }
        }
        return;
    }
#pragma unroll 2
    for (int it = 0; it < BK * BM / NT; ++it) {
        const int idx = tid + it * NT, k = idx >> 7, m = idx & 127;
        const int gm = m0 + m, gs = s0 + k;
        const bool in = Y != nullptr && gm < M0 && gs < s_end;
        if (gm == M0) An[k * RSN + m] = gs < s_end ? 1.0f : 0.0f;
        else cp_async4(&An[k * RSN + m], in ? &Y[(size_t)gs * M0 + gm] : Y, in ? 4 : 0);
    }
}
template <int NTH>
__device__ __forceinline__ void load_b_rowmajor(float *Bs, const float *M, int kdim, int N, int k0, int n0, int tid) {
    if ((N & 3) == 0 && (reinterpret_cast<uintptr_t>(M) & 15) == 0) {
#pragma unroll
        for (int it = 0; it < BK * BN / 4 / NTH; ++it) {
            const int idx = tid + it * NTH, k = idx >> 4, n = (idx & 15) * 4;
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
        cp_async4(&Bs[k * RSB + n], in ? &M[(size_t)gk * N + gn] : M, in ? 4 : 0);
    }
}
__device__ __forceinline__ void load_b_transposed(float *Bs, const float *W, int Kd, int N, int k0, int n0, int tid) {
#pragma unroll
    for (int it = 0; it < BK * BN / NT; ++it) {
        const int idx = tid + it * NT, n = idx / BK, k = idx % BK;
        const int gk = k0 + k, gn = n0 + n;
        const bool in = gk < Kd && gn < N;
        cp_async4(&Bs[k * RSB + n], in ? &W[(size_t)gn * Kd + gk] : W, in ? 4 : 0);
    }
}

// one k-step (32) of a warp's 32 x 32 sub-tile. Fragment layout of m16n8k8 (lane = 4g + t):
//   A: (g, t) (g+8, t) (g, t+4) (g+8, t+4)    B: (t, g) (t+4, g)    C: (g, 2t) (g, 2t+1) (g+8, 2t) (g+8, 2t+1)
template <bool DUAL, bool HAS_RA, int NJ, bool ANAT = false>
__device__ __forceinline__ void mma_stage(float (&acc)[2][NJ][4], float (&racc)[2][NJ][4], const float *As, const float *RAs,
                                          const float *Bs, const float *VBs, int wm, int wn, int g, int t) {
#pragma unroll
    for (int q = 0; q < BK / 8; ++q) {
        unsigned ahi[2][4], alo[2][4], rhi[2][4], rlo[2][4], bhi[NJ][2], blo[NJ][2], vhi[NJ][2], vlo[NJ][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int row = 32 * wm + 16 * i + g + 8 * (e & 1), col = 8 * q + t + 4 * (e >> 1);
                split_tf32(ANAT ? As[col * RSN + row] : As[row * RSA + col], ahi[i][e], alo[i][e]);
                if (DUAL && HAS_RA) split_tf32(RAs[row * RSA + col], rhi[i][e], rlo[i][e]);
            }
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = 8 * q + t + 4 * e, col = 8 * NJ * wn + 8 * j + g;
                split_tf32(Bs[row * RSB + col], bhi[j][e], blo[j][e]);
                if (DUAL) split_tf32(VBs[row * RSB + col], vhi[j][e], vlo[j][e]);
            }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                mma3(acc[i][j], ahi[i], alo[i], bhi[j], blo[j]);
                if (DUAL) {
                    if (HAS_RA) mma3(racc[i][j], rhi[i], rlo[i], bhi[j], blo[j]);
                    mma3(racc[i][j], ahi[i], alo[i], vhi[j], vlo[j]);
                }
            }
    }
}

// TM = rows per CTA tile: 128 rows x 512 threads (one CTA per SM) or 64 rows x 256 threads (two per SM), see gemm_chain.cu
template <bool DUAL, bool HAS_RA, int TM = BM>
__global__ void __launch_bounds__(TM == BM ? 512 : 256, TM == BM ? 1 : 2) k_fwd(const float *__restrict__ Yin, const float *__restrict__ RYin,
                                               const float *__restrict__ W, const float *__restrict__ VW,
                                               int rows, int Kd, int N, char act,
                                               float *__restrict__ Yout, float *__restrict__ RYout,
                                               float *__restrict__ Gout, const float *__restrict__ inv_var,
                                               const int *__restrict__ done) {
    if (done && *done) return;
    extern __shared__ __align__(16) float smem_f[];
    constexpr int AT = TM * RSA;                             // floats per A tile of this variant
    constexpr int STAGE = AT * (DUAL && HAS_RA ? 2 : 1) + B_TILE * (DUAL ? 2 : 1);
    constexpr int NTH = TM == BM ? 512 : 256, WN = 4, NJ = BN / 8 / WN;      // (TM / 32) x 4 warps, 32 x 16 warp tiles
    static_assert(NTH / 32 / WN * 32 == TM, "warp rows must cover the tile");
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3, wm = w / WN, wn = w % WN;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * BN;
    float acc[2][NJ][4] = {}, racc[2][NJ][4] = {};
    // as in gemm_chain.cu: a width that is a whole number of k-steps takes its bias as the accumulators' start value
    const bool bias_init = (Kd % BK) == 0;
    const int nk = bias_init ? Kd / BK : (Kd + 1 + BK - 1) / BK;
    if (bias_init) {
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int gn = n0 + 8 * NJ * wn + 8 * j + 2 * t + (e & 1);
                const float bw = gn < N ? W[(size_t)Kd * N + gn] : 0.0f;
                const float bv = (DUAL && gn < N) ? VW[(size_t)Kd * N + gn] : 0.0f;
#pragma unroll
                for (int i = 0; i < 2; ++i) { acc[i][j][e] = bw; racc[i][j][e] = bv; }
            }
    }
    auto stage_ptrs = [&](int st, float *&As, float *&RAs, float *&Bs, float *&VBs) {
        float *p = smem_f + st * STAGE;
        As = p; p += AT;
        RAs = p; if (DUAL && HAS_RA) p += AT;
        Bs = p; p += B_TILE;
        VBs = p;
    };
    auto load = [&](int st, int k0) {
        float *As, *RAs, *Bs, *VBs;
        stage_ptrs(st, As, RAs, Bs, VBs);
        load_a_rowmajor<NTH, TM>(As, Yin, rows, Kd, m0, k0, true, 1.0f, tid);
        load_b_rowmajor<NTH>(Bs, W, Kd + 1, N, k0, n0, tid);
        if (DUAL) {
            if (HAS_RA) load_a_rowmajor<NTH, TM>(RAs, RYin, rows, Kd, m0, k0, false, 0.0f, tid);
            load_b_rowmajor<NTH>(VBs, VW, Kd + 1, N, k0, n0, tid);
        }
        cp_commit();
    };
    load(0, 0);
    for (int it = 0; it < nk; ++it) {
        if (it + 1 < nk) { load((it + 1) & 1, (it + 1) * BK); cp_wait<1>(); }
        else cp_wait<0>();
        __syncthreads();
        float *As, *RAs, *Bs, *VBs;
        stage_ptrs(it & 1, As, RAs, Bs, VBs);
        mma_stage<DUAL, HAS_RA, NJ>(acc, racc, As, RAs, Bs, VBs, wm, wn, g, t);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int gm = m0 + 32 * wm + 16 * i + g + 8 * (e >> 1), gn = n0 + 8 * NJ * wn + 8 * j + 2 * t + (e & 1);
                if (gm >= rows || gn >= N) continue;
                const float y = act_apply(act, acc[i][j][e]);
                const float d = act_deriv(act, y);
                if (Yout) Yout[(size_t)gm * N + gn] = y;
                if (DUAL) {
                    const float ry = racc[i][j][e] * d;
                    if (RYout) RYout[(size_t)gm * N + gn] = ry;
                    if (Gout) Gout[(size_t)gm * N + gn] = ry * inv_var[gn] * d;
                }
            }
}

__global__ void __launch_bounds__(NT, 2) k_bwd(const float *__restrict__ Gin, const float *__restrict__ W,
                                               const float *__restrict__ Yprev, int rows, int Kd, int N, char act_prev,
                                               float *__restrict__ Gout, const int *__restrict__ done) {
    if (done && *done) return;
    extern __shared__ __align__(16) float smem_f[];
    constexpr int STAGE = A_TILE + B_TILE;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3, wm = w >> 1, wn = w & 1;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[2][4][4] = {}, dummy[2][4][4];
    const int nk = (Kd + BK - 1) / BK;
    auto load = [&](int st, int k0) {
        float *As = smem_f + st * STAGE, *Bs = As + A_TILE;
        load_a_rowmajor<NT>(As, Gin, rows, Kd, m0, k0, false, 0.0f, tid);
        load_b_transposed(Bs, W, Kd, N, k0, n0, tid);
        cp_commit();
    };
    load(0, 0);
    for (int it = 0; it < nk; ++it) {
        if (it + 1 < nk) { load((it + 1) & 1, (it + 1) * BK); cp_wait<1>(); }
        else cp_wait<0>();
        __syncthreads();
        const float *As = smem_f + (it & 1) * STAGE, *Bs = As + A_TILE;
        mma_stage<false, false, 4>(acc, dummy, As, nullptr, Bs, nullptr, wm, wn, g, t);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int gm = m0 + 32 * wm + 16 * i + g + 8 * (e >> 1), gn = n0 + 32 * wn + 8 * j + 2 * t + (e & 1);
                if (gm >= rows || gn >= N) continue;
                Gout[(size_t)gm * N + gn] = acc[i][j][e] * act_deriv(act_prev, Yprev[(size_t)gm * N + gn]);
            }
}

__global__ void __launch_bounds__(NT, 2) k_outer(const float *__restrict__ Yprev, const float *__restrict__ G,
                                                 int rows, int M0, int N, int per_slice, int tiles_n,
                                                 float *__restrict__ partial, int P, int out_off, int accumulate,
                                                 int bias_colsum, const int *__restrict__ done) {
    if (done && *done) return;
    extern __shared__ __align__(16) float smem_f[];
    constexpr int AN_TILE = BK * RSN;                       // natural-orientation A tile
    static_assert(AN_TILE <= A_TILE && (AN_TILE & 3) == 0, "natural A tile must fit the stage and keep B 16-byte aligned");
    constexpr int STAGE = A_TILE + B_TILE;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3, wm = w >> 1, wn = w & 1;
    const int m0 = (blockIdx.x / tiles_n) * BM, n0 = (blockIdx.x % tiles_n) * BN;
    const int slice = blockIdx.y;
    const int s0 = slice * per_slice;
    const int s1 = min(rows, s0 + per_slice);
    float acc[2][4][4] = {}, dummy[2][4][4];
    // bias gradient as column sums of the staged G tiles, idle warps outside the matrix: see gemm_chain.cu k_chain_outer
    const bool colsum = bias_colsum && m0 == 0;
    const bool active = m0 + 32 * wm < M0 + (bias_colsum ? 0 : 1) && n0 + 32 * wn < N;
    float bsum = 0.0f;
    const int nk = s1 > s0 ? (s1 - s0 + BK - 1) / BK : 0;
    auto load = [&](int st, int ks) {
        float *As = smem_f + st * STAGE, *Bs = As + A_TILE;
        load_a_natural(As, Yprev, s1, M0, m0, ks, tid);
        load_b_rowmajor<NT>(Bs, G + (size_t)ks * N, s1 - ks, N, 0, n0, tid);
        cp_commit();
    };
    if (nk) load(0, s0);
    for (int it = 0; it < nk; ++it) {
        if (it + 1 < nk) { load((it + 1) & 1, s0 + (it + 1) * BK); cp_wait<1>(); }
        else cp_wait<0>();
        __syncthreads();
        const float *As = smem_f + (it & 1) * STAGE, *Bs = As + A_TILE;
        if (active) mma_stage<false, false, 4, true>(acc, dummy, As, nullptr, Bs, nullptr, wm, wn, g, t);
        if (colsum) {
#pragma unroll
            for (int kk = 0; kk < BK / 4; ++kk) bsum += Bs[(4 * kk + (tid >> 6)) * RSB + (tid & 63)];
        }
        __syncthreads();
    }
    float *out = partial + (size_t)slice * P + out_off;
    if (colsum) {
        smem_f[tid] = bsum;
        __syncthreads();
        const int gn = n0 + tid;
        if (tid < BN && gn < N) {
            const float sum = ((smem_f[tid] + smem_f[tid + 64]) + smem_f[tid + 128]) + smem_f[tid + 192];
            const size_t o = (size_t)M0 * N + gn;
            out[o] = accumulate ? out[o] + sum : sum;
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int gm = m0 + 32 * wm + 16 * i + g + 8 * (e >> 1), gn = n0 + 32 * wn + 8 * j + 2 * t + (e & 1);
                if (gm >= M0 + (bias_colsum ? 0 : 1) || gn >= N) continue;
                const size_t o = (size_t)gm * N + gn;
                out[o] = accumulate ? out[o] + acc[i][j][e] : acc[i][j][e];
            }
}

__global__ void k_to_float(const double *__restrict__ src, float *__restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = (float)src[i];
}

// tail: the LAST weight layer when the action dimension is small (A <= 32) and the output activation is linear / 0.1x -- the
// FP32 counterpart of k_chain_tail (gemm_chain.cu). One kernel replaces forward(K-1) + seed + backward(K), which as two
// 128 x 64-tile tensor-core GEMMs padded a 17-column layer to 64 columns and took 0.81 of the 3.56 ms of a 200 k-state FVP:
//   Rx_K = Ry_{K-1} W + y_{K-1} VW + VB   (TRPO_FVP.c:795-803),   RG_K = Rx_K f'^2 / sigma^2   (:852-854, :869-882),
//   RG_{K-1} = (RG_K W^T) .* f'(y_{K-1})  (:890-899).
// Plain FP32 FMAs: 13 k per sample is nothing against the 2 KB of activations it has to read (memory bound). A warp takes
// 4 rows at a time; lane l owns the hidden units l, l + 32, ...; the weights sit in shared memory at an odd row stride
// (conflict-free for lane-strided rows) and every weight read feeds 8 (forward) / 4 (backward) FMAs.
constexpr int TAIL_ROWS = 4, TAIL_AS = 33;
template <int JJ>      // JJ = H / 32 hidden units per lane
__global__ void __launch_bounds__(256) k_tail_f32(const float *__restrict__ Y, const float *__restrict__ RY,
                                                  const float *__restrict__ W, const float *__restrict__ VW, int rows, int A,
                                                  char act_prev, float d3, const float *__restrict__ inv_var,
                                                  float *__restrict__ GK, float *__restrict__ Gprev, float *__restrict__ Gprev_lo,
                                                  const int *__restrict__ done) {
    if (done && *done) return;
    constexpr int H = 32 * JJ;
    extern __shared__ __align__(16) float smem_f[];
    float *Ws = smem_f, *VWs = Ws + H * TAIL_AS, *vbs = VWs + H * TAIL_AS, *ivs = vbs + 32;
    for (int idx = threadIdx.x; idx < H * A; idx += blockDim.x) {
        const int j = idx / A, a = idx % A;
        Ws[j * TAIL_AS + a] = W[idx];
        VWs[j * TAIL_AS + a] = VW[idx];
    }
    for (int a = threadIdx.x; a < 32; a += blockDim.x) { vbs[a] = a < A ? VW[(size_t)H * A + a] : 0.0f; ivs[a] = a < A ? inv_var[a] : 0.0f; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = gridDim.x * (blockDim.x >> 5);
    for (long long r0 = (long long)(blockIdx.x * (blockDim.x >> 5) + warp) * TAIL_ROWS; r0 < rows; r0 += (long long)nwarp * TAIL_ROWS) {
        float y[TAIL_ROWS][JJ], ry[TAIL_ROWS][JJ];
#pragma unroll
        for (int r = 0; r < TAIL_ROWS; ++r)
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) {
                const bool in = r0 + r < rows;
                y[r][jj] = in ? Y[(size_t)(r0 + r) * H + lane + 32 * jj] : 0.0f;
                ry[r][jj] = in ? RY[(size_t)(r0 + r) * H + lane + 32 * jj] : 0.0f;
            }
        float g3[TAIL_ROWS];                                   // after the reduction lane a holds RG_K[r][a]
#pragma unroll
        for (int r = 0; r < TAIL_ROWS; ++r) g3[r] = 0.0f;
        for (int a = 0; a < A; ++a) {
            float acc[TAIL_ROWS];
#pragma unroll
            for (int r = 0; r < TAIL_ROWS; ++r) acc[r] = 0.0f;
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) {
                const float w = Ws[(lane + 32 * jj) * TAIL_AS + a], v = VWs[(lane + 32 * jj) * TAIL_AS + a];
#pragma unroll
                for (int r = 0; r < TAIL_ROWS; ++r) acc[r] = fmaf(ry[r][jj], w, fmaf(y[r][jj], v, acc[r]));
            }
#pragma unroll
            for (int r = 0; r < TAIL_ROWS; ++r) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
                if (lane == a) g3[r] = (acc[r] + vbs[a]) * d3 * ivs[a] * d3;
            }
        }
#pragma unroll
        for (int r = 0; r < TAIL_ROWS; ++r)
            if (lane < A && r0 + r < rows) GK[(size_t)(r0 + r) * A + lane] = g3[r];
        // RG_{K-1}[r][j] = f'(y[r][j]) * sum_a RG_K[r][a] W[j][a]
        float g2[TAIL_ROWS][JJ];
#pragma unroll
        for (int r = 0; r < TAIL_ROWS; ++r)
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) g2[r][jj] = 0.0f;
        for (int a = 0; a < A; ++a) {
            float ga[TAIL_ROWS];
#pragma unroll
            for (int r = 0; r < TAIL_ROWS; ++r) ga[r] = __shfl_sync(0xffffffffu, g3[r], a);
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) {
                const float w = Ws[(lane + 32 * jj) * TAIL_AS + a];
#pragma unroll
                for (int r = 0; r < TAIL_ROWS; ++r) g2[r][jj] = fmaf(ga[r], w, g2[r][jj]);
            }
        }
#pragma unroll
        for (int r = 0; r < TAIL_ROWS; ++r) {
            if (r0 + r >= rows) continue;
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) {
                const float gv = g2[r][jj] * act_deriv(act_prev, y[r][jj]);
                Gprev[(size_t)(r0 + r) * H + lane + 32 * jj] = gv;
                if (Gprev_lo) Gprev_lo[(size_t)(r0 + r) * H + lane + 32 * jj] = gv - __uint_as_float(__float_as_uint(gv) & 0xffffe000u);
            }
        }
    }
}
template <int JJ>
int launch_tail_f32(const float *Y, const float *RY, const float *W, const float *VW, int rows, int A, char act_prev, float d3,
                    const float *inv_var, float *GK, float *Gprev, float *Gprev_lo, const int *done, cudaStream_t st) {
    const size_t smem = sizeof(float) * (2 * 32 * JJ * TAIL_AS + 64);
    static DeviceOnce once;
    if (once.pending()) {
        if (cudaFuncSetAttribute(k_tail_f32<JJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        once.mark();
    }
    int grid = (rows + 8 * TAIL_ROWS - 1) / (8 * TAIL_ROWS);
    if (grid > 148 * 2) grid = 148 * 2;
    k_tail_f32<JJ><<<grid, 256, smem, st>>>(Y, RY, W, VW, rows, A, act_prev, d3, inv_var, GK, Gprev, Gprev_lo, done);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// zsum[e] (FP64) = fixed-order sum over slices of the FP32 partial rows
__global__ void __launch_bounds__(256) k_reduce_f32(const float *__restrict__ partial, int rows, int P,
                                                    double *__restrict__ zsum, const int *__restrict__ done) {
    if (done && *done) return;
    __shared__ double sh[8][32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5, e = blockIdx.x * 32 + cx;
    double s = 0.0;
    if (e < P)
        for (int r = ry; r < rows; r += 8) s += (double)partial[(size_t)r * P + e];
    sh[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && e < P) {
        double t = sh[0][cx];
#pragma unroll
        for (int k = 1; k < 8; ++k) t += sh[k][cx];
        zsum[e] = t;
    }
}

constexpr size_t SMEM_DUAL = sizeof(float) * 2 * (2 * A_TILE + 2 * B_TILE);
constexpr size_t SMEM_L0 = sizeof(float) * 2 * (A_TILE + 2 * B_TILE);
constexpr int TM_HALF = 64;
constexpr size_t SMEM_DUAL_H = sizeof(float) * 2 * (2 * TM_HALF * RSA + 2 * B_TILE);
constexpr size_t SMEM_L0_H = sizeof(float) * 2 * (TM_HALF * RSA + 2 * B_TILE);
constexpr size_t SMEM_SINGLE = sizeof(float) * 2 * (A_TILE + B_TILE);

bool configure() {
    static DeviceOnce once;
    if (!once.pending()) return true;
    bool r = true;
    r = r && cudaFuncSetAttribute(k_fwd<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_DUAL) == cudaSuccess;
    r = r && cudaFuncSetAttribute(k_fwd<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_L0) == cudaSuccess;
    r = r && cudaFuncSetAttribute(k_fwd<true, true, TM_HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_DUAL_H) == cudaSuccess;
    r = r && cudaFuncSetAttribute(k_fwd<true, false, TM_HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_L0_H) == cudaSuccess;
    r = r && cudaFuncSetAttribute(k_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_SINGLE) == cudaSuccess;
    r = r && cudaFuncSetAttribute(k_outer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_SINGLE) == cudaSuccess;
    if (r) once.mark();
    return r;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int layer_slices(int tiles, int max_slices) {
    int ns = 296 / tiles;
    if (ns > max_slices) ns = max_slices;
    return ns < 1 ? 1 : ns;
}

}  // namespace

size_t chain_f32_scratch_floats(const NetDesc &net, int chunk, int nslices) {
    size_t maxL = 1, sumL = 0, wts = 0;
    for (int i = 1; i <= net.K; ++i) { sumL += net.L[i]; if ((size_t)net.L[i] > maxL) maxL = net.L[i]; }
    for (int i = 0; i + 1 < net.K; ++i) wts += 4 * (size_t)net.L[i] * net.L[i + 1] + 64;      // transposed hi / lo weights
    // activations + ping-pong buffers, the same again for the TF32 remainders (Ylo, RYlo), slices, weights
    return (size_t)chunk * (2 * sumL + 6 * maxL) + (size_t)nslices * net.P + 2 * (size_t)net.P + wts + 1024;
}

void chain_f32_convert(const double *d_src, float *d_dst, size_t n, cudaStream_t st, long long *launches) {
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    k_to_float<<<blocks, 256, 0, st>>>(d_src, d_dst, n);
    ++*launches;
}

// FP32 counterpart of chain_accumulate(CHAIN_FVP): zsum (FP64) = sum_n per-sample [RGW, RGB, ...]
int chain_f32_accumulate(const NetDesc &net, const ChainScratchF32 &sc, const float *f_theta, const float *f_v,
                         const float *f_inv_var, const float *f_obs, size_t nsamples, double *d_zsum,
                         const int *d_done, const P2PComm *p2p, cudaStream_t st, long long *launches) {
    if (!configure()) return -1;
    const int K = net.K;
    // hidden layers that run on tcgen05 (tc_fwd_f32.cu): their transposed hi / lo weights are rebuilt once per FVP (v changes)
    bool tc[TRPO_MAX_LAYERS] = {};
    for (int i = 0; i + 1 < K; ++i) {
        tc[i] = tc_fwd_eligible(net.L[i], net.L[i + 1]) && (i > 0 || sc.obs_lo != nullptr);
        if (tc[i]) tc_prep_weights(f_theta + net.w_off[i], f_v + net.w_off[i], net.L[i], net.L[i + 1], sc.wt[i], st, launches);
    }
    // the last layer of a narrow-action FVP goes through the fused FP32 tail kernel (forward(K-1) + seed + backward(K))
    const int Hl = K >= 2 ? net.L[K - 1] : 0, Al = net.L[K];
    static const bool no_tail = getenv("TRPO_NO_F32_TAIL") != nullptr;
    const bool tail = !no_tail && K >= 2 && Al <= 32 && (net.ac[K] == 'l' || net.ac[K] == 'o') && (Hl % 32) == 0 && Hl >= 32 && Hl <= 256;
    int chunk_idx = 0;
    for (size_t c0 = 0; c0 < nsamples; c0 += sc.chunk, ++chunk_idx) {
        const int rows = (int)((nsamples - c0 < (size_t)sc.chunk) ? nsamples - c0 : sc.chunk);
        const int accumulate = chunk_idx > 0;
        for (int i = 0; i < (tail ? K - 1 : K); ++i) {
            const float *Yin = (i == 0) ? f_obs + c0 * net.L[0] : sc.Y[i];
            const bool last = (i == K - 1);
            if (tc[i]) {
                const int Kd = net.L[i], N = net.L[i + 1];
                const float *Ylo_in = (i == 0) ? sc.obs_lo + c0 * net.L[0] : sc.Ylo[i];
                const float *RYin = (i == 0) ? nullptr : sc.RY[i & 1], *RYlo_in = (i == 0) ? nullptr : sc.RYlo[i & 1];
                if (i > 0 && !tc[i - 1]) {          // the previous layer ran on the legacy kernels: split its outputs here
                    tc_lo_split(sc.Y[i], sc.Ylo[i], (size_t)rows * Kd, st, launches);
                    tc_lo_split(sc.RY[i & 1], sc.RYlo[i & 1], (size_t)rows * Kd, st, launches);
                }
                const bool next_tc = tc[i + 1];
                if (tc_fwd_layer(Yin, Ylo_in, RYin, RYlo_in, sc.wt[i], f_theta + net.w_off[i] + (size_t)Kd * N,
                                 f_v + net.w_off[i] + (size_t)Kd * N, rows, Kd, N, net.ac[i + 1], sc.Y[i + 1],
                                 next_tc ? sc.Ylo[i + 1] : nullptr, sc.RY[(i + 1) & 1], next_tc ? sc.RYlo[(i + 1) & 1] : nullptr,
                                 d_done, st, launches))
                    return -1;
                continue;
            }
            const bool needY = !last || net.ac[K] == 't' || net.ac[K] == 's';
            dim3 grid(cdiv(net.L[i + 1], BN), cdiv(rows, BM));
            // two 64-row CTAs per SM by default (TRPO_CHAIN_FWD_TM=128: one 128-row CTA), as in the FP64 chain
            static const bool half_tiles = !(getenv("TRPO_CHAIN_FWD_TM") && atoi(getenv("TRPO_CHAIN_FWD_TM")) == 128);
            dim3 grid_h(cdiv(net.L[i + 1], BN), cdiv(rows, TM_HALF));
            float *Yo = needY ? sc.Y[i + 1] : nullptr, *RYo = last ? nullptr : sc.RY[(i + 1) & 1], *Go = last ? sc.G[K & 1] : nullptr;
            const float *Wl = f_theta + net.w_off[i], *VWl = f_v + net.w_off[i];
            if (i == 0 && half_tiles)
                k_fwd<true, false, TM_HALF><<<grid_h, 256, SMEM_L0_H, st>>>(Yin, nullptr, Wl, VWl, rows, net.L[i], net.L[i + 1], net.ac[i + 1],
                                                                          Yo, RYo, Go, f_inv_var, d_done);
            else if (i == 0)
                k_fwd<true, false><<<grid, 512, SMEM_L0, st>>>(Yin, nullptr, Wl, VWl, rows, net.L[i], net.L[i + 1], net.ac[i + 1],
                                                                Yo, RYo, Go, f_inv_var, d_done);
            else if (half_tiles)
                k_fwd<true, true, TM_HALF><<<grid_h, 256, SMEM_DUAL_H, st>>>(Yin, sc.RY[i & 1], Wl, VWl, rows, net.L[i], net.L[i + 1],
                                                                           net.ac[i + 1], Yo, RYo, Go, f_inv_var, d_done);
            else
                k_fwd<true, true><<<grid, 512, SMEM_DUAL, st>>>(Yin, sc.RY[i & 1], Wl, VWl, rows, net.L[i], net.L[i + 1], net.ac[i + 1],
                                                                 Yo, RYo, Go, f_inv_var, d_done);
            ++*launches;
        }
        if (tail) {
            const float d3 = net.ac[K] == 'o' ? 0.1f : 1.0f;
            const float *Wl = f_theta + net.w_off[K - 1], *VWl = f_v + net.w_off[K - 1];
            float *GK = sc.G[K & 1], *Gp = sc.G[(K - 1) & 1];
            int rc;
            switch (Hl / 32) {
                case 1: rc = launch_tail_f32<1>(sc.Y[K - 1], sc.RY[(K - 1) & 1], Wl, VWl, rows, Al, net.ac[K - 1], d3, f_inv_var, GK, Gp, nullptr, d_done, st); break;
                case 2: rc = launch_tail_f32<2>(sc.Y[K - 1], sc.RY[(K - 1) & 1], Wl, VWl, rows, Al, net.ac[K - 1], d3, f_inv_var, GK, Gp, nullptr, d_done, st); break;
                case 3: rc = launch_tail_f32<3>(sc.Y[K - 1], sc.RY[(K - 1) & 1], Wl, VWl, rows, Al, net.ac[K - 1], d3, f_inv_var, GK, Gp, nullptr, d_done, st); break;
                case 4: rc = launch_tail_f32<4>(sc.Y[K - 1], sc.RY[(K - 1) & 1], Wl, VWl, rows, Al, net.ac[K - 1], d3, f_inv_var, GK, Gp, nullptr, d_done, st); break;
                case 5: rc = launch_tail_f32<5>(sc.Y[K - 1], sc.RY[(K - 1) & 1], Wl, VWl, rows, Al, net.ac[K - 1], d3, f_inv_var, GK, Gp, nullptr, d_done, st); break;
                case 6: rc = launch_tail_f32<6>(sc.Y[K - 1], sc.RY[(K - 1) & 1], Wl, VWl, rows, Al, net.ac[K - 1], d3, f_inv_var, GK, Gp, nullptr, d_done, st); break;
                case 7: rc = launch_tail_f32<7>(sc.Y[K - 1], sc.RY[(K - 1) & 1], Wl, VWl, rows, Al, net.ac[K - 1], d3, f_inv_var, GK, Gp, nullptr, d_done, st); break;
                default: rc = launch_tail_f32<8>(sc.Y[K - 1], sc.RY[(K - 1) & 1], Wl, VWl, rows, Al, net.ac[K - 1], d3, f_inv_var, GK, Gp, nullptr, d_done, st); break;
            }
            if (rc) return -1;
            ++*launches;
        }
        for (int i = K; i >= 1; --i) {
            const float *Yprev = (i == 1) ? f_obs + c0 * net.L[0] : sc.Y[i - 1];
            const int M0 = net.L[i - 1], N = net.L[i];
            const int bias_colsum = (M0 % 32) == 0;
            const int tiles_m = bias_colsum ? cdiv(M0, BM) : cdiv(M0 + 1, BM), tiles_n = cdiv(N, BN);
            const int ns = layer_slices(tiles_m * tiles_n, sc.nslices);
            dim3 go(tiles_m * tiles_n, ns);
            k_outer<<<go, NT, SMEM_SINGLE, st>>>(Yprev, sc.G[i & 1], rows, M0, N, cdiv(cdiv(rows, ns), BK) * BK, tiles_n,
                                                sc.partial, net.P, net.w_off[i - 1], accumulate, bias_colsum, d_done);
            ++*launches;
            if (i > 1 && !(tail && i == K)) {
                dim3 gb(cdiv(M0, BN), cdiv(rows, BM));
                k_bwd<<<gb, NT, SMEM_SINGLE, st>>>(sc.G[i & 1], f_theta + net.w_off[i - 1], sc.Y[i - 1], rows, N, M0,
                                                  net.ac[i - 1], sc.G[(i - 1) & 1], d_done);
                ++*launches;
            }
        }
    }
    k_reduce_f32<<<(net.P + 31) / 32, 256, 0, st>>>(sc.partial, sc.nslices, net.P, d_zsum, d_done);
    ++*launches;
    (void)p2p;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
