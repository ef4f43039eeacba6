// tc_fwd_f32.cu -- FP32 mode, forward R-op layer on the 5th-generation tensor cores (tcgen05 / TMEM / TMA).
//
// One hidden layer of the combined forward pass of the Fisher-vector product (TRPO_FVP.c:783-836) for a 128-row block of
// samples and ALL N <= 256 outputs of the layer:
//     X  = Y W + B                      Y_out  = f(X)
//     RX = RY W + Y VW + VB             RY_out = RX f'(X)
// as 3xTF32 products with FP32 accumulation: every FP32 operand is x = hi + lo with hi = x truncated to TF32 (what the
// tensor core reads from the FP32 bits) and lo = x - hi kept in a second array, and a product a*b is issued as
// a_hi*b_hi + a_hi*b_lo + a_lo*b_hi. The lo arrays are produced where the data is produced (this kernel's epilogue for the
// activations, one small kernel per FVP for the transposed weights, one pass per batch for the observations), so the main
// loop is pure data movement + MMA:
//   warp 0   TMA producer: cp.async.bulk.tensor (SASS UTMALDG) of the operand tiles of one k-block (16 columns = 64 bytes,
//            SWIZZLE_64B) into a ring of shared-memory stages, completion on an mbarrier (complete_tx)
//   warp 1   MMA issuer: one thread issues tcgen05.mma.cta_group::1.kind::tf32 (SASS UTCHMMA / UTCMMA), M = 128, N = BN, K = 8,
//            9 per k-step (6 for layer 0, where RY = 0) into two TMEM accumulators (X and RX, BN columns each);
//            tcgen05.commit releases the stage to the producer and finally hands the accumulators to the epilogue
//   warps 2-5  epilogue: tcgen05.ld (SASS LDTM) of 32 rows x 32 columns per warp and step, bias, activation, R-op scaling,
//            hi/lo split, 128-byte row segments straight to global memory.
// Edges need no code: TMA zero-fills rows past the chunk, columns past Kd (376 is not a multiple of 16) and weight rows past N.
// The legacy path (gemm_chain_f32.cu: mma.sync HMMA.1688, 278 TFLOP/s) stays for the layers this kernel does not take
// (last layer, Kd not a multiple of 4, N > 256) and for the backward / outer-product GEMMs.
#include <cuda.h>
#include <stdint.h>
#include <stdlib.h>

#include "trpo_internal.cuh"
#include "tmem_scratch.cuh"
#include "tma_common.cuh"

namespace {
using namespace tma;

constexpr int TC_BM = 128, TC_BK = 16;
constexpr int TC_EPI_WARPS = 16, TC_THREADS = 64 + 32 * TC_EPI_WARPS;       // warp 0: TMA, warp 1: MMA, 16 epilogue warps

// ---- PTX helpers (TMA / mbarrier ones in tma_common.cuh) -------------------------------------------------------------
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory");
}
// K-major operand tile, rows of 64 bytes, SWIZZLE_64B: 8-row groups 512 bytes apart (SBO), LBO unused (1), descriptor version 1
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address
    d |= (uint64_t)1 << 16;                             // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(512 >> 4) << 32;                    // stride byte offset
    d |= (uint64_t)1 << 46;                             // version = 1 (sm_100)
    d |= (uint64_t)4 << 61;                             // layout type SWIZZLE_64B
    return d;
}
// D (TMEM) (+)= A (smem) * B (smem), kind::tf32; accumulate == 0 overwrites D
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float tc_act(char a, float x, float &d) {
    if (a == 't') {                                     // tanh = 1 - 2 / (exp(2x) + 1): ex2.approx + rcp.approx, ~1e-7 absolute
        const float e = __expf(2.0f * x);
        const float y = 1.0f - __fdividef(2.0f, e + 1.0f);
        d = 1.0f - y * y;
        return y;
    }
    if (a == 's') { const float y = __fdividef(1.0f, 1.0f + __expf(-x)); d = y * (1.0f - y); return y; }
    if (a == 'o') { d = 0.1f; return 0.1f * x; }
    d = 1.0f;
    return x;
}
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

template <int BN, bool HAS_RA>
struct TcCfg {
    static constexpr int NA = HAS_RA ? 4 : 2;                                    // Y, Ylo [, RY, RYlo]
    static constexpr int A_BYTES = TC_BM * TC_BK * 4, B_BYTES = BN * TC_BK * 4;   // 8 KB, BN * 64 B
    static constexpr int STAGE_BYTES = NA * A_BYTES + 4 * B_BYTES;               // + W, Wlo, VW, VWlo
    static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 6 ? 6 : (200 * 1024 / STAGE_BYTES);
    static constexpr int EPI_BYTES = TC_EPI_WARPS * 2 * 32 * 33 * 4;            // epilogue staging tiles reuse the operand ring
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES > EPI_BYTES ? STAGES * STAGE_BYTES : EPI_BYTES;
    static constexpr size_t SMEM = (size_t)RING_BYTES + 1024 /* alignment slack */ + 128 /* barriers */ + 2048 /* bias rows */;
    static constexpr uint32_t TM_COLS = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
    static_assert(STAGES >= 2, "pipeline depth");
    static_assert(BN % 32 == 0 && BN <= 256, "tile width");
};

struct TcMaps { CUtensorMap y, ylo, ry, rylo, w, wlo, vw, vwlo; };

// CL > 1: thread-block cluster of CL CTAs (consecutive 128-row blocks). Every CTA needs the SAME weight tiles, and with one CTA
// per SM all 148 of them ask the L2 for the same few cache lines at the same time (measured: 3.4 TB/s of TMA traffic, two thirds
// of it weights, the tensor pipe 29 % busy): each CTA loads 1/CL of every weight tile and TMA multicasts it to the whole cluster.
// A stage may be refilled only when the MMAs of ALL CTAs of the cluster have read it: the empty barriers count CL commits.
template <int BN, bool HAS_RA, int CL>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_fwd(const __grid_constant__ TcMaps maps, const float *__restrict__ bias,
                                                          const float *__restrict__ vbias, int rows, int Kd, int N, char act,
                                                          float *__restrict__ Yout, float *__restrict__ Ylo_out,
                                                          float *__restrict__ RYout, float *__restrict__ RYlo_out,
                                                          const int *__restrict__ done, int dbg) {
    if (done && *done) return;
    using C = TcCfg<BN, HAS_RA>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);       // swizzled tiles: 1024-byte aligned
    uint64_t *full = (uint64_t *)(smem + (size_t)C::RING_BYTES);                         // [STAGES] TMA -> MMA
    uint64_t *empty = full + C::STAGES;                                                  // [STAGES] MMA -> TMA
    uint64_t *acc_ready = empty + C::STAGES;                                             // MMA -> epilogue
    uint32_t *tmem_slot = (uint32_t *)(acc_ready + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_BM;
    const int nkb = (Kd + TC_BK - 1) / TC_BK;

    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.y); tma_prefetch_desc(&maps.ylo); tma_prefetch_desc(&maps.w); tma_prefetch_desc(&maps.wlo);
        tma_prefetch_desc(&maps.vw); tma_prefetch_desc(&maps.vwlo);
        if (HAS_RA) { tma_prefetch_desc(&maps.ry); tma_prefetch_desc(&maps.rylo); }
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }
        mbar_init(acc_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TM_COLS);
    {   // bias rows of W and VW, read by every epilogue thread: shared memory, zero past N
        float *sb = reinterpret_cast<float *>(smem + (size_t)C::RING_BYTES + 128);
        for (int i = threadIdx.x; i < 512; i += TC_THREADS) {
            const int n = i & 255;
            sb[i] = n < N ? (i < 256 ? bias[n] : vbias[n]) : 0.0f;
        }
    }
    tmem_fence_before_sync();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                       // every CTA's barriers exist before a peer's TMA / commit can reach them
    tmem_fence_after_sync();
    const uint32_t tm = *tmem_slot;                       // X at columns [0, BN), RX at [BN, 2 BN)

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % C::STAGES, it = kb / C::STAGES;
                if (it > 0) mbar_wait(&empty[s], (it - 1) & 1);                // the MMAs that read this stage have completed
                uint8_t *st = smem + (size_t)s * C::STAGE_BYTES;
                mbar_expect_tx(&full[s], C::STAGE_BYTES);
                const int k0 = kb * TC_BK;
                tma_load_2d(st + 0 * C::A_BYTES, &maps.y, &full[s], k0, m0);
                tma_load_2d(st + 1 * C::A_BYTES, &maps.ylo, &full[s], k0, m0);
                if (HAS_RA) {
                    tma_load_2d(st + 2 * C::A_BYTES, &maps.ry, &full[s], k0, m0);
                    tma_load_2d(st + 3 * C::A_BYTES, &maps.rylo, &full[s], k0, m0);
                }
                uint8_t *sb = st + C::NA * C::A_BYTES;
                if (CL == 1) {
                    tma_load_2d(sb + 0 * C::B_BYTES, &maps.w, &full[s], k0, 0);
                    tma_load_2d(sb + 1 * C::B_BYTES, &maps.wlo, &full[s], k0, 0);
                    tma_load_2d(sb + 2 * C::B_BYTES, &maps.vw, &full[s], k0, 0);
                    tma_load_2d(sb + 3 * C::B_BYTES, &maps.vwlo, &full[s], k0, 0);
                } else {
                    // this CTA's slice (BN / CL weight rows) of each of the four tiles, delivered to every CTA of the cluster
                    constexpr int SL = BN / CL;
                    const int n0 = (int)crank * SL;
                    uint8_t *sl = sb + (size_t)n0 * TC_BK * 4;
                    tma_load_2d_mc(sl + 0 * C::B_BYTES, &maps.w, &full[s], k0, n0, CMASK);
                    tma_load_2d_mc(sl + 1 * C::B_BYTES, &maps.wlo, &full[s], k0, n0, CMASK);
                    tma_load_2d_mc(sl + 2 * C::B_BYTES, &maps.vw, &full[s], k0, n0, CMASK);
                    tma_load_2d_mc(sl + 3 * C::B_BYTES, &maps.vwlo, &full[s], k0, n0, CMASK);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            const uint32_t tmX = tm, tmRX = tm + BN;
            uint32_t accX = 0, accRX = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % C::STAGES, it = kb / C::STAGES;
                mbar_wait(&full[s], it & 1);
                tmem_fence_after_sync();
                const uint32_t sa = smem_u32(smem + (size_t)s * C::STAGE_BYTES), sbb = sa + C::NA * C::A_BYTES;
#pragma unroll
                for (int kk = 0; kk < TC_BK / 8; ++kk) {
                    const uint32_t ko = kk * 32;                               // 8 TF32 = 32 bytes along K inside the 64-byte rows
                    const uint64_t dY = umma_desc_k_sw64(sa + 0 * C::A_BYTES + ko), dYl = umma_desc_k_sw64(sa + 1 * C::A_BYTES + ko);
                    const uint64_t dW = umma_desc_k_sw64(sbb + 0 * C::B_BYTES + ko), dWl = umma_desc_k_sw64(sbb + 1 * C::B_BYTES + ko);
                    const uint64_t dV = umma_desc_k_sw64(sbb + 2 * C::B_BYTES + ko), dVl = umma_desc_k_sw64(sbb + 3 * C::B_BYTES + ko);
                    // X = Y W: compensation terms first
                    if (dbg & 8) continue;
                    umma_tf32(tmX, dYl, dW, idesc, accX); accX = 1;
                    umma_tf32(tmX, dY, dWl, idesc, 1);
                    umma_tf32(tmX, dY, dW, idesc, 1);
                    // RX = Y VW (+ RY W)
                    umma_tf32(tmRX, dYl, dV, idesc, accRX); accRX = 1;
                    umma_tf32(tmRX, dY, dVl, idesc, 1);
                    umma_tf32(tmRX, dY, dV, idesc, 1);
                    if (HAS_RA) {
                        const uint64_t dR = umma_desc_k_sw64(sa + 2 * C::A_BYTES + ko), dRl = umma_desc_k_sw64(sa + 3 * C::A_BYTES + ko);
                        umma_tf32(tmRX, dRl, dW, idesc, 1);
                        umma_tf32(tmRX, dR, dWl, idesc, 1);
                        umma_tf32(tmRX, dR, dW, idesc, 1);
                    }
                }
                if (CL == 1) umma_commit(&empty[s]);                           // arrives when the MMAs above have read the stage
                else umma_commit_mc(&empty[s], CMASK);                         // ... on the empty barrier of every CTA of the cluster
            }
            umma_commit(acc_ready);                                            // ... and when every MMA of the tile has completed
        }
    } else {
        // ===================== epilogue: 16 warps; TMEM lane quadrant = warp % 4, column group = (warp - 2) / 4 =====================
        // tcgen05.ld hands every thread ONE ROW of the accumulator tile (32 consecutive columns per step). Stored from there a
        // warp's store instruction would touch 32 different 128-byte lines, so each 32 x 32 block goes through a padded per-warp
        // shared-memory tile (the operand ring is free by now) and leaves as whole 128-byte rows; the TF32 remainders are formed
        // on the way out. Measured with 4 epilogue warps: 1.0 of the 1.6 ms the two hidden layers took at 200 k states was this
        // epilogue, one warp per scheduler exposing every latency -- hence 16 warps (4 per scheduler).
        const int q = warp & 3, cg = (warp - 2) >> 2;
        float *sbias = reinterpret_cast<float *>(smem + (size_t)C::RING_BYTES + 128);   // [2][256], after the barriers
        mbar_wait(acc_ready, 0);
        tmem_fence_after_sync();
        const uint32_t tq = tm + ((uint32_t)(32 * q) << 16);
        float *tile = reinterpret_cast<float *>(smem) + (size_t)(warp - 2) * (2 * 32 * 33);         // [y, ry][32 rows][33]
        const int rows_here = min(32, rows - (m0 + 32 * q));                             // rows of this warp inside the chunk (may be <= 0)
        for (int c0 = 32 * cg; c0 < BN && c0 < N; c0 += 32 * (TC_EPI_WARPS / 4)) {
            if (dbg & 4) break;
            uint32_t xr[32], rr[32];
            tmem_ld32(tq + c0, xr);
            tmem_ld32(tq + BN + c0, rr);
            if (dbg & 2) continue;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float d;
                const float y = tc_act(act, __uint_as_float(xr[j]) + sbias[c0 + j], d);
                tile[(0 * 32 + lane) * 33 + j] = y;
                tile[(1 * 32 + lane) * 33 + j] = (__uint_as_float(rr[j]) + sbias[256 + c0 + j]) * d;
            }
            __syncwarp();
            if (c0 + lane < N && !(dbg & 1)) {
                for (int r = 0; r < rows_here; ++r) {
                    const size_t o = (size_t)(m0 + 32 * q + r) * N + c0 + lane;
                    const float y = tile[(0 * 32 + r) * 33 + lane], ry = tile[(1 * 32 + r) * 33 + lane];
                    Yout[o] = y;
                    RYout[o] = ry;
                    if (Ylo_out) { Ylo_out[o] = tf32_lo(y); RYlo_out[o] = tf32_lo(ry); }
                }
            }
            __syncwarp();
        }
    }
    tmem_fence_before_sync();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                       // no CTA leaves while a peer may still multicast into it or signal it
    if (warp == 1) tmem_dealloc(tm, C::TM_COLS);
}

// Wt[n][k] = W[k][n] (hi: truncated to TF32, lo: the remainder), for the weight matrix and the direction of one layer
__global__ void k_tc_prep_weights(const float *__restrict__ W, const float *__restrict__ VW, int Kd, int N,
                                  float *__restrict__ Wt, float *__restrict__ Wtlo, float *__restrict__ VWt, float *__restrict__ VWtlo) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Kd * N) return;
    const int n = idx / Kd, k = idx % Kd;
    const float w = W[(size_t)k * N + n], v = VW[(size_t)k * N + n];
    const float wh = __uint_as_float(__float_as_uint(w) & 0xffffe000u), vh = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    Wt[idx] = wh; Wtlo[idx] = w - wh; VWt[idx] = vh; VWtlo[idx] = v - vh;
}
__global__ void k_tc_lo(const float *__restrict__ x, float *__restrict__ lo, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) lo[i] = tf32_lo(x[i]);
}

// row-major FP32 matrix [nrows x ncols] (row stride ld floats), box = box_rows x 16 columns, SWIZZLE_64B, zero fill
bool make_map(CUtensorMap *m, const float *base, int nrows, int ncols, int ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)ncols, (cuuint64_t)nrows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, bool HAS_RA, int CL>
int launch_tc(const TcMaps &maps, const float *bias, const float *vbias, int rows, int Kd, int N, char act, float *Yout,
              float *Ylo_out, float *RYout, float *RYlo_out, const int *done, cudaStream_t st) {
    using C = TcCfg<BN, HAS_RA>;
    static DeviceOnce once;
    if (once.pending()) {
        if (cudaFuncSetAttribute(k_tc_fwd<BN, HAS_RA, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess) return -1;
        once.mark();
    }
    const int tiles = (rows + TC_BM - 1) / TC_BM;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((tiles + CL - 1) / CL * CL);       // whole clusters: surplus CTAs see only zero-filled rows and store nothing
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = C::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k_tc_fwd<BN, HAS_RA, CL>, maps, bias, vbias, rows, Kd, N, act, Yout, Ylo_out, RYout, RYlo_out, done, getenv("TRPO_TC_DEBUG") ? atoi(getenv("TRPO_TC_DEBUG")) : 0) == cudaSuccess ? 0 : -1;
}
template <int BN, bool HAS_RA>
int launch_tc_cl(int cl, const TcMaps &maps, const float *bias, const float *vbias, int rows, int Kd, int N, char act, float *Yout,
                 float *Ylo_out, float *RYout, float *RYlo_out, const int *done, cudaStream_t st) {
    if (cl == 4) return launch_tc<BN, HAS_RA, 4>(maps, bias, vbias, rows, Kd, N, act, Yout, Ylo_out, RYout, RYlo_out, done, st);
    if (cl == 2) return launch_tc<BN, HAS_RA, 2>(maps, bias, vbias, rows, Kd, N, act, Yout, Ylo_out, RYout, RYlo_out, done, st);
    return launch_tc<BN, HAS_RA, 1>(maps, bias, vbias, rows, Kd, N, act, Yout, Ylo_out, RYout, RYlo_out, done, st);
}

}  // namespace

bool tc_fwd_eligible(int Kd, int N) {
    static const bool off = getenv("TRPO_NO_TCGEN05") != nullptr;
    return !off && encode_fn() != nullptr && (Kd % 4) == 0 && Kd >= 16 && (N % 32) == 0 && N >= 32 && N <= 256;
}

void tc_lo_split(const float *x, float *lo, size_t n, cudaStream_t st, long long *launches) {
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    k_tc_lo<<<blocks, 256, 0, st>>>(x, lo, n);
    ++*launches;
}

void tc_prep_weights(const float *W, const float *VW, int Kd, int N, float *wt4, cudaStream_t st, long long *launches) {
    const size_t sz = (size_t)Kd * N;
    k_tc_prep_weights<<<(int)((sz + 255) / 256), 256, 0, st>>>(W, VW, Kd, N, wt4, wt4 + sz, wt4 + 2 * sz, wt4 + 3 * sz);
    ++*launches;
}

// One forward layer. Yin / Ylo_in: [rows x Kd]; RYin / RYlo_in: the same or NULL (layer 0); wt4: the four transposed weight
// arrays of tc_prep_weights; bias / vbias: row Kd of W and VW (N floats). Ylo_out / RYlo_out may be NULL (next layer is not ours).
int tc_fwd_layer(const float *Yin, const float *Ylo_in, const float *RYin, const float *RYlo_in, const float *wt4, const float *bias,
                 const float *vbias, int rows, int Kd, int N, char act, float *Yout, float *Ylo_out, float *RYout, float *RYlo_out,
                 const int *done, cudaStream_t st, long long *launches) {
    const int BN = N <= 64 ? 64 : N <= 128 ? 128 : 256;
    // cluster size: weight tiles are multicast over CL consecutive row blocks (TRPO_TC_CLUSTER = 1 | 2 | 4 overrides)
    static const int cl_env = getenv("TRPO_TC_CLUSTER") ? atoi(getenv("TRPO_TC_CLUSTER")) : 0;
    int cl = (cl_env == 1 || cl_env == 2 || cl_env == 4) ? cl_env : 4;
    if ((rows + TC_BM - 1) / TC_BM < 2 * cl) cl = 1;
    const size_t sz = (size_t)Kd * N;
    TcMaps maps;
    bool ok = make_map(&maps.y, Yin, rows, Kd, Kd, TC_BM) && make_map(&maps.ylo, Ylo_in, rows, Kd, Kd, TC_BM) &&
              make_map(&maps.w, wt4, N, Kd, Kd, BN / cl) && make_map(&maps.wlo, wt4 + sz, N, Kd, Kd, BN / cl) &&
              make_map(&maps.vw, wt4 + 2 * sz, N, Kd, Kd, BN / cl) && make_map(&maps.vwlo, wt4 + 3 * sz, N, Kd, Kd, BN / cl);
    if (RYin) ok = ok && make_map(&maps.ry, RYin, rows, Kd, Kd, TC_BM) && make_map(&maps.rylo, RYlo_in, rows, Kd, Kd, TC_BM);
    else { maps.ry = maps.y; maps.rylo = maps.ylo; }
    if (!ok) return -1;
    int rc;
    if (RYin) rc = BN == 64 ? launch_tc_cl<64, true>(cl, maps, bias, vbias, rows, Kd, N, act, Yout, Ylo_out, RYout, RYlo_out, done, st)
                : BN == 128 ? launch_tc_cl<128, true>(cl, maps, bias, vbias, rows, Kd, N, act, Yout, Ylo_out, RYout, RYlo_out, done, st)
                            : launch_tc_cl<256, true>(cl, maps, bias, vbias, rows, Kd, N, act, Yout, Ylo_out, RYout, RYlo_out, done, st);
    else rc = BN == 64 ? launch_tc_cl<64, false>(cl, maps, bias, vbias, rows, Kd, N, act, Yout, Ylo_out, RYout, RYlo_out, done, st)
            : BN == 128 ? launch_tc_cl<128, false>(cl, maps, bias, vbias, rows, Kd, N, act, Yout, Ylo_out, RYout, RYlo_out, done, st)
                        : launch_tc_cl<256, false>(cl, maps, bias, vbias, rows, Kd, N, act, Yout, Ylo_out, RYout, RYlo_out, done, st);
    if (rc == 0) ++*launches;
    return rc;
}
