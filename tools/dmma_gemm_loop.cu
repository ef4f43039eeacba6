// Where does a 128 x 64 DMMA GEMM main loop lose against the bare DMMA stream? (profiles/r02_summary.md section 5)
// The single-product kernels of the GEMM chain (k_chain_outer / k_chain_bwd: 8 warps, 32 x 32 warp tiles, k-step 32, two CTAs
// per SM) keep the FP64 tensor pipe 81 - 83 % busy although the dominant stall is "math pipe throttle". This probe runs that
// main loop on tiles that never leave shared memory and adds the real kernel's ingredients one at a time:
//   mode 0  fragments + DMMAs only                      mode 1  + one block barrier per k-step
//   mode 2  + cp.async double buffering of both tiles (16-byte copies, the k_chain_outer loaders without bounds checks)
//   mode 3  mode 2 with a 3-stage ring (the barrier no longer waits for the copy issued one step earlier)
//   mode 4  the tiles arrive by TMA instead (cp.async.bulk.tensor.2d, 16-column boxes, SWIZZLE_128B, completion on an mbarrier):
//           12 copies per k-step issued by one thread instead of 12 LDGSTS per thread; fragments read through the swizzle with the
//           contraction index permuted (lane t takes k = 8*(q/2) + 2t + (q&1)) so that every half warp hits 16 distinct banks
// each with 1 and 2 CTAs per SM and with the A tile row-major ([m][k], stride 36) or natural ([k][m], stride 132).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/dmma_gemm_loop tools/dmma_gemm_loop.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

#include "../trpo-robot-control_b200/csrc/dmma_common.cuh"
#include "../trpo-robot-control_b200/csrc/tma_common.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int BM = 128, BN = 64, BK = 32, RSA = BK + 4, RSB = BN + 4, RSN = BM + 4, NT = 256;
constexpr int A_TILE = BM * RSA, B_TILE = BK * RSB;       // natural A: BK * RSN = 4224 <= 4608
constexpr int STAGE = A_TILE + B_TILE;

template <bool ANAT>
__device__ __forceinline__ void mma_stage(double (&acc)[4][4][2], const double *As, const double *Bs, int wm, int wn, int g, int t) {
#pragma unroll
    for (int q = 0; q < BK / 4; ++q) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = ANAT ? As[(4 * q + t) * RSN + 32 * wm + 8 * i + g] : As[(32 * wm + 8 * i + g) * RSA + 4 * q + t];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[(4 * q + t) * RSB + 32 * wn + 8 * j + g];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
    }
}

template <bool ANAT>
__device__ __forceinline__ void load_stage(double *As, double *Bs, const double *gA, const double *gB, int ks, int tid) {
    if (ANAT) {
#pragma unroll
        for (int it = 0; it < BK * BM / 2 / NT; ++it) {
            const int idx = tid + it * NT, k = idx >> 6, m = (idx & 63) * 2;
            cp_async16(&As[k * RSN + m], gA + (size_t)(ks + k) * BM + m, 16);
        }
    } else {
#pragma unroll
        for (int it = 0; it < BM * BK / 2 / NT; ++it) {
            const int idx = tid + it * NT, m = idx >> 4, k = (idx & 15) * 2;
            cp_async16(&As[m * RSA + k], gA + (size_t)m * 4096 + ks + k, 16);
        }
    }
#pragma unroll
    for (int it = 0; it < BK * BN / 2 / NT; ++it) {
        const int idx = tid + it * NT, k = idx >> 5, n = (idx & 31) * 2;
        cp_async16(&Bs[k * RSB + n], gB + (size_t)(ks + k) * BN + n, 16);
    }
    cp_async_commit();
}

template <int MODE, bool ANAT>
__global__ void __launch_bounds__(NT, 2) k_loop(const double *gA, const double *gB, double *out, int nk) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3, wm = w >> 1, wn = w & 1;
    constexpr int NS = MODE == 3 ? 3 : 2;
    for (int i = tid; i < NS * STAGE; i += NT) smem[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double acc[4][4][2] = {};
    gA += (size_t)blockIdx.x * 512;                       // 4096-sample slices of an L2-resident buffer
    gB += (size_t)blockIdx.x * 64;
    if (MODE >= 2) {
        for (int s = 0; s < NS - 1; ++s) load_stage<ANAT>(smem + s * STAGE, smem + s * STAGE + A_TILE, gA, gB, (s * BK) & 4095, tid);
    }
    for (int it = 0; it < nk; ++it) {
        if (MODE >= 2) { if (MODE == 3) cp_async_wait_group<1>(); else cp_async_wait_group<0>(); }
        if (MODE >= 1) __syncthreads(); else asm volatile("" ::: "memory");   // mode 0: the fragments are re-read every k-step
        if (MODE >= 2) {
            const int st = (it + NS - 1) % NS;
            load_stage<ANAT>(smem + st * STAGE, smem + st * STAGE + A_TILE, gA, gB, ((it + NS - 1) * BK) & 4095, tid);
        }
        const double *As = smem + (MODE >= 2 ? it % NS : 0) * STAGE, *Bs = As + A_TILE;
        mma_stage<ANAT>(acc, As, Bs, wm, wn, g, t);
    }
    if (MODE >= 2) cp_async_wait_group<0>();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            out[((size_t)(blockIdx.x * NT + tid) * 16 + i * 4 + j) * 2] = acc[i][j][0];
            out[((size_t)(blockIdx.x * NT + tid) * 16 + i * 4 + j) * 2 + 1] = acc[i][j][1];
        }
}


// ---- mode 4: TMA-fed stages ------------------------------------------------------------------------------------------
// stage = 8 A boxes (m-blocks of 16 columns) + 4 B boxes (n-blocks of 16), each 32 k-rows x 128 bytes, SWIZZLE_128B:
// byte (k, col) of a box = k*128 + (((col >> 1) ^ (k & 7)) << 4) + (col & 1) * 8
constexpr int BOX_BYTES = BK * 128, TMA_STAGE_BYTES = (BM / 16 + BN / 16) * BOX_BYTES;
__global__ void __launch_bounds__(NT, 2) k_loop_tma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                    double *out, int nk) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full[2];
    unsigned char *smem_raw = smem_dyn + ((1024u - (tma::smem_u32(smem_dyn) & 1023u)) & 1023u);      // swizzle atoms are 1 KB
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3, wm = w >> 1, wn = w & 1;
    if (tid == 0) {
        tma::mbar_init(&full[0], 1); tma::mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double acc[4][4][2] = {};
    // lane offsets through the swizzle: [c = q & 1][sub-block = i & 1]
    int off[2][2];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int ib = 0; ib < 2; ++ib) off[c][ib] = (2 * t + c) * 128 + (((4 * ib + (g >> 1)) ^ (2 * t + c)) << 4) + (g & 1) * 8;
    auto issue = [&](int st, int ks) {
        unsigned char *base = smem_raw + st * TMA_STAGE_BYTES;
        tma::mbar_expect_tx(&full[st], TMA_STAGE_BYTES);
#pragma unroll
        for (int mb = 0; mb < BM / 16; ++mb) tma::tma_load_2d(base + mb * BOX_BYTES, &mapA, &full[st], 16 * mb, 4 * blockIdx.x + ks);
#pragma unroll
        for (int nb = 0; nb < BN / 16; ++nb) tma::tma_load_2d(base + (BM / 16 + nb) * BOX_BYTES, &mapB, &full[st], 16 * nb, blockIdx.x + ks);
    };
    if (tid == 0) issue(0, 0);
    for (int it = 0; it < nk; ++it) {
        __syncthreads();                                   // everybody is done with the other stage
        if (tid == 0 && it + 1 < nk) issue((it + 1) & 1, ((it + 1) * BK) & 4095);
        tma::mbar_wait(&full[it & 1], (it >> 1) & 1);
        const unsigned char *A = smem_raw + (it & 1) * TMA_STAGE_BYTES + 2 * wm * BOX_BYTES;
        const unsigned char *B = smem_raw + (it & 1) * TMA_STAGE_BYTES + (BM / 16 + 2 * wn) * BOX_BYTES;
#pragma unroll
        for (int q = 0; q < BK / 4; ++q) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const double *>(A + (i >> 1) * BOX_BYTES + (q >> 1) * 1024 + off[q & 1][i & 1]);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double *>(B + (j >> 1) * BOX_BYTES + (q >> 1) * 1024 + off[q & 1][j & 1]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            out[((size_t)(blockIdx.x * NT + tid) * 16 + i * 4 + j) * 2] = acc[i][j][0];
            out[((size_t)(blockIdx.x * NT + tid) * 16 + i * 4 + j) * 2 + 1] = acc[i][j][1];
        }
}


// layout probe: one box of a matrix whose element (row, col) holds row * 1000 + col, dumped as it lies in shared memory
__global__ void k_dump_box(const __grid_constant__ CUtensorMap map, double *out) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full;
    unsigned char *base = smem_dyn + ((1024u - (tma::smem_u32(smem_dyn) & 1023u)) & 1023u);
    if (threadIdx.x == 0) {
        tma::mbar_init(&full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma::mbar_expect_tx(&full, BOX_BYTES);
        tma::tma_load_2d(base, &map, &full, 16, 3);
    }
    __syncthreads();
    tma::mbar_wait(&full, 0);
    for (int i = threadIdx.x; i < BOX_BYTES / 8; i += blockDim.x) out[i] = reinterpret_cast<const double *>(base)[i];
}

bool make_map_f64(CUtensorMap *m, const double *base, size_t nrows, int ncols) {
    tma::EncodeTiledFn fn = tma::encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)ncols, (cuuint64_t)nrows};
    const cuuint64_t strides[1] = {(cuuint64_t)ncols * sizeof(double)};
    const cuuint32_t box[2] = {16, (cuuint32_t)BK};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void run_tma(const double *gA, size_t rowsA, const double *gB, size_t rowsB, double *out, int ctas_per_sm, int nk) {
    CUtensorMap mA, mB;
    if (!make_map_f64(&mA, gA, rowsA, BM) || !make_map_f64(&mB, gB, rowsB, BN)) { printf("tensor map encode failed\n"); return; }
    const size_t smem = 2 * TMA_STAGE_BYTES + 1024;
    CK(cudaFuncSetAttribute(k_loop_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_loop_tma<<<grid, NT, smem>>>(mA, mB, out, nk);
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        k_loop_tma<<<grid, NT, smem>>>(mA, mB, out, nk);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    const double flop = 2.0 * BM * BN * BK * (double)nk * grid;
    printf("%-58s %d CTA/SM  %8.3f ms  %6.2f TFLOP/s\n", "mode 4 TMA boxes (SWIZZLE_128B) + mbarrier, A [k][m]", ctas_per_sm, best, flop / best / 1e9);
}

template <int MODE, bool ANAT>
void run(const char *name, const double *gA, const double *gB, double *out, int ctas_per_sm) {
    constexpr int NS = MODE == 3 ? 3 : 2;
    const size_t smem = sizeof(double) * NS * STAGE;
    CK(cudaFuncSetAttribute(k_loop<MODE, ANAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nk = 4096, grid = 148 * ctas_per_sm;
    if (NS == 3 && ctas_per_sm == 2) { printf("%-58s  (3 stages x 2 CTAs do not fit)\n", name); return; }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_loop<MODE, ANAT><<<grid, NT, smem>>>(gA, gB, out, nk);
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        k_loop<MODE, ANAT><<<grid, NT, smem>>>(gA, gB, out, nk);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    const double flop = 2.0 * BM * BN * BK * (double)nk * grid;
    printf("%-58s %d CTA/SM  %8.3f ms  %6.2f TFLOP/s\n", name, ctas_per_sm, best, flop / best / 1e9);
}

int main() {
    double *gA, *gB, *out;
    const size_t nA = (size_t)4096 * 4096 + 148 * 2 * 512 + 4096, nB = (size_t)4200 * 64 + 148 * 2 * 64;
    const size_t nOut = (size_t)148 * 2 * NT * 32;
    CK(cudaMalloc(&gA, nA * 8)); CK(cudaMalloc(&gB, nB * 8)); CK(cudaMalloc(&out, nOut * 8));
    {   // small integers: every product and partial sum is exact, so the two loaders must agree bitwise whatever the k order
        double *h = (double *)malloc((nA > nB ? nA : nB) * 8);
        unsigned x = 12345u;
        for (size_t i = 0; i < nA; ++i) { x = x * 1664525u + 1013904223u; h[i] = (double)((int)(x >> 28) - 8); }
        CK(cudaMemcpy(gA, h, nA * 8, cudaMemcpyHostToDevice));
        for (size_t i = 0; i < nB; ++i) { x = x * 1664525u + 1013904223u; h[i] = (double)((int)(x >> 29) - 4); }
        CK(cudaMemcpy(gB, h, nB * 8, cudaMemcpyHostToDevice));
        free(h);
    }
    if (getenv("DUMP_BOX")) {
        double *h = (double *)malloc(nA * 8);
        for (size_t r = 0; r < 64; ++r) for (int c = 0; c < BM; ++c) h[r * BM + c] = r * 1000.0 + c;
        CK(cudaMemcpy(gA, h, 64 * BM * 8, cudaMemcpyHostToDevice));
        CUtensorMap mA;
        make_map_f64(&mA, gA, nA / BM, BM);
        k_dump_box<<<1, 128, BOX_BYTES + 1024>>>(mA, out);
        CK(cudaMemcpy(h, out, BOX_BYTES, cudaMemcpyDeviceToHost));
        for (int r = 0; r < 10; ++r) { for (int c = 0; c < 16; ++c) printf("%6.0f ", h[r * 16 + c]); printf("\n"); }
        return 0;
    }
    for (int c = 1; c <= 2 && !getenv("CHECK_ONLY"); ++c) {
        run<0, false>("mode 0 fragments + DMMA, A [m][k]", gA, gB, out, c);
        run<0, true >("mode 0 fragments + DMMA, A [k][m]", gA, gB, out, c);
        run<1, false>("mode 1 + barrier per k-step, A [m][k]", gA, gB, out, c);
        run<1, true >("mode 1 + barrier per k-step, A [k][m]", gA, gB, out, c);
        run<2, false>("mode 2 + cp.async double buffer, A [m][k]", gA, gB, out, c);
        run<2, true >("mode 2 + cp.async double buffer, A [k][m]", gA, gB, out, c);
        run<3, false>("mode 3 + cp.async 3-stage ring, A [m][k]", gA, gB, out, c);
        run<3, true >("mode 3 + cp.async 3-stage ring, A [k][m]", gA, gB, out, c);
        run_tma(gA, nA / BM, gB, nB / BN, out, c, 4096);
    }
    for (int nkc : {1, 2, 128}) {   // the TMA-fed loop against the cp.async one on the same data (no wrap-around of the sample index)
        double *h2 = (double *)malloc(nOut * 8), *h4 = (double *)malloc(nOut * 8);
        const size_t smem2 = sizeof(double) * 2 * STAGE;
        CK(cudaFuncSetAttribute(k_loop<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        CK(cudaFuncSetAttribute(k_loop_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TMA_STAGE_BYTES + 1024));
        k_loop<2, true><<<296, NT, smem2>>>(gA, gB, out, nkc);
        CK(cudaGetLastError());
        CK(cudaMemcpy(h2, out, nOut * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemset(out, 0, nOut * 8));
        CUtensorMap mA, mB;
        make_map_f64(&mA, gA, nA / BM, BM); make_map_f64(&mB, gB, nB / BN, BN);
        k_loop_tma<<<296, NT, 2 * TMA_STAGE_BYTES + 1024>>>(mA, mB, out, nkc);
        CK(cudaGetLastError());
        CK(cudaMemcpy(h4, out, nOut * 8, cudaMemcpyDeviceToHost));
        size_t bad = 0; double mx = 0;
        for (size_t i = 0; i < nOut; ++i) { if (h2[i] != h4[i]) ++bad; if (fabs(h2[i]) > mx) mx = fabs(h2[i]); }
        printf("check nk=%d: TMA loop vs cp.async loop, %zu of %zu accumulators differ (max |value| %.0f)\n", nkc, bad, nOut, mx);
        if (getenv("CHECK_ONLY")) { for (int i = 0; i < 12; ++i) printf("  %g/%g", h2[i], h4[i]); printf("\n"); }
        free(h2); free(h4);
    }
    return 0;
}
