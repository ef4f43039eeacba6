// FP64 peak micro-benchmark for B200 (sm_100a): DFMA pipe, DMMA (mma.sync f64) shapes, and a mix.
// Used only to measure the roofline denominator for the FVP kernels (MEASURED_PEAKS.json has no FP64 entry).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dmma884(double *out, int iters, double a, double b) {
    double c0[CHAINS], c1[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dmma1684(double *out, int iters, double a, double b) {
    double c[CHAINS][4];
    double af[2] = {a, a + 1};
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) dmma1684(c[i], af, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dmma1688(double *out, int iters, double a, double b) {
    double c[CHAINS][4];
    double af[4] = {a, a + 1, a + 2, a + 3};
    double bf[2] = {b, b + 1};
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) dmma1688(c[i], af, bf);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dmma16816(double *out, int iters, double a, double b) {
    double c[CHAINS][4];
    double af[8] = {a, a + 1, a + 2, a + 3, a + 4, a + 5, a + 6, a + 7};
    double bf[4] = {b, b + 1, b + 2, b + 3};
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) dmma16816(c[i], af, bf);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA stream with distinct A/B operand registers per instruction (the microbenchmark above reuses one a and one b)
template <int ACCS>
__global__ void __launch_bounds__(256) k_dmma_varied(double *out, int iters, double a0, double b0) {
    double c0[ACCS], c1[ACCS], a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = a0 + i * 1e-3 + threadIdx.x * 1e-6; b[i] = b0 + i * 1e-4; }
#pragma unroll
    for (int i = 0; i < ACCS; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < ACCS; ++i) dmma884(c0[i], c1[i], a[(i + u) & 7], b[(i * 3 + u) & 7]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACCS; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA stream fed from shared memory like the fused kernel's layer 1: per k-step 2 fragment loads (LDS.64), 3 DMMAs
template <int ACCS>
__global__ void __launch_bounds__(256) k_dmma_lds(double *out, int iters, double a0) {
    __shared__ double w[2][8 * ACCS * 32];
    for (int i = threadIdx.x; i < 2 * 8 * ACCS * 32; i += blockDim.x) (&w[0][0])[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double x0[ACCS], x1[ACCS], r0[ACCS], r1[ACCS];
#pragma unroll
    for (int i = 0; i < ACCS; ++i) { x0[i] = threadIdx.x + i; x1[i] = i; r0[i] = 1; r1[i] = 2; }
    const int lane = threadIdx.x & 31;
    double ay = a0, ar = a0 * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int k = 0; k < 16; ++k) {
#pragma unroll
            for (int i = 0; i < ACCS; ++i) {
                const double bw = w[0][((k & 7) * ACCS + i) * 32 + lane], bv = w[1][((k & 7) * ACCS + i) * 32 + lane];
                dmma884(r0[i], r1[i], ar, bw);
                dmma884(x0[i], x1[i], ay, bw);
                dmma884(r0[i], r1[i], ay, bv);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACCS; ++i) s += x0[i] + x1[i] + r0[i] + r1[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix: DM dmma884 per DF dfma in the same warp (are the pipes shared?)
template <int DM, int DF>
__global__ void __launch_bounds__(256) k_mix(double *out, int iters, double a, double b) {
    double c0[DM], c1[DM], acc[DF];
#pragma unroll
    for (int i = 0; i < DM; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < DF; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < DM; ++i) dmma884(c0[i], c1[i], a, b);
#pragma unroll
        for (int i = 0; i < DF; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < DM; ++i) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < DF; ++i) s += acc[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP32 FFMA peak for the optional FP32 mode
template <int CHAINS>
__global__ void __launch_bounds__(256) k_ffma(float *out, int iters, float a, float b) {
    float acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// legacy warp-level TF32 / BF16 tensor MMA (candidate for a 3xTF32 FP32 mode)
template <int CHAINS>
__global__ void __launch_bounds__(256) k_mma_tf32(float *out, int iters, float a, float b) {
    float c[CHAINS][4];
    unsigned af[4], bf[2];
    for (int i = 0; i < 4; ++i) af[i] = __float_as_uint(a + i);
    for (int i = 0; i < 2; ++i) bf[i] = __float_as_uint(b + i);
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(af[0]), "r"(af[1]), "r"(af[2]), "r"(af[3]), "r"(bf[0]), "r"(bf[1]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// tanh cost on the FP64 pipe
__global__ void __launch_bounds__(256) k_tanh(double *out, int iters, double a) {
    double x[4] = {a + threadIdx.x * 1e-3, a * 0.5, -a, a * 0.25};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = tanh(x[i]) + 0.3;
    }
    double s = x[0] + x[1] + x[2] + x[3];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char **argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    printf("device=%s sms=%d cc=%d.%d\n", p.name, sms, p.major, p.minor);
    double *out; CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024 * 4));
    const int iters = 4096;
    const int reps = 5;
    for (int bps = 1; bps <= 8; bps *= 2) {   // blocks per SM (256 threads each)
        int grid = sms * bps;
        double thr = (double)grid * 256;
        float ms;
        ms = time_ms([&] { k_dfma<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, reps);
        printf("bps=%d dfma16      : %8.3f ms  %7.2f TFLOP/s\n", bps, ms, thr * 16 * iters * 2 / ms / 1e9);
        ms = time_ms([&] { k_dmma884<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, reps);
        printf("bps=%d dmma884 x8  : %8.3f ms  %7.2f TFLOP/s\n", bps, ms, thr / 32 * 8 * iters * 512.0 / ms / 1e9);
        ms = time_ms([&] { k_dmma884<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, reps);
        printf("bps=%d dmma884 x16 : %8.3f ms  %7.2f TFLOP/s\n", bps, ms, thr / 32 * 16 * iters * 512.0 / ms / 1e9);
        ms = time_ms([&] { k_dmma1684<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, reps);
        printf("bps=%d dmma1684 x8 : %8.3f ms  %7.2f TFLOP/s\n", bps, ms, thr / 32 * 8 * iters * 1024.0 / ms / 1e9);
        ms = time_ms([&] { k_dmma1688<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, reps);
        printf("bps=%d dmma1688 x8 : %8.3f ms  %7.2f TFLOP/s\n", bps, ms, thr / 32 * 8 * iters * 2048.0 / ms / 1e9);
        ms = time_ms([&] { k_dmma16816<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, reps);
        printf("bps=%d dmma16816 x8: %8.3f ms  %7.2f TFLOP/s\n", bps, ms, thr / 32 * 8 * iters * 4096.0 / ms / 1e9);
        ms = time_ms([&] { k_mix<8, 8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, reps);
        printf("bps=%d mix 8dmma+8dfma : %8.3f ms  %7.2f TFLOP/s (dmma part %7.2f, dfma part %7.2f)\n", bps, ms,
               (thr / 32 * 8 * iters * 512.0 + thr * 8 * iters * 2) / ms / 1e9, thr / 32 * 8 * iters * 512.0 / ms / 1e9, thr * 8 * iters * 2 / ms / 1e9);
        ms = time_ms([&] { k_mix<8, 32><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); }, reps);
        printf("bps=%d mix 8dmma+32dfma: %8.3f ms  %7.2f TFLOP/s (dmma part %7.2f, dfma part %7.2f)\n", bps, ms,
               (thr / 32 * 8 * (iters / 4) * 512.0 + thr * 32 * (iters / 4) * 2) / ms / 1e9, thr / 32 * 8 * (iters / 4) * 512.0 / ms / 1e9, thr * 32 * (iters / 4) * 2 / ms / 1e9);
        ms = time_ms([&] { k_ffma<16><<<grid, 256>>>((float *)out, iters, 1.0000001f, 1e-9f); }, reps);
        printf("bps=%d ffma16      : %8.3f ms  %7.2f TFLOP/s\n", bps, ms, thr * 16 * iters * 2 / ms / 1e9);
        ms = time_ms([&] { k_tanh<<<grid, 256>>>(out, iters / 16, 0.7); }, reps);
        printf("bps=%d tanh(f64)   : %8.3f ms  %7.2f Gtanh/s\n", bps, ms, thr * 4 * (iters / 16) / ms / 1e6);
    }
    {
        const int grid = sms;
        const double thr = (double)grid * 256;
        float ms = time_ms([&] { k_dmma_varied<16><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); }, reps);
        printf("dmma884 16 accs, varied a/b regs, 8 warps/SM: %8.3f ms  %7.2f TFLOP/s\n", ms, thr / 32 * 16 * 4 * (iters / 4) * 512.0 / ms / 1e9);
        ms = time_ms([&] { k_dmma_varied<8><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); }, reps);
        printf("dmma884  8 accs, varied a/b regs, 8 warps/SM: %8.3f ms  %7.2f TFLOP/s\n", ms, thr / 32 * 8 * 4 * (iters / 4) * 512.0 / ms / 1e9);
        ms = time_ms([&] { k_dmma_lds<8><<<grid, 256>>>(out, iters / 64, 1.0000001); }, reps);
        printf("dmma884 LDS-fed (2 LDS.64 per 3 DMMA, 16 accs), 8 warps/SM: %8.3f ms  %7.2f TFLOP/s\n", ms, thr / 32 * 8 * 3 * 16 * (iters / 64) * 512.0 / ms / 1e9);
        ms = time_ms([&] { k_dmma_lds<4><<<grid, 256>>>(out, iters / 64, 1.0000001); }, reps);
        printf("dmma884 LDS-fed (2 LDS.64 per 3 DMMA,  8 accs), 8 warps/SM: %8.3f ms  %7.2f TFLOP/s\n", ms, thr / 32 * 4 * 3 * 16 * (iters / 64) * 512.0 / ms / 1e9);
    }
    for (int bps = 1; bps <= 4; bps *= 2) {
        const int grid = sms * bps;
        const double thr = (double)grid * 256;
        float ms = time_ms([&] { k_mma_tf32<8><<<grid, 256>>>((float *)out, iters, 1.0000001f, 1e-9f); }, reps);
        printf("bps=%d mma.sync m16n8k8 tf32 x8: %8.3f ms  %7.2f TFLOP/s\n", bps, ms, thr / 32 * 8 * iters * (2.0 * 16 * 8 * 8) / ms / 1e9);
    }
    // DMMA / DFMA dependent-issue latency: one warp per SM sub-partition (128 threads per SM), CH independent chains
    {
        const int grid = sms;
        float ms;
        ms = time_ms([&] { k_dmma884<1><<<grid, 128>>>(out, iters, 1.0000001, 1e-9); }, reps);
        printf("latency dmma884 1 chain : %8.3f ms  => %.1f cycles per dependent DMMA at 1.965 GHz\n", ms, ms * 1e-3 * 1.965e9 / iters);
        ms = time_ms([&] { k_dmma884<2><<<grid, 128>>>(out, iters, 1.0000001, 1e-9); }, reps);
        printf("latency dmma884 2 chains: %8.3f ms  => %.1f cycles per DMMA issue\n", ms, ms * 1e-3 * 1.965e9 / iters / 2);
        ms = time_ms([&] { k_dmma884<4><<<grid, 128>>>(out, iters, 1.0000001, 1e-9); }, reps);
        printf("latency dmma884 4 chains: %8.3f ms  => %.1f cycles per DMMA issue\n", ms, ms * 1e-3 * 1.965e9 / iters / 4);
        ms = time_ms([&] { k_dmma884<8><<<grid, 128>>>(out, iters, 1.0000001, 1e-9); }, reps);
        printf("latency dmma884 8 chains: %8.3f ms  => %.1f cycles per DMMA issue\n", ms, ms * 1e-3 * 1.965e9 / iters / 8);
        ms = time_ms([&] { k_dfma<1><<<grid, 128>>>(out, iters, 1.0000001, 1e-9); }, reps);
        printf("latency dfma 1 chain    : %8.3f ms  => %.1f cycles per dependent DFMA\n", ms, ms * 1e-3 * 1.965e9 / iters);
    }
    CK(cudaFree(out));
    return 0;
}
