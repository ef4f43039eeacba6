// peak_probe.cu -- roofline denominators measured on the device the context runs on, in the same process as the benchmark.
// MEASURED_PEAKS.json (driver-written) carries HBM bandwidth and bf16 tensor throughput but no FP64 figure, and the FVP kernels
// are bound by the FP64 pipe: bench.py calls these probes right before its timed region (tools/fp64_peak.cu is the long form).
#include <cuda_runtime.h>

#include "../../include/trpo_b200.h"

namespace {

template <int CHAINS>
__global__ void __launch_bounds__(256) k_probe_dmma(double *out, int iters, double a, double b) {
    double c0[CHAINS], c1[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(256) k_probe_tf32(float *out, int iters, float a, float b) {
    float c[CHAINS][4];
    unsigned af[4], bf[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) af[i] = __float_as_uint(a + i);
    bf[0] = __float_as_uint(b); bf[1] = __float_as_uint(b * 2);
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

template <typename F>
double best_ms(F launch) {
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1.0;
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return cudaGetLastError() == cudaSuccess ? (double)best : -1.0;
}

}  // namespace

// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) with 16 independent accumulators per warp, 2 CTAs of 8 warps per SM: TFLOP/s, or -1.
extern "C" double trpo_probe_fp64_peak_tflops(int device) {
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp p;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) return -1.0;
    double *out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * 1024 * 1024) != cudaSuccess) return -1.0;
    const int grid = p.multiProcessorCount * 2, iters = 4096;
    const double ms = best_ms([&] { k_probe_dmma<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
    cudaFree(out);
    if (ms <= 0) return -1.0;
    return (double)grid * 8 * 16 * iters * 512.0 / ms / 1e9;
}

// legacy mma.sync.m16n8k8 TF32 (SASS HMMA.1688.F32.TF32), the instruction of the first FP32-mode kernels: TFLOP/s of single
// TF32 products (a 3xTF32 product costs three of them), or -1.
extern "C" double trpo_probe_tf32_mma_sync_tflops(int device) {
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp p;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) return -1.0;
    float *out = nullptr;
    if (cudaMalloc(&out, sizeof(float) * 1024 * 1024) != cudaSuccess) return -1.0;
    const int grid = p.multiProcessorCount * 2, iters = 4096;
    const double ms = best_ms([&] { k_probe_tf32<8><<<grid, 256>>>(out, iters, 1.0000001f, 1e-9f); });
    cudaFree(out);
    if (ms <= 0) return -1.0;
    return (double)grid * 8 * 8 * iters * (2.0 * 16 * 8 * 8) / ms / 1e9;
}
