// dmma_common.cuh -- device helpers shared by the fused FVP kernel and the DMMA GEMM-chain kernels.
#pragma once

// FP64 tensor-pipe MMA: D(8x8) += A(8x4) * B(4x8). Fragment layout (lane = 4*g + t):
//   A: row g, col t      B: row t, col g      C/D: row g, cols 2t and 2t+1
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// 8-byte async global->shared copy (LDGSTS); src_bytes == 0 zero-fills the destination
__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

// Branch-free FP64 tanh for N values in lock step: tanh(x) = sign(x) * (1 - 2 / (exp(2|x|) + 1)).
// The library tanh() has data-dependent branches, so the unrolled per-element calls cannot be interleaved and every
// call runs as one ~30-deep dependent DFMA chain (11.6 cycles each): with 2 warps per scheduler that was a third of
// the kernel's time. Here every stage is applied to all N values before the next one, so the FP64 pipe sees N
// independent chains. exp(2a) = 2^n * e^{2h}, h = a - n*ln2/2 (|h| <= 0.174), degree-12 Taylor polynomial in h
// (truncation 1.7e-16), 2^n applied through the exponent bits; 1/(E+1) from MUFU.RCP64H + two Newton steps.
// Max absolute error 2.6e-16 over [-25, 25] (tests/test_host_logic.py holds the same algorithm in numpy).
template <int N>
__device__ __forceinline__ void tanh_vec(double (&x)[N], double (&d)[N]) {
    constexpr double MAGIC = 6755399441055744.0;            // 1.5 * 2^52: rounds to nearest integer in the low word
    constexpr double L2E2 = 2.8853900817779268;             // 2 / ln 2
    constexpr double LN2H_HI = 0.3465735901845619, LN2H_LO = 9.541074646352939e-11;   // ln2/2 = HI + LO, HI has 21 trailing zero bits
    constexpr double C[13] = {1.0, 2.0, 2.0, 1.3333333333333333, 0.6666666666666666, 0.26666666666666666,
                              0.08888888888888889, 0.025396825396825397, 0.006349206349206349, 0.0014109347442680777,
                              0.0002821869488536155, 5.130671797338464e-05, 8.551119662230774e-06};   // 2^k / k!
    double a[N], h[N], q[N];
    int ni[N];
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = fmin(fabs(x[i]), 20.0);       // tanh(20) == 1 in double
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double tt = fma(a[i], L2E2, MAGIC);
        ni[i] = __double2loint(tt);
        const double nf = tt - MAGIC;
        h[i] = fma(nf, -LN2H_HI, a[i]);
        h[i] = fma(nf, -LN2H_LO, h[i]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = fma(C[12], h[i], C[11]);
#pragma unroll
    for (int k = 10; k >= 0; --k)
#pragma unroll
        for (int i = 0; i < N; ++i) q[i] = fma(q[i], h[i], C[k]);
    double s[N], y[N], e[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double E = __hiloint2double(__double2hiint(q[i]) + (ni[i] << 20), __double2loint(q[i]));
        s[i] = E + 1.0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(s[i]));
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) e[i] = fma(-s[i], y[i], 1.0);
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = fma(y[i], e[i], y[i]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double r = copysign(fma(-2.0, y[i], 1.0), x[i]);
        x[i] = r;
        d[i] = fma(-r, r, 1.0);
    }
}

