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
// 16-byte variant (both addresses 16-byte aligned); src_bytes in {0, 8, 16}: the remainder is zero-filled
__device__ __forceinline__ void cp_async16(double *dst_smem, const double *src, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

// Branch-free FP64 tanh for N values in lock step: tanh(x) = sign(x) * (1 - 2 / (exp(2|x|) + 1)).
// The library tanh() has data-dependent branches, so the unrolled per-element calls cannot be interleaved and every
// call runs as one ~30-deep dependent DFMA chain (11.6 cycles each): with 2 warps per scheduler that was a third of
// the fused kernel's time. Here every stage is applied to all N values before the next one, so the FP64 pipe sees N
// independent chains. exp(2a) = 2^n * 2^(j/64) * e^r with k = rint(2a * 64/ln2), n = k >> 6, j = k & 63,
// r = 2a - k ln2/64 (|r| <= 0.0055): a 64-entry table (shared memory) and a degree-5 Taylor polynomial (truncation
// 3.5e-17); 2^n goes in through the exponent bits; 1/(E+1) from MUFU.RCP64H + two Newton steps. 19 FP64-pipe
// instructions per value. Max absolute error 2.6e-16 over [-25, 25] (tests/test_host_logic.py holds the same
// algorithm in numpy).
__constant__ double c_exp2_tab[64] = {
    1, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.1023825833078409, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.2021567314527031, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.2553807570246911, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.3396675240533029,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.5590044002378369, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.6457554781539649, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.7186192981224779, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.9784560263879509};

// every kernel that calls tanh_vec copies the table into its shared memory once
__device__ __forceinline__ void load_exp2_table(double *tab_smem) {
    for (int i = threadIdx.x; i < 64; i += blockDim.x) tab_smem[i] = c_exp2_tab[i];
}

template <int N>
__device__ __forceinline__ void tanh_vec(double (&x)[N], double (&d)[N], const double *__restrict__ tab) {
    constexpr double MAGIC = 6755399441055744.0;            // 1.5 * 2^52: rounds to nearest integer in the low word
    constexpr double INV = 92.33248261689366;               // 64 / ln 2
    constexpr double LN2_64_HI = 0.010830424667801708, LN2_64_LO = 2.8447437476627285e-11;   // HI has 24 trailing zero bits
    double z[N], r[N], q[N];
    int k[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { const double a = fmin(fabs(x[i]), 20.0); z[i] = a + a; }   // tanh(20) == 1 in double
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double tt = fma(z[i], INV, MAGIC);
        k[i] = __double2loint(tt);
        const double kf = tt - MAGIC;
        r[i] = fma(kf, -LN2_64_HI, z[i]);
        r[i] = fma(kf, -LN2_64_LO, r[i]);
    }
    double tj[N];
#pragma unroll
    for (int i = 0; i < N; ++i) tj[i] = tab[k[i] & 63];
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = fma(1.0 / 120.0, r[i], 1.0 / 24.0);
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = fma(q[i], r[i], 1.0 / 6.0);
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = fma(q[i], r[i], 0.5);
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = fma(q[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = fma(q[i], r[i], 1.0);
    double s[N], y[N], e[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double m = q[i] * tj[i];
        const double E = __hiloint2double(__double2hiint(m) + ((k[i] >> 6) << 20), __double2loint(m));
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
        const double v = copysign(fma(-2.0, y[i], 1.0), x[i]);
        x[i] = v;
        d[i] = fma(-v, v, 1.0);
    }
}
