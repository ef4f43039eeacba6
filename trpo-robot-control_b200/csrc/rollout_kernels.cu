// rollout_kernels.cu -- the steps either side of the update inside the reference's training loop (SURVEY.md section 8,
// rows f-3 / f-4): return + GAE advantage + standardisation (TRPO_Lightweight.c:565-653), the time-augmented observation
// matrix the value-function ("baseline") network reads (TRPO_Baseline.c:98-103) and the squared-error sum of its
// objective (TRPO_Baseline.c:221-226). All of it is HBM-bound streaming work over N-length vectors; reductions are
// fixed-order (thread-strided partial -> shuffle tree -> shared tree -> one final block) so results are reproducible.
#include "trpo_internal.cuh"

namespace {

constexpr int RB_THREADS = 256;
constexpr int RB_BLOCKS = 592;    // 4 x 148 SMs; also the size of the block-partial scratch the callers hand in (<= 1024)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double block_sum(double v, double *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (w == 0) {
        t = warp_sum(t);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}

// out[n][0..O) = obs[n][0..O), out[n][O] = (n mod EpLen) / EpLen
__global__ void __launch_bounds__(RB_THREADS) k_vf_augment(const double *__restrict__ obs, size_t n, int O, size_t ep_len,
                                                           double *__restrict__ out) {
    const size_t total = n * (size_t)(O + 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t s = i / (O + 1);
        const int j = (int)(i - s * (O + 1));
        out[i] = (j < O) ? obs[s * O + j] : (double)(s % ep_len) / (double)ep_len;
    }
}

// One warp per episode, walking it backwards in 32-step chunks. Both quantities are first-order linear recurrences
//   Return[t]    = Reward[t] + gamma       * Return[t+1]
//   Advantage[t] = delta[t]  + gamma * lam * Advantage[t+1],   delta[t] = Reward[t] + gamma * V[t+1] - V[t]  (V[EpLen] = 0)
// (TRPO_Lightweight.c:565-580,627-641 writes them as O(EpLen^2) pow() sums); inside a chunk they are evaluated with a
// 5-step shuffle scan, the carry from the chunk behind enters with weight a^(32 - lane). Coalesced loads and stores.
__global__ void __launch_bounds__(RB_THREADS) k_gae(const double *__restrict__ reward, const double *__restrict__ baseline,
                                                    size_t num_ep, int ep_len, double gamma, double lam,
                                                    double *__restrict__ ret, double *__restrict__ adv) {
    const size_t ep = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ep >= num_ep) return;
    const double *R = reward + ep * (size_t)ep_len, *V = baseline + ep * (size_t)ep_len;
    double *RET = ret + ep * (size_t)ep_len, *ADV = adv + ep * (size_t)ep_len;
    const double gl = gamma * lam;
    const double wr = pow(gamma, (double)(32 - lane)), wa = pow(gl, (double)(32 - lane));
    double carry_r = 0.0, carry_a = 0.0;
    for (int t0 = ((ep_len - 1) / 32) * 32; t0 >= 0; t0 -= 32) {
        const int t = t0 + lane;
        const bool in = t < ep_len;
        const double r = in ? R[t] : 0.0;
        const double v = in ? V[t] : 0.0;
        const double vn = (t + 1 < ep_len) ? V[t + 1] : 0.0;
        double xr = r, xa = in ? (r + gamma * vn - v) : 0.0;
        double pr = gamma, pa = gl;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double yr = __shfl_down_sync(0xffffffffu, xr, off), ya = __shfl_down_sync(0xffffffffu, xa, off);
            if (lane + off < 32) { xr += pr * yr; xa += pa * ya; }
            pr *= pr; pa *= pa;
        }
        xr += wr * carry_r;
        xa += wa * carry_a;
        if (in) { RET[t] = xr; ADV[t] = xa; }
        carry_r = __shfl_sync(0xffffffffu, xr, 0);
        carry_a = __shfl_sync(0xffffffffu, xa, 0);
    }
}

// block partials of sum_i (a[i] - c_i)^2 with c_i = b[i] (b != NULL) or the scalar *shift_sum / shift_div
__global__ void __launch_bounds__(RB_THREADS) k_sqdiff_blocks(const double *__restrict__ a, const double *__restrict__ b,
                                                              const double *__restrict__ shift_sum, double shift_div, size_t n,
                                                              double *__restrict__ block_partials) {
    __shared__ double red[32];
    const double c = b ? 0.0 : shift_sum[0] / shift_div;
    double acc = 0.0;
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (size_t)gridDim.x * blockDim.x) {
        const double d = a[s] - (b ? b[s] : c);
        acc += d * d;
    }
    const double sblk = block_sum(acc, red);
    if (threadIdx.x == 0) block_partials[blockIdx.x] = sblk;
}

__global__ void __launch_bounds__(1024) k_partials_final(const double *__restrict__ a, int n, double *__restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += a[i];
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) out[0] = s;
}

// a = (a - mean) / std with mean = sum/N, std = sqrt(sqdev/N) (population, TRPO_Lightweight.c:645-653)
__global__ void __launch_bounds__(RB_THREADS) k_standardise(double *__restrict__ a, size_t n, const double *__restrict__ sum,
                                                            const double *__restrict__ sqdev, double n_total) {
    const double mean = sum[0] / n_total, sd = sqrt(sqdev[0] / n_total);
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (size_t)gridDim.x * blockDim.x)
        a[s] = (a[s] - mean) / sd;
}

__global__ void __launch_bounds__(RB_THREADS) k_fill(double *__restrict__ a, double v, size_t n) {
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (size_t)gridDim.x * blockDim.x) a[s] = v;
}

inline int grid_for(size_t n) {
    size_t b = (n + RB_THREADS - 1) / RB_THREADS;
    return (int)(b < 1 ? 1 : (b > RB_BLOCKS ? RB_BLOCKS : b));
}

}  // namespace

void launch_vf_augment(const double *d_obs, size_t n, int O, size_t ep_len, double *d_out, cudaStream_t st, long long *launches) {
    k_vf_augment<<<grid_for(n * (size_t)(O + 1)), RB_THREADS, 0, st>>>(d_obs, n, O, ep_len, d_out);
    ++*launches;
}

void launch_gae(const double *d_reward, const double *d_baseline, size_t num_ep, int ep_len, double gamma, double lam,
                double *d_ret, double *d_adv, cudaStream_t st, long long *launches) {
    const size_t warps_per_block = RB_THREADS / 32;
    const unsigned blocks = (unsigned)((num_ep + warps_per_block - 1) / warps_per_block);
    k_gae<<<blocks, RB_THREADS, 0, st>>>(d_reward, d_baseline, num_ep, ep_len, gamma, lam, d_ret, d_adv);
    ++*launches;
}

void launch_sqdiff(const double *d_a, const double *d_b, const double *d_shift_sum, double shift_div, size_t n,
                   double *d_block_partials, double *d_out, cudaStream_t st, long long *launches) {
    k_sqdiff_blocks<<<RB_BLOCKS, RB_THREADS, 0, st>>>(d_a, d_b, d_shift_sum, shift_div, n, d_block_partials);
    k_partials_final<<<1, 1024, 0, st>>>(d_block_partials, RB_BLOCKS, d_out);
    *launches += 2;
}

void launch_standardise(double *d_a, size_t n, const double *d_sum, const double *d_sqdev, double n_total, cudaStream_t st,
                        long long *launches) {
    k_standardise<<<grid_for(n), RB_THREADS, 0, st>>>(d_a, n, d_sum, d_sqdev, n_total);
    ++*launches;
}

void launch_fill(double *d_a, double v, size_t n, cudaStream_t st, long long *launches) {
    k_fill<<<grid_for(n), RB_THREADS, 0, st>>>(d_a, v, n);
    ++*launches;
}
