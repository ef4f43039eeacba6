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

// ---------------------------------------------------------------------------------------------------------------------
// Rollout producer: the reference's lightweight arm simulator (TRPO_Lightweight.c:349-540; on the FPGA build the role of
// TRPO_RunLightweight, TRPO_Lightweight_FPGA.c:548-556). One WARP per episode: lane j owns neuron j of every layer (all
// widths <= 32), the activations of the previous layer are broadcast by shuffles in the reference's summation order
// (bias first, then k ascending), the scalar simulator state (joint angles, link positions, reward) is computed
// redundantly by every lane -- SIMT issues it once per warp either way. Random numbers: either the caller's raw rand()
// draws in the reference's order (3 per episode for the object position, then 2 per action component per step), which
// makes the batch comparable to the reference's, or a counter-based generator (splitmix64 of seed + draw index).
namespace {

constexpr int ROLL_MAX_LAYERS = 8;
struct RolloutNet {
    int K;
    int L[ROLL_MAX_LAYERS + 1];
    int w_off[ROLL_MAX_LAYERS];
    int logstd_off, P;
    char ac[ROLL_MAX_LAYERS + 1];
};

__device__ __forceinline__ double roll_act(char a, double x) {
    switch (a) {
        case 't': return tanh(x);
        case 'o': return 0.1 * x;
        case 's': return 1.0 / (1.0 + exp(-x));
        default:  return x;
    }
}

__device__ __forceinline__ int roll_draw(const int *__restrict__ draws, unsigned long long seed, size_t idx) {
    if (draws) return draws[idx];
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (int)(z >> 33);                                  // 31 bits, the range of glibc's rand()
}

constexpr double ROLL_RAND_MAX = 2147483647.0;

__global__ void __launch_bounds__(128) k_arm_rollout(RolloutNet net, const double *__restrict__ theta,
                                                     const int *__restrict__ draws, unsigned long long seed,
                                                     size_t num_ep, int ep_len,
                                                     double *__restrict__ observ, double *__restrict__ mean,
                                                     double *__restrict__ action, double *__restrict__ reward) {
    extern __shared__ double sm[];
    double *th = sm;                                        // the whole parameter vector
    double *obs_sm = sm + net.P + (threadIdx.x >> 5) * 16;  // 15 observation components of this warp's current step
    for (int i = threadIdx.x; i < net.P; i += blockDim.x) th[i] = theta[i];
    __syncthreads();
    const size_t ep = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ep >= num_ep) return;
    const int K = net.K, O = net.L[0], A = net.L[K];
    const double pi = 3.1415926535897931, TimeStepLen = 0.02, coeff = 1;
    const size_t per_ep = 3 + (size_t)2 * A * ep_len;
    const size_t d0 = ep * per_ep;
    double t1 = 0, t2 = -pi / 2.0, t3 = pi / 2.0;
    double dof2x = 0, dof2y = 0, dof2z = 0.07518, wrx = 0.07375, wry = 0, wrz = 0.07518, grx = 0.11315, gry = 0, grz = 0.06268;
    const double objx = ((double)roll_draw(draws, seed, d0 + 0) / ROLL_RAND_MAX) * 0.076 + 0.084;
    const double objy = ((double)roll_draw(draws, seed, d0 + 1) / ROLL_RAND_MAX) * 0.100 - 0.05;
    const double objz = ((double)roll_draw(draws, seed, d0 + 2) / ROLL_RAND_MAX) * 0.100;
    const double sd = lane < A ? exp(th[net.logstd_off + lane]) : 1.0;
    for (int step = 0; step < ep_len; ++step) {
        const size_t row = ep * (size_t)ep_len + step;
        if (lane == 0) {
            obs_sm[0] = 0; obs_sm[1] = 0; obs_sm[2] = 0.01768;
            obs_sm[3] = dof2x; obs_sm[4] = dof2y; obs_sm[5] = dof2z;
            obs_sm[6] = wrx; obs_sm[7] = wry; obs_sm[8] = wrz;
            obs_sm[9] = grx; obs_sm[10] = gry; obs_sm[11] = grz;
            obs_sm[12] = objx; obs_sm[13] = objy; obs_sm[14] = objz;
        }
        __syncwarp();
        if (lane < O) observ[row * O + lane] = obs_sm[lane];
        // policy forward: lane j = neuron j
        double y = 0.0;
        {
            const int N = net.L[1];
            const double *W = th + net.w_off[0];
            double x = lane < N ? W[O * N + lane] : 0.0;
            for (int k = 0; k < O; ++k) x += obs_sm[k] * (lane < N ? W[k * N + lane] : 0.0);
            y = roll_act(net.ac[1], x);
        }
        for (int i = 1; i < K; ++i) {
            const int M = net.L[i], N = net.L[i + 1];
            const double *W = th + net.w_off[i];
            double x = lane < N ? W[M * N + lane] : 0.0;
            for (int k = 0; k < M; ++k) {
                const double yk = __shfl_sync(0xffffffffu, y, k);
                x += yk * (lane < N ? W[k * N + lane] : 0.0);
            }
            y = roll_act(net.ac[i + 1], x);
        }
        __syncwarp();                                       // obs_sm is rewritten at the top of the next step
        // sample the action (Box-Muller on two draws per component, TRPO_Lightweight.c:463-474)
        double ac = 0.0;
        if (lane < A) {
            const size_t di = d0 + 3 + (size_t)2 * A * step + 2 * lane;
            const double u1 = ((double)roll_draw(draws, seed, di) + 1.0) / (ROLL_RAND_MAX + 1.0);
            const double u2 = ((double)roll_draw(draws, seed, di + 1) + 1.0) / (ROLL_RAND_MAX + 1.0);
            const double z0 = sqrt(-2.0 * log(u1)) * cos(2 * pi * u2);
            ac = z0 * sd + y;
            mean[row * A + lane] = y;
            action[row * A + lane] = ac;
        }
        const double a0 = __shfl_sync(0xffffffffu, ac, 0), a1 = __shfl_sync(0xffffffffu, ac, 1), a2 = __shfl_sync(0xffffffffu, ac, 2);
        t1 += a0 * coeff * TimeStepLen;
        t2 += a1 * coeff * TimeStepLen;
        t3 += a2 * coeff * TimeStepLen;
        const double s1 = sin(t1), c1 = cos(t1), s2 = sin(t2), c2 = cos(t2), s3 = sin(t3), c3 = cos(t3);
        const double c2c3 = c2 * c3, s2s3 = s2 * s3, c2s3 = c2 * s3, s2c3 = s2 * c3;
        dof2x = 0.0575 * c1 * c2;
        dof2y = 0.0575 * s1 * c2;
        dof2z = 0.01768 - 0.0575 * s2;
        wrx = dof2x + 0.07375 * c1 * (c2c3 - s2s3);
        wry = dof2y + 0.07375 * s1 * (c2c3 - s2s3);
        wrz = dof2z - 0.07375 * (c2s3 + s2c3);
        grx = dof2x + 0.11315 * c1 * (c2c3 - s2s3) - 0.0125 * c1 * (c2s3 + s2c3);
        gry = dof2y + 0.11315 * s1 * (c2c3 - s2s3) - 0.0125 * s1 * (c2s3 + s2c3);
        grz = dof2z - 0.11315 * (c2s3 + s2c3) + 0.0125 * (s2s3 - c2c3);
        double re = 0;
        re -= 100 * (objx - grx) * (objx - grx); re -= a0 * a0;
        re -= 100 * (objy - gry) * (objy - gry); re -= a1 * a1;
        re -= 100 * (objz - grz) * (objz - grz); re -= a2 * a2;
        if (lane == 0) reward[row] = re;
    }
}

}  // namespace

int launch_arm_rollout(const NetDesc &net, const double *d_theta, const int *d_draws, unsigned long long seed, size_t num_ep,
                       int ep_len, double *d_obs, double *d_mean, double *d_action, double *d_reward, cudaStream_t st,
                       long long *launches) {
    if (net.K > ROLL_MAX_LAYERS || net.L[0] != 15 || net.L[net.K] != 3) return 1;     // the simulator's observation / action layout
    RolloutNet rn;
    rn.K = net.K; rn.logstd_off = net.logstd_off; rn.P = net.P;
    for (int i = 0; i <= net.K; ++i) { if (i && net.L[i] > 32) return 1; rn.L[i] = net.L[i]; rn.ac[i] = net.ac[i]; }
    for (int i = 0; i < net.K; ++i) rn.w_off[i] = net.w_off[i];
    const int threads = 128, warps = threads / 32;
    const size_t smem = sizeof(double) * ((size_t)net.P + 16 * warps);
    if (smem > 200 * 1024) return 1;
    // per device, so set on every launch (a process may drive several GPUs); only deep 32-wide policies exceed 48 KB
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_arm_rollout, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    const unsigned blocks = (unsigned)((num_ep + warps - 1) / warps);
    k_arm_rollout<<<blocks, threads, smem, st>>>(rn, d_theta, d_draws, seed, num_ep, ep_len, d_obs, d_mean, d_action, d_reward);
    ++*launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
