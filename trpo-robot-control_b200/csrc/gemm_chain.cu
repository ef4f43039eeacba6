// gemm_chain.cu -- generic (any NumLayers / any widths) path of the Fisher-vector product and policy gradient.
//
// The per-sample loops of the reference (TRPO_FVP.c:771-924, TRPO_Update.c:254-371) are restructured as a chain of
// batched FP64 GEMMs over a chunk of samples whose activations stay L2-resident:
//   forward  (per layer i)   [Y | RY]_{i+1} = act( [Y_i,1] * [W_i;B_i] ),  RX = RY_i*W_i + [Y_i,1]*[VW_i;VB_i]   (:783-836)
//   backward (per layer i)   G_{i-1} = (G_i * W_{i-1}^T) .* f'(Y_{i-1})                                            (:857-900)
//   outer    (per layer i)   [RGW_{i-1};RGB_{i-1}] += [Y_{i-1},1]^T * G_i   (split-K over sample slices)            (:890-921)
// The flat parameter layout already stores [W_i;B_i] as one (L_i+1) x L_{i+1} row-major matrix, so the bias is the
// "ones" row of the augmented operand and no packing is needed.
// Determinism: every output element has exactly one owner thread per (slice); slices are summed in fixed order.
#include "trpo_internal.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4, NT = 256;

__device__ __forceinline__ double act_apply(char a, double x) {
    switch (a) {
        case 't': return tanh(x);
        case 'o': return 0.1 * x;
        case 's': return 1.0 / (1.0 + exp(-x));
        default:  return x;
    }
}
__device__ __forceinline__ double act_deriv(char a, double y) {   // f'(x) expressed through y = f(x)
    switch (a) {
        case 't': return 1.0 - y * y;
        case 'o': return 0.1;
        case 's': return y * (1.0 - y);
        default:  return 1.0;
    }
}

// A-tile loader for row-major activations X[rows x ld]: As[k][m] = X[(m0+m)*ld + k0+k]; k == ld -> ones_val.
__device__ __forceinline__ void load_act_tile(double (*As)[BM], const double *X, int rows, int ld, int m0, int k0,
                                              double ones_val, int tid) {
    const int m = tid >> 2, ks = (tid & 3) * 4;
    const int gm = m0 + m;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int gk = k0 + ks + q;
        double v = 0.0;
        if (gm < rows) {
            if (gk < ld) v = X ? X[(size_t)gm * ld + gk] : 0.0;
            else if (gk == ld) v = ones_val;
        }
        As[ks + q][m] = v;
    }
}

// B-tile loader for a row-major matrix M[kdim x N]: Bs[k][n] = M[(k0+k)*N + n0+n]
__device__ __forceinline__ void load_rowmajor_tile(double (*Bs)[BN], const double *M, int kdim, int N, int k0, int n0, int tid) {
    const int k = tid >> 4, ns = (tid & 15) * 4;
    const int gk = k0 + k;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int gn = n0 + ns + q;
        Bs[k][ns + q] = (gk < kdim && gn < N) ? M[(size_t)gk * N + gn] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// forward: DUAL = also propagate R{} (FVP); otherwise ordinary forward only (policy gradient / line search).
template <bool DUAL>
__global__ void __launch_bounds__(NT) k_chain_fwd(const double *__restrict__ Yin, const double *__restrict__ RYin,
                                                  const double *__restrict__ W, const double *__restrict__ VW,
                                                  int rows, int Kd, int N, char act,
                                                  double *__restrict__ Yout, double *__restrict__ RYout,
                                                  double *__restrict__ Gout, const double *__restrict__ inv_var,
                                                  const int *__restrict__ done) {
    if (done && *done) return;
    __shared__ __align__(16) double As[BK][BM], Bs[BK][BN];
    __shared__ __align__(16) double RAs[DUAL ? BK : 1][BM], VBs[DUAL ? BK : 1][BN];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    double ax[TM][TN] = {}, arx[TM][TN] = {};
    for (int k0 = 0; k0 < Kd + 1; k0 += BK) {
        load_act_tile(As, Yin, rows, Kd, m0, k0, 1.0, tid);
        load_rowmajor_tile(Bs, W, Kd + 1, N, k0, n0, tid);
        if (DUAL) {
            load_act_tile(RAs, RYin, rows, Kd, m0, k0, 0.0, tid);
            load_rowmajor_tile(VBs, VW, Kd + 1, N, k0, n0, tid);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            double a[TM], b[TN], ra[TM], vb[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
            if (DUAL) {
#pragma unroll
                for (int i = 0; i < TM; ++i) ra[i] = RAs[kk][ty * TM + i];
#pragma unroll
                for (int j = 0; j < TN; ++j) vb[j] = VBs[kk][tx * TN + j];
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    ax[i][j] = fma(a[i], b[j], ax[i][j]);
                    if (DUAL) {
                        arx[i][j] = fma(ra[i], b[j], arx[i][j]);
                        arx[i][j] = fma(a[i], vb[j], arx[i][j]);
                    }
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gm = m0 + ty * TM + i;
        if (gm >= rows) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gn = n0 + tx * TN + j;
            if (gn >= N) continue;
            const double y = act_apply(act, ax[i][j]);
            const double d = act_deriv(act, y);
            if (Yout) Yout[(size_t)gm * N + gn] = y;
            if (DUAL) {
                const double ry = arx[i][j] * d;
                if (RYout) RYout[(size_t)gm * N + gn] = ry;
                // last layer: R-gradient seed RG_K = Ry_K / sigma^2 (TRPO_FVP.c:852-854), already times f'(y_K) (:869-882)
                if (Gout) Gout[(size_t)gm * N + gn] = ry * inv_var[gn] * d;
            }
        }
    }
}

// backward: Gout[s][n] = f'(Yprev[s][n]) * sum_k Gin[s][k] * W[n][k]       (W is [N x Kd] row-major)
__global__ void __launch_bounds__(NT) k_chain_bwd(const double *__restrict__ Gin, const double *__restrict__ W,
                                                  const double *__restrict__ Yprev, int rows, int Kd, int N, char act_prev,
                                                  double *__restrict__ Gout, const int *__restrict__ done) {
    if (done && *done) return;
    __shared__ __align__(16) double As[BK][BM], Bs[BK][BN];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    double acc[TM][TN] = {};
    for (int k0 = 0; k0 < Kd; k0 += BK) {
        load_act_tile(As, Gin, rows, Kd, m0, k0, 0.0, tid);
        {   // Bs[k][n] = W[(n0+n)*Kd + k0+k]: walk k contiguously, store transposed
            const int n = tid >> 2, ks = (tid & 3) * 4, gn = n0 + n;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gk = k0 + ks + q;
                Bs[ks + q][n] = (gn < N && gk < Kd) ? W[(size_t)gn * Kd + gk] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            double a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gm = m0 + ty * TM + i;
        if (gm >= rows) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gn = n0 + tx * TN + j;
            if (gn >= N) continue;
            const double d = act_deriv(act_prev, Yprev[(size_t)gm * N + gn]);
            Gout[(size_t)gm * N + gn] = acc[i][j] * d;
        }
    }
}

// outer: out[slice][m*N + n] (+)= sum_{s in slice} [Yprev,1][s][m] * G[s][n],  m in [0, M0]  (row M0 = bias gradient)
__global__ void __launch_bounds__(NT) k_chain_outer(const double *__restrict__ Yprev, const double *__restrict__ G,
                                                    int rows, int M0, int N, int per_slice, int tiles_n,
                                                    double *__restrict__ partial, int P, int out_off, int accumulate,
                                                    const int *__restrict__ done) {
    if (done && *done) return;
    __shared__ __align__(16) double As[BK][BM], Bs[BK][BN];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = (blockIdx.x / tiles_n) * BM, n0 = (blockIdx.x % tiles_n) * BN;
    const int slice = blockIdx.y;
    const int s0 = slice * per_slice;
    const int s1 = min(rows, s0 + per_slice);
    double acc[TM][TN] = {};
    for (int k0 = s0; k0 < s1; k0 += BK) {
        {   // As[k][m] = Yprev[(k0+k)*M0 + m0+m], ones at m == M0
            const int k = tid >> 4, ms = (tid & 15) * 4, gs = k0 + k;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gmm = m0 + ms + q;
                double v = 0.0;
                if (gs < s1) {
                    if (gmm < M0) v = Yprev[(size_t)gs * M0 + gmm];
                    else if (gmm == M0) v = 1.0;
                }
                As[k][ms + q] = v;
            }
        }
        {
            const int k = tid >> 4, ns = (tid & 15) * 4, gs = k0 + k;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int gn = n0 + ns + q;
                Bs[k][ns + q] = (gs < s1 && gn < N) ? G[(size_t)gs * N + gn] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            double a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    double *out = partial + (size_t)slice * P + out_off;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gm = m0 + ty * TM + i;
        if (gm > M0) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gn = n0 + tx * TN + j;
            if (gn >= N) continue;
            const size_t o = (size_t)gm * N + gn;
            out[o] = accumulate ? out[o] + acc[i][j] : acc[i][j];
        }
    }
}

// policy-gradient seed (TRPO_Update.c:297-301) followed by f'(y_K) (:310-324): G[s][j] and GL[s][j]
__global__ void k_pg_seed(const double *__restrict__ mean, const double *__restrict__ action, const double *__restrict__ adv,
                          const double *__restrict__ logstd, const double *__restrict__ YK, char actK,
                          int rows, int A, double *__restrict__ G, double *__restrict__ GL) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * A) return;
    const int s = idx / A, j = idx % A;
    const double sd = exp(logstd[j]);
    const double t = (action[idx] - mean[idx]) / sd;
    double g = adv[s] * t / sd;
    g *= act_deriv(actK, YK[idx]);
    G[idx] = g;
    GL[idx] = adv[s] * (t * t - 1.0);
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace

size_t chain_scratch_bytes(const NetDesc &net, int chunk, int nslices) {
    size_t maxL = 0, sumL = 0;
    for (int i = 1; i <= net.K; ++i) { sumL += net.L[i]; if ((size_t)net.L[i] > maxL) maxL = net.L[i]; }
    if ((size_t)net.L[0] > maxL) maxL = net.L[0];
    return sizeof(double) * ((size_t)chunk * (sumL + 4 * maxL) + (size_t)nslices * net.P);
}

int chain_accumulate(const NetDesc &net, const ChainScratch &sc, ChainMode mode,
                     const double *d_theta, const double *d_v, const double *d_inv_var,
                     const double *d_obs, const double *d_mean, const double *d_action, const double *d_adv,
                     size_t nsamples, double *d_zsum, const int *d_done, const P2PComm *p2p, cudaStream_t st,
                     long long *launches) {
    const int K = net.K, A = net.L[K];
    const bool fvp = (mode == CHAIN_FVP);
    int chunk_idx = 0;
    for (size_t c0 = 0; c0 < nsamples; c0 += sc.chunk, ++chunk_idx) {
        const int rows = (int)((nsamples - c0 < (size_t)sc.chunk) ? nsamples - c0 : sc.chunk);
        const int accumulate = chunk_idx > 0;
        const int per_slice = cdiv(cdiv(rows, sc.nslices), BK) * BK;
        // ---- forward ----
        for (int i = 0; i < K; ++i) {
            const double *Yin = (i == 0) ? d_obs + c0 * net.L[0] : sc.Y[i];
            const bool last = (i == K - 1);
            dim3 grid(cdiv(rows, BM), cdiv(net.L[i + 1], BN));
            if (fvp) {
                const double *RYin = (i == 0) ? nullptr : sc.RY[i & 1];
                const bool needY = !last || net.ac[K] == 't' || net.ac[K] == 's';
                k_chain_fwd<true><<<grid, NT, 0, st>>>(Yin, RYin, d_theta + net.w_off[i], d_v + net.w_off[i], rows,
                                                      net.L[i], net.L[i + 1], net.ac[i + 1],
                                                      needY ? sc.Y[i + 1] : nullptr, last ? nullptr : sc.RY[(i + 1) & 1],
                                                      last ? sc.G[K & 1] : nullptr, d_inv_var, d_done);
            } else {
                k_chain_fwd<false><<<grid, NT, 0, st>>>(Yin, nullptr, d_theta + net.w_off[i], nullptr, rows,
                                                       net.L[i], net.L[i + 1], net.ac[i + 1], sc.Y[i + 1], nullptr,
                                                       nullptr, nullptr, d_done);
            }
            ++*launches;
        }
        if (!fvp) {
            // seed G_K and the per-sample LogStd gradient (kept in RY[0] as a [rows x A] matrix)
            const int n = rows * A;
            k_pg_seed<<<cdiv(n, 256), 256, 0, st>>>(d_mean + c0 * A, d_action + c0 * A, d_adv + c0, d_theta + net.logstd_off,
                                                    sc.Y[K], net.ac[K], rows, A, sc.G[K & 1], sc.RY[0]);
            ++*launches;
            dim3 g1(cdiv(1, BM) * cdiv(A, BN), sc.nslices);
            k_chain_outer<<<g1, NT, 0, st>>>(nullptr, sc.RY[0], rows, 0, A, per_slice, cdiv(A, BN), sc.partial, net.P,
                                             net.logstd_off, accumulate, d_done);
            ++*launches;
        }
        // ---- backward + outer products ----
        for (int i = K; i >= 1; --i) {
            const double *Yprev = (i == 1) ? d_obs + c0 * net.L[0] : sc.Y[i - 1];
            const int M0 = net.L[i - 1], N = net.L[i];
            const int tiles_m = cdiv(M0 + 1, BM), tiles_n = cdiv(N, BN);
            dim3 go(tiles_m * tiles_n, sc.nslices);
            k_chain_outer<<<go, NT, 0, st>>>(Yprev, sc.G[i & 1], rows, M0, N, per_slice, tiles_n, sc.partial, net.P,
                                             net.w_off[i - 1], accumulate, d_done);
            ++*launches;
            if (i > 1) {
                dim3 gb(cdiv(rows, BM), cdiv(M0, BN));
                k_chain_bwd<<<gb, NT, 0, st>>>(sc.G[i & 1], d_theta + net.w_off[i - 1], sc.Y[i - 1], rows, N, M0,
                                               net.ac[i - 1], sc.G[(i - 1) & 1], d_done);
                ++*launches;
            }
        }
    }
    launch_reduce_partials(sc.partial, sc.nslices, net.P, d_zsum, d_done, p2p, st, launches);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int chain_forward(const NetDesc &net, const ChainScratch &sc, const double *d_theta, const double *d_obs,
                  size_t nsamples, double *d_mean_out, cudaStream_t st, long long *launches) {
    const int K = net.K, A = net.L[K];
    for (size_t c0 = 0; c0 < nsamples; c0 += sc.chunk) {
        const int rows = (int)((nsamples - c0 < (size_t)sc.chunk) ? nsamples - c0 : sc.chunk);
        for (int i = 0; i < K; ++i) {
            const double *Yin = (i == 0) ? d_obs + c0 * net.L[0] : sc.Y[i];
            double *Yout = (i == K - 1) ? d_mean_out + c0 * A : sc.Y[i + 1];
            dim3 grid(cdiv(rows, BM), cdiv(net.L[i + 1], BN));
            k_chain_fwd<false><<<grid, NT, 0, st>>>(Yin, nullptr, d_theta + net.w_off[i], nullptr, rows, net.L[i],
                                                   net.L[i + 1], net.ac[i + 1], Yout, nullptr, nullptr, nullptr, nullptr);
            ++*launches;
        }
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
