// Internal declarations shared by the CUDA translation units of libtrpo_b200.so (not part of the C-ABI).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#define TRPO_MAX_LAYERS 16

#ifdef __cplusplus
#include <atomic>
// cudaFuncSetAttribute applies to the CURRENT device only; a process may drive several GPUs (one thread per GPU), so the
// "already configured" memo of a kernel set is kept per device.
struct DeviceOnce {
    std::atomic<unsigned long long> mask{0};
    bool pending() const { int d = 0; cudaGetDevice(&d); return !((mask.load() >> (d & 63)) & 1ull); }
    void mark() { int d = 0; cudaGetDevice(&d); mask.fetch_or(1ull << (d & 63)); }
};
#endif

// Network description in device-friendly form. The flat parameter vector (TRPO_FVP.c:704-725) is, per weight layer i,
// an augmented row-major matrix [(L_i + 1) x L_{i+1}] at offset w_off[i]: L_i rows of W[i] followed by the row B[i].
struct NetDesc {
    int K;                           // weight layers = NumLayers - 1
    int L[TRPO_MAX_LAYERS];          // L[0..K]
    int w_off[TRPO_MAX_LAYERS];      // offset of W[i]; B[i] is at w_off[i] + L[i]*L[i+1]
    int logstd_off;
    int P;
    char ac[TRPO_MAX_LAYERS];        // ac[i] = activation producing layer i (i >= 1)
};

// Device-side CG state block (one per context).
struct CgState {
    double rdotr;
    double pdotz;
    double xnorm;
    int    done;        // set when rdotr < ResidualTh: the remaining iterations become no-ops
    int    iters;       // FVPs executed
};
// The per-iteration trace ("Residual Norm" / "Soln Norm" of TRPO_CG.c:56) lives in a separate device buffer of
// 2 * trace_cap doubles (rdotr[0..cap), xnorm[0..cap)) that the context grows to MaxIter + 2: the reference places no
// bound on MaxIter (TRPO_CG.c:45).

// Scratch for the GEMM-chain path: activations of one sample chunk, row-major [chunk x L_i].
struct ChainScratch {
    double *Y[TRPO_MAX_LAYERS];      // Y[i] for i = 1..K (Y[0] is the observation matrix itself)
    double *RY[2];                   // R{y} ping-pong, [chunk x maxL]
    double *G[2];                    // R-gradient / gradient ping-pong, [chunk x maxL]
    int chunk;                       // samples per chunk
    int nslices;                     // split-K slices of the outer-product GEMM
    double *partial;                 // [nslices x P] un-normalised partial sums, fixed-order reduced
    double *wperm, *vperm;           // row-permuted copies of the weights / the direction for the TMA-fed forward kernel (chain_tma_perm_total)
    double *tailw;                   // TMA-fed tail kernel: permuted, padded W / VW of the last layer and the W^T image (chain_tma_tail_doubles)
    // streamed staging of the observation matrix: piece i of piece_rows rows has landed when piece_events[i] has completed
    // (recorded on the copy stream); the chunk loop makes the compute stream wait only for the pieces a chunk touches
    const cudaEvent_t *piece_events;
    int n_pieces;
    size_t piece_rows;
};

enum ChainMode { CHAIN_FVP = 0, CHAIN_PG = 1 };

// Peer-memory all-reduce state (one process per GPU, buffers exchanged as CUDA IPC handles). Every rank owns
//   slots[2][world][P] doubles  and  flags[2][world] sequence numbers;
// rank s PUSHES its un-normalised FVP sum into slots[parity][s] of every rank over NVLink and then publishes the
// sequence number into flags[parity][s]; the consumer (CG update / FVP finalise) waits on its LOCAL flags and sums the
// world slots in rank order, so every rank forms bitwise the same sum. world == 0 means "not in use".
#define TRPO_MAX_RANKS 8
struct P2PComm {
    int world, rank;
    double *slots[TRPO_MAX_RANKS];               // slots[r]: base of rank r's slot area (peer-mapped; own for r == rank)
    unsigned long long *flags[TRPO_MAX_RANKS];   // flags[r]: base of rank r's flag area
    unsigned long long *ll[TRPO_MAX_RANKS];      // persistent solve kernel: tagged words [2][world][P][2] (low-latency exchange), or NULL
    unsigned long long *seq_dev;                 // completed all-reduces (local, advanced by the consumer)
    unsigned int *block_counter;                 // last-block detection of the push kernel (local)
    int *error;                                  // set when a wait timed out (local)
    int P;
};

// ---- gemm_chain.cu ------------------------------------------------------------------------------------------------
// Enqueue the whole un-normalised sum  zsum[0..P) = sum_n per-sample [RGW,RGB,...] (FVP) or [GW,GB,...,GLogStd] (PG).
// The LogStd block of zsum is written by the finalise step, not here (FVP) / by the seed kernel (PG).
int chain_accumulate(const NetDesc &net, const ChainScratch &sc, ChainMode mode,
                     const double *d_theta, const double *d_v, const double *d_inv_var,
                     const double *d_obs, const double *d_mean, const double *d_action, const double *d_adv,
                     size_t nsamples, double *d_zsum, const int *d_done, const P2PComm *p2p, cudaStream_t st,
                     long long *launches);
// Forward only: writes the policy mean of every sample of [s0, s0+n) into d_out [n x A].
int chain_forward(const NetDesc &net, const ChainScratch &sc, const double *d_theta, const double *d_obs,
                  size_t nsamples, double *d_mean_out, cudaStream_t st, long long *launches);
size_t chain_scratch_bytes(const NetDesc &net, int chunk, int nslices);

// gemm_chain_tma.cu: TMA-fed variants of the single-product GEMM-chain kernels (operands in SWIZZLE_128B boxes, mbarrier completion)
bool chain_tma_enabled();
bool chain_tma_outer_eligible(const double *Yprev, const double *G, int M0, int N, int ldg);
bool chain_tma_bwd_eligible(const double *Gin, const double *W, const double *Yprev, const double *Gout, int Kd, int N);
int chain_tma_tiles_m(int M0);
int chain_tma_outer(const double *Yprev, const double *G, int ldg, int rows, int M0, int N, int per_slice, int tiles_n, int nslices,
                    double *partial, int P, int out_off, int accumulate, const int *done, cudaStream_t st);
int chain_tma_bwd(const double *Gin, const double *W, const double *Yprev, int rows, int Kd, int N, char act_prev, double *Gout,
                  const int *done, cudaStream_t st);
bool chain_tma_fwd_eligible(const double *Yin, const double *RYin, const double *Yout, const double *RYout, const double *Gout, int Kd, int N);
size_t chain_tma_perm_doubles(int Kd, int N);         // one layer's permuted copy: Kd rounded up to 16 rows
size_t chain_tma_perm_offset(const NetDesc &net, int layer);     // layer == K: the total
void chain_tma_permute_rows(const double *W, double *Wp, int Kd, int N, cudaStream_t st);
size_t chain_tma_tail_doubles(int H);
bool chain_tma_tail_eligible(const double *Y, const double *RY, const double *Gprev, int H, int A);
void chain_tma_tail_prepare(const double *W, const double *VW, double *scratch, int H, int A, cudaStream_t st);
int chain_tma_tail(const double *Y, const double *RY, const double *VW, const double *scratch, int rows, int H, int A, char act_prev,
                   double d3, const double *inv_var, double *GK, int ldgk, double *Gprev, const int *done, cudaStream_t st);
int chain_tma_fwd(const double *Yin, const double *RYin, const double *Wp, const double *Vp, const double *W, const double *VW,
                  int rows, int Kd, int N, char act, double *Yout, double *RYout, double *Gout, const double *inv_var,
                  const int *done, cudaStream_t st);

// ---- gemm_chain_f32.cu (optional FP32 mode) ------------------------------------------------------------------------
struct ChainScratchF32 {
    float *Y[TRPO_MAX_LAYERS];
    float *RY[2];
    float *G[2];
    int chunk, nslices;
    float *partial;                  // [nslices x P] FP32 partial sums
    // tcgen05 forward layers (tc_fwd_f32.cu): TF32 remainders (x - trunc_tf32(x)) of the activations, and per layer the
    // transposed weights / directions [hi, lo, V hi, V lo] x [N x Kd]
    float *Ylo[TRPO_MAX_LAYERS];
    float *RYlo[2];
    float *wt[TRPO_MAX_LAYERS];
    const float *obs_lo;             // remainders of the observation matrix (whole shard), or NULL
};
size_t chain_f32_scratch_floats(const NetDesc &net, int chunk, int nslices);
void chain_f32_convert(const double *d_src, float *d_dst, size_t n, cudaStream_t st, long long *launches);
int chain_f32_accumulate(const NetDesc &net, const ChainScratchF32 &sc, const float *f_theta, const float *f_v,
                         const float *f_inv_var, const float *f_obs, size_t nsamples, double *d_zsum,
                         const int *d_done, const P2PComm *p2p, cudaStream_t st, long long *launches);

// ---- tc_fwd_f32.cu (FP32 mode: forward R-op layer on tcgen05 / TMEM / TMA) --------------------------------------------
bool tc_fwd_eligible(int Kd, int N);
void tc_lo_split(const float *x, float *lo, size_t n, cudaStream_t st, long long *launches);
void tc_prep_weights(const float *W, const float *VW, int Kd, int N, float *wt4, cudaStream_t st, long long *launches);
int tc_fwd_layer(const float *Yin, const float *Ylo_in, const float *RYin, const float *RYlo_in, const float *wt4, const float *bias,
                 const float *vbias, int rows, int Kd, int N, char act, float *Yout, float *Ylo_out, float *RYout, float *RYlo_out,
                 const int *done, cudaStream_t st, long long *launches);

// ---- fvp_fused.cu -------------------------------------------------------------------------------------------------
// Fused DMMA kernel for 4-layer nets whose padded weights fit in shared memory. Returns 0 if it handled the launch,
// 1 if the shape is not eligible.
bool fused_eligible(const NetDesc &net);
int  fused_partial_rows();     // number of per-CTA partial rows the fused kernel writes
int  fused_fvp_accumulate(const NetDesc &net, const double *d_theta, const double *d_v, const double *d_inv_var,
                          const double *d_obs, size_t nsamples, double *d_partial, double *d_zsum,
                          const int *d_done, const P2PComm *p2p, const int *stream_ready, size_t stream_chunk,
                          int *stream_error, cudaStream_t st, long long *launches);

// d_mean == NULL: the kernel computes the network output itself and stores it to d_mean_out (if not NULL)
int  fused_pg_accumulate(const NetDesc &net, const double *d_theta, const double *d_inv_std, const double *d_obs,
                         const double *d_mean, const double *d_action, const double *d_adv, size_t nsamples,
                         double *d_partial, double *d_zsum, double *d_mean_out, cudaStream_t st, long long *launches);

// Whole CG solve as one persistent cooperative kernel (fused-eligible shapes, single GPU or peer-memory exchange).
// Returns 0 if enqueued, 1 if the shape / batch is not eligible (caller falls back to the per-iteration launches), -1 on error.
int fused_cg_solve(const NetDesc &net, const double *d_theta, const double *d_inv_var, const double *d_obs, size_t nsamples,
                   double n_total, double *d_partial, const double *d_b, double *d_x, double *d_r, double *d_p, double *d_z,
                   double *d_zsum, double *d_dots, unsigned int *d_gbar, CgState *d_state, double *d_trace, int trace_cap,
                   size_t max_iter, double residual_th, double damping, const P2PComm *p2p, const int *stream_ready,
                   size_t stream_chunk, int *stream_error, unsigned long long *d_timeline, cudaStream_t st, long long *launches);

// ---- cg_kernels.cu ------------------------------------------------------------------------------------------------
// p2p != NULL: the fixed-order row sum is pushed straight into every rank's slot (fused reduce + all-reduce send)
void launch_reduce_partials(const double *d_partial, int rows, int P, double *d_zsum, const int *d_done,
                            const P2PComm *p2p, cudaStream_t st, long long *launches);
// z = zsum/N + damping*v (LogStd block: 2v + damping*v), standalone FVP finalise (TRPO_FVP.c:928-931)
void launch_fvp_finalise(const double *d_zsum, const double *d_v, double *d_out, int P, int logstd_off,
                         double n_total, double damping, const P2PComm *p2p, cudaStream_t st, long long *launches);
void launch_cg_init(const double *d_b, double *d_x, double *d_r, double *d_p, int P, double residual_th,
                    CgState *d_state, double *d_trace, int trace_cap, cudaStream_t st, long long *launches);
void launch_cg_update(const double *d_zsum, double *d_x, double *d_r, double *d_p, double *d_z, int P, int logstd_off,
                      double n_total, double damping, double residual_th, CgState *d_state, double *d_trace, int trace_cap,
                      const P2PComm *p2p, cudaStream_t st, long long *launches);
// out[0] = sum_i a[i]*b[i] with the same fixed-order reduction (used for shs, gnorm, b.x)
void launch_dot(const double *d_a, const double *d_b, int n, double *d_out, cudaStream_t st, long long *launches);
void launch_axpby(double *d_out, const double *d_x, double a, const double *d_y, double b, int n,
                  cudaStream_t st, long long *launches);
// surrogate loss sum_n exp(lld_n) * Adv_n (TRPO_Update.c:969-983), fixed-order reduction into d_out[0]
void launch_surrogate(const double *d_mean_new, const double *d_mean_old, const double *d_action, const double *d_adv,
                      const double *d_std_old, const double *d_logstd_new, int A, size_t nsamples,
                      double *d_block_partials, double *d_out, cudaStream_t st, long long *launches);
void launch_sum(const double *d_a, size_t n, double *d_block_partials, double *d_out, cudaStream_t st, long long *launches);

// ---- rollout_kernels.cu (rows f-3 / f-4: advantage estimation and the baseline objective) ------------------------------
// out[n][0..O) = obs[n][0..O), out[n][O] = (n mod EpLen)/EpLen: the baseline network's input (TRPO_Baseline.c:98-103)
void launch_vf_augment(const double *d_obs, size_t n, int O, size_t ep_len, double *d_out, cudaStream_t st, long long *launches);
// discounted return and GAE(gamma, lam) advantage of every episode (TRPO_Lightweight.c:565-641), un-standardised
void launch_gae(const double *d_reward, const double *d_baseline, size_t num_ep, int ep_len, double gamma, double lam,
                double *d_ret, double *d_adv, cudaStream_t st, long long *launches);
// out[0] = sum_i (a[i] - c_i)^2, c_i = b[i] if b != NULL else *d_shift_sum / shift_div; fixed-order
void launch_sqdiff(const double *d_a, const double *d_b, const double *d_shift_sum, double shift_div, size_t n,
                   double *d_block_partials, double *d_out, cudaStream_t st, long long *launches);
// a = (a - sum/N) / sqrt(sqdev/N)  (TRPO_Lightweight.c:645-653)
void launch_standardise(double *d_a, size_t n, const double *d_sum, const double *d_sqdev, double n_total, cudaStream_t st,
                        long long *launches);
void launch_fill(double *d_a, double v, size_t n, cudaStream_t st, long long *launches);
// Rollout producer for the lightweight arm simulator (TRPO_Lightweight.c:349-540): one warp per episode. d_draws = raw rand()
// values in the reference's order, or NULL for the counter-based generator. Returns 1 if the network is not a
// 15-...-3 policy with hidden widths <= 32.
int launch_arm_rollout(const NetDesc &net, const double *d_theta, const int *d_draws, unsigned long long seed, size_t num_ep,
                       int ep_len, double *d_obs, double *d_mean, double *d_action, double *d_reward, cudaStream_t st,
                       long long *launches);
