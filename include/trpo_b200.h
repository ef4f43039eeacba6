/*
 * trpo_b200.h -- C-ABI of the B200-native natural-gradient solve (FVP + CG + TRPO update).
 *
 * Drop-in boundary for the reference's FPGA host files:
 *   FVP_FPGA  /root/reference/src/TRPO_FVP_FPGA.c:13   (prototype /root/reference/src/include/TRPO.h:98)
 *   CG_FPGA   /root/reference/src/TRPO_CG_FPGA.c:13    (prototype TRPO.h:101)
 * and GPU counterparts of the CPU entry points they are tested against:
 *   FVP / FVPFast  /root/reference/src/TRPO_FVP.c:11,548   (TRPO.h:88,92)
 *   CG             /root/reference/src/TRPO_CG.c:11        (TRPO.h:95)
 *   TRPO_Update    /root/reference/src/TRPO_Update.c:10    (TRPO.h:104)
 *
 * Everything here is plain C: pointers, sizes and doubles. No torch / C++ types cross the boundary.
 * All functions return the reference's convention where they mirror a reference function
 * (elapsed compute seconds >= 0, or -1.0 after printing "[ERROR] ..." on stderr), and
 * 0 / negative error code for the context API (trpo_last_error() gives the message).
 */
#ifndef TRPO_B200_H
#define TRPO_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * TRPOparam: identical field order and ABI to /root/reference/src/include/TRPO.h:6-49
 * (11 x 8 bytes on LP64, passed BY VALUE). If the reference header was included first, reuse its type.
 * The FPGA-only fields PaddedLayerSize / NumBlocks are ignored (callers leave them uninitialised).
 * ---------------------------------------------------------------------------------------------- */
#ifndef TRPO_H
typedef struct {
    char   *ModelFile;        /* W, B per layer then LogStd: text, TRPO_FVP.c:677-696 */
    char   *BaselineFile;     /* unused on this path */
    char   *ResultFile;       /* unused on this path */
    char   *DataFile;         /* rows Mean[A] Std[A] Observ[O] Action[A] Advantage: TRPO_FVP.c:740-759 */
    size_t  NumLayers;        /* input + hidden + output, e.g. 4 */
    char   *AcFunc;           /* NumLayers chars, AcFunc[0] unused: 'l','t','o','s' (TRPO_FVP.c:806-834) */
    size_t *LayerSize;        /* NumLayers entries */
    size_t  NumSamples;
    double  CG_Damping;
    size_t *PaddedLayerSize;  /* FPGA only -- ignored */
    size_t *NumBlocks;        /* FPGA only -- ignored */
} TRPOparam;
#endif

/* ------------------------------------------------------------------------------------------------
 * File-based drop-in entry points (reference signatures). Each call loads ModelFile / DataFile, stages the
 * batch on the GPU, runs, copies the P-length result back and releases everything, like the FPGA hosts
 * (max_load ... max_unload, TRPO_FVP_FPGA.c:197-198,370-371). Return: seconds spent in the device compute
 * section (file parsing excluded, as in TRPO_FVP.c:768-769,933-934), or -1.0 on failure.
 * ---------------------------------------------------------------------------------------------- */

/* Replaces FVP_FPGA (TRPO_FVP_FPGA.c:13). Result = (1/N) sum_n F_n * Input + CG_Damping * Input. */
double FVP_GPU(TRPOparam param, double *Result, double *Input);

/* Replaces CG_FPGA (TRPO_CG_FPGA.c:13). Solves (F + damping I) x = b; prints the reference's
 * "CG Iter[%zu] Residual Norm=%.12e, Soln Norm=%.12e" lines (TRPO_CG.c:56). NumThreads is accepted and ignored. */
double CG_GPU(TRPOparam param, double *Result, double *b, size_t MaxIter, double ResidualTh, size_t NumThreads);

/* GPU counterpart of TRPO_Update (TRPO_Update.c:10): policy gradient, CG, shs FVP, line search. */
double TRPO_Update_GPU(TRPOparam param, double *Result, size_t NumThreads);

/* GPU counterpart of TRPO_Lightweight / TRPO_Lightweight_FPGA (TRPO_Lightweight.c:12, TRPO_Lightweight_FPGA.c:12,
 * prototypes TRPO.h:110,113): the whole training loop on the lightweight arm simulator -- rollouts, return / GAE,
 * baseline fit, TRPO update -- for NumIter iterations; reads param.ModelFile / param.BaselineFile, writes
 * `<param.ResultFile>%03d.txt` every 100 iterations and after the last one, prints the reference's log lines, returns the
 * loop's wall-clock seconds or -1. The rollouts consume the C library's rand() stream after srand(0) exactly as the
 * reference's loop does. The L-BFGS of the baseline fit is the CALLER's libLBFGS (the reference vendors src/lbfgs.c):
 * `lbfgs` / `lbfgs_parameter_init` are taken from the process's loaded symbols or from the shared object named by the
 * environment variable TRPO_LBFGS_LIB. _ex exposes the constants the reference hard-codes (:32-33,52-55). */
double TRPO_Lightweight_GPU(TRPOparam param, const int NumIter, const size_t NumThreads);
double TRPO_Lightweight_GPU_ex(TRPOparam param, const int NumIter, size_t NumEpBatch, size_t EpLen, double gamma, double lam);

/* The reference's own symbol names, so that Test_FVP_FPGA / Test_CG_FPGA / Test_TRPO_Lightweight_FPGA
 * (TRPOCpuCode.c:189,273,457) link unchanged.
 * Compiled only into libtrpo_b200_dropin.so (they would clash with the MaxCompiler build otherwise). */
double FVP_FPGA(TRPOparam param, double *Result, double *Input);
double CG_FPGA(TRPOparam param, double *Result, double *b, size_t MaxIter, double ResidualTh, size_t NumThreads);
double TRPO_Lightweight_FPGA(TRPOparam param, const int NumIter, const size_t NumThreads);

/* NumParamsCalc (TRPO_Util.c:7-17) under a non-clashing name. */
size_t trpo_num_params(const size_t *LayerSize, size_t NumLayers);

/* ------------------------------------------------------------------------------------------------
 * Persistent context API: pay file parsing / H2D staging once per rollout batch, not once per call
 * (the reference re-reads both text files on EVERY FVP, TRPO_FVP.c:670-762).
 * One context = one GPU = one host thread at a time.
 * ---------------------------------------------------------------------------------------------- */
typedef struct trpo_ctx trpo_ctx;

/* FP64: the reference's arithmetic (parity <= 1e-10). FP32: optional mode for the FVP inside CG -- FP32 storage, 3xTF32
 * tensor-core products, FP32 per-slice sums, FP64 reduction and CG; stated tolerance 1e-4 relative (norm-wise) on the FVP.
 * The policy gradient and the line search of trpo_ctx_update stay FP64 in either mode. */
enum { TRPO_PRECISION_FP64 = 0, TRPO_PRECISION_FP32 = 1 };

/* Kernel path selection (trpo_ctx_set_path). AUTO picks the fused DMMA kernel when the network fits it. */
enum { TRPO_PATH_AUTO = 0, TRPO_PATH_GEMM_CHAIN = 1, TRPO_PATH_FUSED = 2 };

/* Diagnostics of the last CG / update on a context. */
typedef struct {
    int    cg_iters;            /* FVPs executed by the last CG */
    double cg_rdotr[34];        /* value printed as "Residual Norm" at iteration i: the first 34 entries of the trace */
    double cg_xnorm[34];        /* value printed as "Soln Norm"; trpo_ctx_get_cg_trace returns the whole trace      */
    double shs, lm, gnorm, fval;
    int    ls_steps, ls_accepted;
    double ls_actual[16], ls_expected[16], ls_ratio[16];
} trpo_info;

const char *trpo_last_error(void);

/* device < 0: use the current CUDA device. Returns NULL on failure. */
trpo_ctx *trpo_ctx_create(const size_t *LayerSize, const char *AcFunc, size_t NumLayers, int device, int precision);
void      trpo_ctx_destroy(trpo_ctx *ctx);

size_t trpo_ctx_num_params(const trpo_ctx *ctx);
/* Use an external CUDA stream (cudaStream_t as void*) for all work of this context; NULL = own stream. */
int    trpo_ctx_set_stream(trpo_ctx *ctx, void *cuda_stream);
void  *trpo_ctx_get_stream(const trpo_ctx *ctx);
int    trpo_ctx_set_path(trpo_ctx *ctx, int path);
int    trpo_ctx_get_path(const trpo_ctx *ctx);      /* the path the last launch actually used */
/* GEMM-chain path: samples per pass over the kernel chain (rounded up to a multiple of 128). 0 = automatic: as many whole
 * waves of CTAs as fit 4 GB of activation scratch. A batch larger than the chunk is processed in several passes whose
 * per-slice partial sums accumulate in place (the sum TRPO_FVP.c:903-921 forms sample by sample). */
int    trpo_ctx_set_chunk(trpo_ctx *ctx, size_t chunk_samples);
size_t trpo_ctx_get_chunk(const trpo_ctx *ctx);     /* chunk of the last GEMM-chain launch (0 before the first) */
int    trpo_ctx_sync(trpo_ctx *ctx);
/* 1 if the last trpo_ctx_cg / _cg_device ran as the single persistent cooperative kernel (shapes the fused kernels take, one
 * GPU or the peer-memory exchange), 0 if it ran as per-iteration launches. The environment variable TRPO_NO_FUSED_SOLVE
 * forces the latter. */
int    trpo_ctx_solve_kernel_used(const trpo_ctx *ctx);
/* Diagnostics: phase time stamps of the persistent solve kernel. Call with max_iters > 0 and out == NULL to enable (0 disables);
 * after a solve call with out != NULL to read max_iters x 8 nanosecond stamps of CTA 0 per CG iteration: pass start, own pass
 * done, all passes done, slice sums formed, peers' slices arrived, p.z complete, r.r complete, new direction published. */
int    trpo_ctx_solve_timeline(trpo_ctx *ctx, size_t max_iters, unsigned long long *out);
/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
long long trpo_ctx_launch_count(const trpo_ctx *ctx);

/* Optional CUDA-event timing of the dominant kernel (the per-sample FVP sum) on the context stream.
 * enable != 0 starts recording an event pair around every FVP-sum launch (up to 4096 pairs, then it stops recording);
 * trpo_ctx_kernel_time_ms synchronises, returns the accumulated milliseconds, stores the number of timed launches
 * in *launches (may be NULL) and resets the accumulator. */
int    trpo_ctx_kernel_timing(trpo_ctx *ctx, int enable);
double trpo_ctx_kernel_time_ms(trpo_ctx *ctx, int *launches);

/* Model: theta = flat [W0,B0,...,LogStd] (TRPO_FVP.c:704-725), P doubles on the host. */
int trpo_ctx_set_model(trpo_ctx *ctx, const double *theta);

/* Rollout batch (host pointers, row-major). When Observ is PINNED host memory and the batch is large, the copy is issued
 * asynchronously in chunks and the first fused FVP after this call overlaps it (the kernel polls a chunk counter the
 * copy engine advances); the caller must keep Observ unchanged until the next synchronising call returns. Std has A entries (the data file's Std columns, last row wins,
 * TRPO_FVP.c:746-748). Mean/Action/Advantage may be NULL when only FVP/CG are used.
 * In multi-GPU mode every rank passes ITS shard (NumSamples = local count). */
int trpo_ctx_set_batch(trpo_ctx *ctx, size_t NumSamples, const double *Observ, const double *Std,
                       const double *Mean, const double *Action, const double *Advantage);
/* Same with DEVICE pointers (adopted, not copied; must stay valid until the next set_batch). */
int trpo_ctx_set_batch_device(trpo_ctx *ctx, size_t NumSamples, const double *dObserv, const double *Std_host,
                              const double *dMean, const double *dAction, const double *dAdvantage);

/* Synchronous host-buffer calls (H2D of the P-length input, compute, D2H of the result). */
int trpo_ctx_fvp(trpo_ctx *ctx, const double *Input, double *Result, double CG_Damping);
int trpo_ctx_cg(trpo_ctx *ctx, const double *b, double *Result, size_t MaxIter, double ResidualTh, double CG_Damping);
int trpo_ctx_policy_gradient(trpo_ctx *ctx, double *b_out);
/* Policy mean of every staged sample (ordinary forward pass, TRPO_Update.c:259-291 / the Mean column a rollout producer
 * writes, TRPO_Lightweight_FPGA.c:548-556): Mean_out is NumSamples x A on the host. */
int trpo_ctx_forward(trpo_ctx *ctx, double *Mean_out);
int trpo_ctx_update(trpo_ctx *ctx, double *Result, double CG_Damping);
int trpo_ctx_get_info(const trpo_ctx *ctx, trpo_info *info);
/* Whole per-iteration trace of the last CG (entries 0..cg_iters: what TRPO_CG.c:56 prints); MaxIter is unbounded as in the
 * reference (TRPO_CG.c:45). Copies at most max_entries values into each array (either may be NULL); returns the count. */
int trpo_ctx_get_cg_trace(const trpo_ctx *ctx, double *rdotr_out, double *xnorm_out, size_t max_entries);

/* Asynchronous device-buffer calls: enqueue on the context stream and return (inputs/outputs are device
 * pointers to P doubles). The CG state never leaves the device; trpo_ctx_get_info after trpo_ctx_sync. */
int trpo_ctx_fvp_device(trpo_ctx *ctx, const double *dInput, double *dResult, double CG_Damping);
int trpo_ctx_cg_device(trpo_ctx *ctx, const double *db, double *dResult, size_t MaxIter, double ResidualTh, double CG_Damping);

/* Device scratch helpers for C callers that do not link the CUDA runtime themselves. */
double *trpo_device_alloc(size_t n_doubles);
void    trpo_device_free(double *p);
int     trpo_memcpy_h2d(double *dst_dev, const double *src_host, size_t n_doubles);
int     trpo_memcpy_d2h(double *dst_host, const double *src_dev, size_t n_doubles);
/* Page-locked host arrays: with a pinned Observ source trpo_ctx_set_batch streams the copy under the first FVP. */
double *trpo_host_alloc_pinned(size_t n_doubles);
void    trpo_host_free_pinned(double *p);

/* ------------------------------------------------------------------------------------------------
 * The training loop around the update: what TRPO_Lightweight.c:541-694 does
 * between the rollouts and the TRPO update, on the batch already staged in HBM.
 * ---------------------------------------------------------------------------------------------- */
/* Stage one batch of rollouts: NumEpBatch episodes of EpLen steps, row = ep*EpLen + step; Reward has one entry per row
 * (the arrays TRPO_Lightweight.c:112-118 fills, or the streams TRPO_RunLightweight returns,
 * TRPO_Lightweight_FPGA.c:548-556). The advantage is produced on the device by trpo_vf_advantage. */
int trpo_ctx_set_rollout(trpo_ctx *ctx, size_t NumEpBatch, size_t EpLen, const double *Observ, const double *Std,
                         const double *Mean, const double *Action, const double *Reward);

/* Rollout PRODUCER on the device: the reference's lightweight arm simulator (TRPO_Lightweight.c:349-540; the role
 * TRPO_RunLightweight plays on the FPGA, TRPO_Lightweight_FPGA.c:548-556) driven by the model last given to
 * trpo_ctx_set_model, one warp per episode; Observ / Mean / Action / Reward land in the context's batch exactly as after
 * trpo_ctx_set_rollout, Std = exp(LogStd). RandDraws (host) holds the raw rand() values in the reference's order --
 * per episode 3 (object position), then per step 2 per action component: NumEpBatch * (3 + 6 * EpLen) ints -- which makes
 * the batch the one the reference would have produced from the same stream; NULL selects a counter-based generator
 * seeded by Seed. The policy must be 15-...-3 with hidden widths <= 32. */
int trpo_ctx_rollout_arm(trpo_ctx *ctx, size_t NumEpBatch, size_t EpLen, const int *RandDraws, unsigned long long Seed);
/* Copy the staged rollout back to the host (any pointer may be NULL). */
int trpo_ctx_get_rollout(trpo_ctx *ctx, double *Observ, double *Mean, double *Action, double *Reward);

/* Value-function ("baseline") network on the policy context's device and stream. LayerSizeBase[0] must be
 * ObservSpaceDim + 1 (observation followed by step/EpLen), the last layer 1 (TRPO_Lightweight.c:36, TRPO_Baseline.c:98-103).
 * Its parameter vector x is [W0,B0,...] WITHOUT a LogStd tail: trpo_vf_num_params == NumParamsCalc(LayerSizeBase) - 1
 * (TRPO_Lightweight.c:44). Shares the policy context's communicator in multi-GPU mode. */
typedef struct trpo_vf trpo_vf;
trpo_vf *trpo_vf_create(trpo_ctx *policy, const size_t *LayerSizeBase, const char *AcFunc, size_t NumLayers);
void     trpo_vf_destroy(trpo_vf *vf);
size_t   trpo_vf_num_params(const trpo_vf *vf);
/* (Re)build the time-augmented observation matrix from the batch staged in the policy context; call after every
 * trpo_ctx_set_batch / set_rollout (trpo_vf_advantage does it itself). NumSamples must be a multiple of EpLen. */
int trpo_vf_bind_batch(trpo_vf *vf, size_t EpLen);
int trpo_vf_set_target(trpo_vf *vf, const double *Target);                 /* NumSamples doubles on the host */
int trpo_vf_predict(trpo_vf *vf, const double *x, double *Predict_out);    /* TRPO_Lightweight.c:582-625 */
/* Return, GAE(gamma, lam) advantage and its standardisation (TRPO_Lightweight.c:565-653) with the baseline predicted
 * from x. The standardised advantage becomes the policy context's Advantage (ready for trpo_ctx_update), the return
 * becomes this network's target (ready for trpo_vf_evaluate). Return_out / Advantage_out may be NULL. */
int trpo_vf_advantage(trpo_vf *vf, const double *x, double gamma, double lam, double *Return_out, double *Advantage_out);
/* libLBFGS objective callback, a drop-in for the reference's `evaluate` (TRPO_Baseline.c:29, lbfgs.h lbfgs_evaluate_t):
 *     lbfgs(PaddedParams, x, &fx, trpo_vf_evaluate, NULL, vf, &param);
 * returns 0.01*MSE + 0.001*|x|^2 and writes g = d/dx (n >= trpo_vf_num_params; the padding of g is zeroed).
 * After a failure (see trpo_last_error; the reference returns -1 for an unsupported activation, which L-BFGS would take for a
 * very good objective value) it returns +inf with a zero gradient and records the failure: see trpo_vf_failed. */
double trpo_vf_evaluate(void *vf, const double *x, double *g, const int n, const double step);
/* libLBFGS cannot tell a failed evaluation from a low objective, so a failure is also recorded on the network: the callback
 * then returns +inf with a zero gradient, and this returns 1 once (and clears the record). Check it after lbfgs() returns.
 * Lifetime: a trpo_vf borrows its policy context's stream and communicator -- destroy it BEFORE the policy context and
 * re-bind (trpo_vf_bind_batch / trpo_vf_advantage) after trpo_ctx_set_stream on the policy context. */
int trpo_vf_failed(trpo_vf *vf);

/* Binary replacement for the text data file (TRPO_FVP.c:731-762 re-parses N x (3A+O+1) decimal numbers on every call):
 * a 64-byte header {"TRPOB200", version, flags, N, O, A} followed by Std[A], Observ[N*O], Mean[N*A], Action[N*A],
 * Advantage[N] as little-endian doubles. FVP_GPU / CG_GPU / TRPO_Update_GPU accept either format in param.DataFile. */
typedef struct {
    char               magic[8];        /* "TRPOB200" */
    unsigned int       version;         /* 1 */
    unsigned int       flags;           /* bit 0: Mean / Action / Advantage sections present */
    unsigned long long NumSamples, ObservSpaceDim, ActionSpaceDim;
    unsigned long long reserved[3];
} trpo_batch_file_header;               /* 64 bytes */
int trpo_batch_file_write(const char *path, size_t NumSamples, size_t ObservSpaceDim, size_t ActionSpaceDim,
                          const double *Observ, const double *Std, const double *Mean, const double *Action,
                          const double *Advantage);
int trpo_batch_file_from_text(const char *text_path, const char *bin_path, size_t NumSamples, size_t ObservSpaceDim,
                              size_t ActionSpaceDim);
/* Read the header of a binary batch file; returns -1 (and leaves *hdr zeroed) if path is not one. */
int trpo_batch_file_probe(const char *path, trpo_batch_file_header *hdr);
/* Read the first NumSamples rows into caller-allocated host arrays (Mean/Action/Advantage may be NULL). */
int trpo_batch_file_read(const char *path, size_t NumSamples, double *Observ, double *Std, double *Mean, double *Action,
                         double *Advantage);
/* Stage the first NumSamples rows (0 = all) of a binary batch file through pinned buffers. */
int trpo_ctx_set_batch_file(trpo_ctx *ctx, const char *path, size_t NumSamples);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU: one process (or thread) per GPU. Samples are sharded; each FVP ends with ONE all-reduce of the
 * P-length un-normalised sum (ncclDouble, ncclSum) before the replicated CG update.
 * Rank 0 creates the id and ships the 128 bytes to the others by any means (bench.py: torch.distributed).
 * ---------------------------------------------------------------------------------------------- */
int trpo_nccl_unique_id(char id_out[128]);
int trpo_ctx_init_comm(trpo_ctx *ctx, const char id[128], int rank, int world_size);
/* Peer-memory (NVLink / NVSwitch) all-reduce for the FVP sums, replacing the ncclAllReduce: the send side is fused into
 * the kernel that reduces the per-CTA partial rows (each rank pushes its sum into every peer's memory), the receive side
 * into the CG-update / FVP-finalise kernel (waits on local flags, sums the ranks in fixed order). One PROCESS per GPU:
 * every rank exports its communication buffer as a 64-byte CUDA IPC handle (after trpo_ctx_init_comm), the application
 * all-gathers the handles (rank order) and attaches them. A barrier across ranks must separate attach from first use. */
int trpo_ctx_p2p_export(trpo_ctx *ctx, char handle_out[64]);
int trpo_ctx_p2p_attach(trpo_ctx *ctx, const char *handles /* world_size x 64 bytes */);
enum { TRPO_COMM_NCCL = 0, TRPO_COMM_P2P = 1 };
int trpo_ctx_set_comm_mode(trpo_ctx *ctx, int mode);       /* P2P becomes the default once attached */
/* Non-zero if a peer / staging wait has timed out and no call has reported it yet (synchronises). The synchronous host-buffer
 * calls (trpo_ctx_fvp / _cg / _update / trpo_vf_evaluate) check the same flags themselves: they fail with -1, reset the flag, and
 * the device results of that call are poisoned (NaN) / the solve stopped. Callers of the asynchronous *_device calls poll this. */
int trpo_ctx_comm_error(trpo_ctx *ctx);

/* Roofline denominators measured on the device (< 0.1 s each; device < 0: the current one). MEASURED_PEAKS.json has no FP64 entry:
 * the FP64 probe runs mma.sync.m8n8k4.f64 (DMMA, the pipe DFMA shares) with 16 independent accumulators per warp; the TF32 probe
 * the legacy mma.sync.m16n8k8 path. TFLOP/s, or a negative value on failure. */
double trpo_probe_fp64_peak_tflops(int device);
double trpo_probe_tf32_mma_sync_tflops(int device);

/* Total sample count over all ranks (the 1/N of TRPO_FVP.c:930). Computed by init_comm+set_batch via all-reduce. */
size_t trpo_ctx_global_samples(const trpo_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* TRPO_B200_H */
