// trpo_shim.cu -- the thin C-ABI shim between the C host code and the CUDA kernels (include/trpo_b200.h).
// Owns device memory, the stream, the device-resident CG state and the optional NCCL communicator.
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/trpo_b200.h"
#include "trpo_internal.cuh"

// --------------------------------------------------------------------------------------------------------------
// errors
static thread_local char g_err[512] = "";
static int fail(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return -1;
}
extern "C" const char *trpo_last_error(void) { return g_err; }

#define KTIME_MAX 4096
#define STAGE_CHUNKS 8
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// --------------------------------------------------------------------------------------------------------------
// NCCL through dlopen: the library has no link-time NCCL dependency; in a torch process this resolves to the
// libnccl.so.2 torch already loaded, in a plain C program to the system one.
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSum_ = 0 };
enum { ncclInt64_ = 4, ncclUint64_ = 5, ncclFloat64_ = 8 };
struct NcclApi {
    void *handle;
    int (*GetUniqueId)(ncclUniqueId *);
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*CommDestroy)(ncclComm_t);
    const char *(*GetErrorString)(int);
};
static NcclApi g_nccl = {};
static int nccl_load() {
    if (g_nccl.handle) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail("cannot dlopen libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail("libnccl is missing required symbols");
    g_nccl.handle = h;
    return 0;
}
#define NC(x) do { int r_ = (x); if (r_ != 0) return fail("%s failed: %s", #x, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error"); } while (0)

// --------------------------------------------------------------------------------------------------------------
struct trpo_ctx {
    NetDesc net;
    int device;
    cudaStream_t stream;
    bool own_stream;
    int path_req, path_used;
    long long launches;

    double *d_theta, *d_inv_var, *d_std;
    double *d_inv_std_model;   // exp(-LogStd) of the current parameters (policy-gradient seed, TRPO_Update.c:298-300)
    // batch
    size_t n_local, n_total;
    double *d_obs, *d_mean, *d_action, *d_adv;
    bool own_batch;
    size_t cap_obs, cap_mean, cap_adv;
    size_t n_full;             // rows for which Mean / Action / Advantage are staged (0: the batch carries observations only)
    size_t chunk_override;     // trpo_ctx_set_chunk: samples per GEMM-chain pass (0 = automatic)
    // rollout staging (rows f-3/f-4): per-step rewards of the staged batch and its episode length
    double *d_reward;
    size_t cap_reward, ep_len;
    int *d_draws;              // raw rand() draws of the rollout producer
    size_t cap_draws;
    double *h_logstd;          // LogStd block of the last trpo_ctx_set_model (host copy, A entries)
    // work vectors (P each)
    double *d_in, *d_out, *d_zsum, *d_x, *d_r, *d_p, *d_z, *d_b, *d_xnew;
    double *d_scal;            // small scalar scratch (16 doubles)
    double *d_blockpart;       // block partial sums (1024 doubles)
    double *d_mean_new;        // [n_local x A] line-search forward output
    size_t cap_mean_new;
    CgState *d_state;
    CgState *h_state;          // pinned
    double *d_trace, *h_trace; // per-iteration CG trace: rdotr[0..trace_cap), xnorm[0..trace_cap) (h_trace pinned)
    int trace_cap;
    double *d_dots;            // persistent solve kernel: per-CTA partial dot products [4][160]
    unsigned int *d_gbar;      // ... and its grid-barrier state (arrivals, generation)
    unsigned long long *d_timeline;   // optional per-iteration phase stamps of the solve kernel (trpo_ctx_solve_timeline)
    size_t timeline_iters;
    int *h_flags;              // pinned: [0] peer-memory wait error, [1] streamed-staging wait error (read back after a sync)
    double *h_scal;            // pinned
    double *h_vec;             // pinned, P doubles: host vectors go through here while a batch copy occupies the copy engine
    // gemm-chain scratch
    ChainScratch sc;
    double *sc_base;
    size_t sc_bytes;
    // fused path partials
    double *d_fused_partial;
    // optional FP32 mode (FVP inside CG): FP32 copies of the model, direction and observations + FP32 scratch
    int precision;
    float *f_theta, *f_v, *f_inv_var, *f_obs;
    float *f_obs_lo;           // TF32 remainders of f_obs for the tcgen05 forward layer (NULL: layer 0 stays on the legacy kernel)
    size_t cap_fobs;
    ChainScratchF32 scf;
    float *scf_base;
    size_t scf_floats;
    // optional event timing of the FVP-sum kernel(s)
    bool ktime_on;
    int ktime_n;
    cudaEvent_t *ktime_ev;     // 2 * KTIME_MAX events
    // comm
    ncclComm_t comm;
    bool borrowed_comm;        // comm belongs to another context (the value-function network shares the policy's)
    int rank, world;
    // streamed staging of the observation matrix (pinned host source, fused path): the first FVP after set_batch
    // overlaps the host-to-device copy, the kernel polling d_ready for the chunks it needs
    cudaStream_t copy_stream;
    cudaEvent_t ev_compute, ev_copy;
    int *d_ready;              // [0]: chunks landed, [1]: error flag
    int *h_ready_vals;         // pinned 0..STAGE_CHUNKS
    bool copy_inflight;        // an asynchronous batch copy has been issued and not yet joined into c->stream
    bool stream_first_fvp;     // the next fused FVP may start before the copy has finished
    size_t stage_chunk;        // samples per staged chunk
    // GEMM-chain path: the pinned observation matrix is copied in pieces with an event each; the first FVP's chunk loop waits
    // per piece, so chunk k computes while chunk k + 1 is still crossing PCIe (Humanoid-size batch: 3 GB, 55 ms)
    cudaEvent_t ev_piece[64];
    int n_pieces;
    size_t piece_rows;
    bool pieces_pending;
    // CUDA graph of a whole CG solve (single GPU, fused path): captured on the second identical call, replayed afterwards
    struct CgKey { const double *db; double *dres; size_t iters; double th, damping; const double *obs; size_t n; cudaStream_t st; int path; } cg_key;
    int cg_key_seen;
    cudaGraphExec_t cg_exec;
    long long cg_graph_launches;
    // peer-memory all-reduce
    P2PComm p2p;               // world == 0 until attached
    bool p2p_on;
    bool solve_kernel_used;    // the last CG ran as the single persistent kernel
    char *p2p_buf;             // own communication buffer (header + slots)
    void *p2p_peer[TRPO_MAX_RANKS];
    trpo_info info;
};

extern "C" size_t trpo_num_params(const size_t *LayerSize, size_t NumLayers) {
    size_t n = 0;
    for (size_t i = 0; i + 1 < NumLayers; ++i) n += LayerSize[i] * LayerSize[i + 1] + LayerSize[i + 1];
    return n + LayerSize[NumLayers - 1];
}

static int ensure_chain_scratch(trpo_ctx *c) {
    // chunk: many whole waves of CTAs per kernel so the fixed per-launch cost (first-tile latency, epilogue, tail, ~10 us
    // per kernel measured) is amortised; up to 4 GB of scratch (of 180 GB). Wide layers are compute bound even from HBM (>= 100 flop/B), so the chunk does not
    // have to stay L2 resident.
    // the R{y} / gradient ping-pong buffers only ever hold layers 1..K (Ry0 = 0, no gradient w.r.t. the observations)
    size_t maxL = 1, sumL = 0;
    for (int i = 1; i <= c->net.K; ++i) { sumL += c->net.L[i]; if ((size_t)c->net.L[i] > maxL) maxL = c->net.L[i]; }
    size_t per_sample = 8 * (sumL + 4 * maxL);
    // rows per chunk = 128 * m with m * (n-tiles of the widest layer) a multiple of the 148 SMs: whole waves of CTAs
    size_t maxH = 1;
    for (int i = 1; i <= c->net.K; ++i) if ((size_t)c->net.L[i] > maxH) maxH = c->net.L[i];
    const size_t tiles_n = (maxH + 63) / 64;
    size_t gcd = 148, b = tiles_n;
    while (b) { size_t r = gcd % b; gcd = b; b = r; }
    const size_t unit = 128 * (148 / gcd);
    size_t chunk = (((size_t)4 << 30) / per_sample / unit) * unit;
    if (chunk < unit) chunk = unit;
    if (chunk > 64 * unit) chunk = 64 * unit;
    if (chunk > 1048576) chunk = (1048576 / unit) * unit;
    {   // the GEMM kernels index rows * width with 32-bit ints
        size_t widest = c->net.L[0] + 1;
        for (int i = 1; i <= c->net.K; ++i) if ((size_t)c->net.L[i] + 1 > widest) widest = c->net.L[i] + 1;
        while (chunk > unit && chunk * widest >= ((size_t)1 << 31)) chunk -= unit;
    }
    if (c->chunk_override) chunk = ((c->chunk_override + 127) / 128) * 128;      // trpo_ctx_set_chunk (tests, memory-tight callers)
    if (c->n_local && chunk > ((c->n_local + 127) / 128) * 128) chunk = ((c->n_local + 127) / 128) * 128;
    int nslices = 148;
    if ((size_t)nslices * 16 > chunk) nslices = (int)(chunk / 16);
    if (nslices < 1) nslices = 1;
    size_t bytes = chain_scratch_bytes(c->net, (int)chunk, nslices);
    if (c->sc_base && bytes <= c->sc_bytes && c->sc.chunk == (int)chunk && c->sc.nslices == nslices) return 0;
    if (c->sc_base) cudaFree(c->sc_base);
    c->sc_base = nullptr;
    CU(cudaMalloc(&c->sc_base, bytes));
    CU(cudaMemsetAsync(c->sc_base, 0, bytes, c->stream));
    c->sc_bytes = bytes;
    double *p = c->sc_base;
    for (int i = 1; i <= c->net.K; ++i) { c->sc.Y[i] = p; p += chunk * c->net.L[i]; }
    for (int j = 0; j < 2; ++j) { c->sc.RY[j] = p; p += chunk * maxL; }
    for (int j = 0; j < 2; ++j) { c->sc.G[j] = p; p += chunk * maxL; }
    c->sc.partial = p;
    p += (size_t)nslices * c->net.P;
    if ((p - c->sc_base) & 1) ++p;                       // 16-byte alignment for the tensor maps
    c->sc.wperm = p;
    c->sc.vperm = p + chain_tma_perm_offset(c->net, c->net.K);
    c->sc.tailw = c->sc.vperm + chain_tma_perm_offset(c->net, c->net.K);
    c->sc.chunk = (int)chunk;
    c->sc.nslices = nslices;
    return 0;
}

extern "C" trpo_ctx *trpo_ctx_create(const size_t *LayerSize, const char *AcFunc, size_t NumLayers, int device, int precision) {
    if (!LayerSize || !AcFunc || NumLayers < 2 || NumLayers > TRPO_MAX_LAYERS) { fail("bad network description"); return nullptr; }
    if (precision != TRPO_PRECISION_FP64 && precision != TRPO_PRECISION_FP32) { fail("unknown precision %d", precision); return nullptr; }
    for (size_t i = 1; i < NumLayers; ++i) {
        const char a = AcFunc[i];
        if (a != 'l' && a != 't' && a != 'o' && a != 's') {
            // the reference prints this and carries on with stale values (TRPO_FVP.c:831-833); we refuse instead
            fprintf(stderr, "[ERROR] AC Function for Layer[%zu] is %c. Unsupported.\n", i, a);
            fail("unsupported activation '%c' for layer %zu", a, i);
            return nullptr;
        }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fail("no CUDA device available: this library has no CPU fallback");
        return nullptr;
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (cudaSetDevice(device) != cudaSuccess) { fail("cudaSetDevice(%d) failed", device); return nullptr; }
    trpo_ctx *c = (trpo_ctx *)calloc(1, sizeof(trpo_ctx));
    c->device = device;
    c->net.K = (int)NumLayers - 1;
    int pos = 0;
    for (size_t i = 0; i < NumLayers; ++i) { c->net.L[i] = (int)LayerSize[i]; c->net.ac[i] = i ? AcFunc[i] : 'l'; }
    for (int i = 0; i < c->net.K; ++i) { c->net.w_off[i] = pos; pos += c->net.L[i] * c->net.L[i + 1] + c->net.L[i + 1]; }
    c->net.logstd_off = pos;
    c->net.P = pos + c->net.L[c->net.K];
    c->world = 1;
    c->precision = precision;
    c->h_logstd = (double *)calloc((size_t)c->net.L[c->net.K], sizeof(double));
    if (precision == TRPO_PRECISION_FP32) c->path_req = TRPO_PATH_GEMM_CHAIN;      // the FP32 mode is a GEMM-chain mode
    const size_t P = c->net.P, A = c->net.L[c->net.K];
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    c->own_stream = true;
    double **vecs[] = {&c->d_theta, &c->d_in, &c->d_out, &c->d_zsum, &c->d_x, &c->d_r, &c->d_p, &c->d_z, &c->d_b, &c->d_xnew};
    for (auto v : vecs) ok = ok && cudaMalloc(v, P * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_inv_var, A * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_std, A * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_inv_std_model, A * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_scal, 16 * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_blockpart, 1024 * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_state, sizeof(CgState)) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_state, sizeof(CgState)) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_scal, 16 * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_vec, P * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_flags, 2 * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_dots, 4 * 160 * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_gbar, 2 * sizeof(unsigned int)) == cudaSuccess;
    ok = ok && cudaMemset(c->d_gbar, 0, 2 * sizeof(unsigned int)) == cudaSuccess;
    c->trace_cap = 64;
    ok = ok && cudaMalloc(&c->d_trace, 2 * (size_t)c->trace_cap * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_trace, 2 * (size_t)c->trace_cap * sizeof(double)) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_compute, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 64 && ok; ++i) ok = cudaEventCreateWithFlags(&c->ev_piece[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&c->d_ready, 2 * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMemset(c->d_ready, 0, 2 * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_ready_vals, (STAGE_CHUNKS + 1) * sizeof(int)) == cudaSuccess;
    if (ok) for (int i = 0; i <= STAGE_CHUNKS; ++i) c->h_ready_vals[i] = i;
    if (ok && fused_eligible(c->net))
        ok = cudaMalloc(&c->d_fused_partial, (size_t)fused_partial_rows() * P * sizeof(double)) == cudaSuccess;
    if (ok && precision == TRPO_PRECISION_FP32) {
        ok = cudaMalloc(&c->f_theta, P * sizeof(float)) == cudaSuccess && cudaMalloc(&c->f_v, P * sizeof(float)) == cudaSuccess &&
             cudaMalloc(&c->f_inv_var, A * sizeof(float)) == cudaSuccess;
    }
    if (ok) ok = cudaMemset(c->d_state, 0, sizeof(CgState)) == cudaSuccess;
    if (!ok) { fail("device allocation failed: %s", cudaGetErrorString(cudaGetLastError())); trpo_ctx_destroy(c); return nullptr; }
    return c;
}

static void free_batch(trpo_ctx *c) {
    if (c->own_batch) { cudaFree(c->d_obs); cudaFree(c->d_mean); cudaFree(c->d_action); cudaFree(c->d_adv); }
    c->d_obs = c->d_mean = c->d_action = c->d_adv = nullptr;
    c->cap_obs = c->cap_mean = c->cap_adv = 0;
    c->n_full = 0;
    c->own_batch = false;
}

extern "C" void trpo_ctx_destroy(trpo_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && !c->borrowed_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (int r = 0; r < TRPO_MAX_RANKS; ++r) if (c->p2p_peer[r]) cudaIpcCloseMemHandle(c->p2p_peer[r]);
    if (c->p2p_buf) cudaFree(c->p2p_buf);
    free_batch(c);
    double *vecs[] = {c->d_theta, c->d_in, c->d_out, c->d_zsum, c->d_x, c->d_r, c->d_p, c->d_z, c->d_b, c->d_xnew,
                      c->d_inv_var, c->d_std, c->d_inv_std_model, c->d_scal, c->d_blockpart, c->d_mean_new, c->sc_base, c->d_fused_partial,
                      c->d_reward};
    for (double *v : vecs) if (v) cudaFree(v);
    float *fvecs[] = {c->f_theta, c->f_v, c->f_inv_var, c->f_obs, c->f_obs_lo, c->scf_base};
    for (float *v : fvecs) if (v) cudaFree(v);
    if (c->d_draws) cudaFree(c->d_draws);
    free(c->h_logstd);
    if (c->d_state) cudaFree(c->d_state);
    if (c->h_state) cudaFreeHost(c->h_state);
    if (c->h_scal) cudaFreeHost(c->h_scal);
    if (c->h_vec) cudaFreeHost(c->h_vec);
    if (c->h_flags) cudaFreeHost(c->h_flags);
    if (c->d_dots) cudaFree(c->d_dots);
    if (c->d_gbar) cudaFree(c->d_gbar);
    if (c->d_timeline) cudaFree(c->d_timeline);
    if (c->d_trace) cudaFree(c->d_trace);
    if (c->h_trace) cudaFreeHost(c->h_trace);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->ev_compute) cudaEventDestroy(c->ev_compute);
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    for (int i = 0; i < 64; ++i) if (c->ev_piece[i]) cudaEventDestroy(c->ev_piece[i]);
    if (c->d_ready) cudaFree(c->d_ready);
    if (c->h_ready_vals) cudaFreeHost(c->h_ready_vals);
    if (c->cg_exec) cudaGraphExecDestroy(c->cg_exec);
    if (c->ktime_ev) { for (int i = 0; i < 2 * KTIME_MAX; ++i) if (c->ktime_ev[i]) cudaEventDestroy(c->ktime_ev[i]); free(c->ktime_ev); }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    free(c);
}

extern "C" size_t trpo_ctx_num_params(const trpo_ctx *c) { return c ? (size_t)c->net.P : 0; }
extern "C" int trpo_ctx_set_stream(trpo_ctx *c, void *s) {
    if (!c) return fail("null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (c->own_stream) cudaStreamDestroy(c->stream);
    if (s) { c->stream = (cudaStream_t)s; c->own_stream = false; }
    else { CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    return 0;
}
extern "C" void *trpo_ctx_get_stream(const trpo_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int trpo_ctx_set_path(trpo_ctx *c, int path) {
    if (!c) return fail("null context");
    if (c->precision == TRPO_PRECISION_FP32 && path == TRPO_PATH_FUSED) return fail("the FP32 mode has no fused kernel");
    if (c->precision == TRPO_PRECISION_FP32) path = TRPO_PATH_GEMM_CHAIN;
    if (path == TRPO_PATH_FUSED && !fused_eligible(c->net)) return fail("network shape is not eligible for the fused kernel");
    c->path_req = path;
    return 0;
}
extern "C" int trpo_ctx_get_path(const trpo_ctx *c) { return c ? c->path_used : 0; }
extern "C" int trpo_ctx_set_chunk(trpo_ctx *c, size_t chunk_samples) {
    if (!c) return fail("null context");
    c->chunk_override = chunk_samples;
    return 0;
}
extern "C" size_t trpo_ctx_get_chunk(const trpo_ctx *c) {
    if (!c) return 0;
    return c->precision == TRPO_PRECISION_FP32 ? (size_t)c->scf.chunk : (size_t)c->sc.chunk;
}
extern "C" int trpo_ctx_sync(trpo_ctx *c) {
    if (!c) return fail("null context");
    CU(cudaSetDevice(c->device));
    if (c->copy_inflight) CU(cudaEventSynchronize(c->ev_copy));      // a streamed batch copy counts as work of this context
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" long long trpo_ctx_launch_count(const trpo_ctx *c) { return c ? c->launches : 0; }

extern "C" int trpo_ctx_set_model(trpo_ctx *c, const double *theta) {
    if (!c || !theta) return fail("null argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(c->d_theta, theta, c->net.P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    {
        const int A = c->net.L[c->net.K];
        double is[TRPO_MAX_LAYERS * 64];
        if (A > (int)(sizeof(is) / sizeof(is[0]))) return fail("action dimension too large");
        for (int j = 0; j < A; ++j) { is[j] = 1.0 / exp(theta[c->net.logstd_off + j]); c->h_logstd[j] = theta[c->net.logstd_off + j]; }
        CU(cudaMemcpyAsync(c->d_inv_std_model, is, A * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));      // `is` is a stack buffer
    }
    if (c->precision == TRPO_PRECISION_FP32) chain_f32_convert(c->d_theta, c->f_theta, c->net.P, c->stream, &c->launches);
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

static int set_std(trpo_ctx *c, const double *Std) {
    const int A = c->net.L[c->net.K];
    double iv[TRPO_MAX_LAYERS * 64];
    if (A > (int)(sizeof(iv) / sizeof(iv[0]))) return fail("action dimension too large");
    for (int j = 0; j < A; ++j) iv[j] = 1.0 / (Std[j] * Std[j]);
    CU(cudaMemcpyAsync(c->d_std, Std, A * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_inv_var, iv, A * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (c->precision == TRPO_PRECISION_FP32) {
        chain_f32_convert(c->d_inv_var, c->f_inv_var, A, c->stream, &c->launches);
        // FP32 copy of the observations (they are already on the stream: set_std runs after the batch copy / adoption)
        const size_t n = c->n_local * (size_t)c->net.L[0];
        const bool want_lo = c->net.K >= 2 && tc_fwd_eligible(c->net.L[0], c->net.L[1]);
        if (n > c->cap_fobs || (want_lo && !c->f_obs_lo)) {
            if (c->f_obs) cudaFree(c->f_obs);
            if (c->f_obs_lo) cudaFree(c->f_obs_lo);
            c->f_obs = c->f_obs_lo = nullptr;
            c->cap_fobs = 0;
            CU(cudaMalloc(&c->f_obs, n * sizeof(float)));
            if (want_lo) CU(cudaMalloc(&c->f_obs_lo, n * sizeof(float)));
            c->cap_fobs = n;
        }
        chain_f32_convert(c->d_obs, c->f_obs, n, c->stream, &c->launches);
        if (c->f_obs_lo) tc_lo_split(c->f_obs, c->f_obs_lo, n, c->stream, &c->launches);
    }
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

static int update_global_samples(trpo_ctx *c) {
    c->n_total = c->n_local;
    if (c->comm) {
        unsigned long long *d = (unsigned long long *)c->d_scal;
        unsigned long long v = c->n_local;
        CU(cudaMemcpyAsync(d, &v, sizeof(v), cudaMemcpyHostToDevice, c->stream));
        NC(g_nccl.AllReduce(d, d, 1, ncclUint64_, ncclSum_, c->comm, c->stream));
        CU(cudaMemcpyAsync(&v, d, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->n_total = (size_t)v;
    }
    return 0;
}

extern "C" int trpo_ctx_set_batch(trpo_ctx *c, size_t N, const double *Observ, const double *Std,
                                  const double *Mean, const double *Action, const double *Advantage) {
    if (!c || !Observ || !Std || N == 0) return fail("bad batch arguments");
    CU(cudaSetDevice(c->device));
    const size_t O = c->net.L[0], A = c->net.L[c->net.K];
    if (!c->own_batch) free_batch(c);
    c->own_batch = true;
    if (N * O > c->cap_obs) { cudaFree(c->d_obs); c->d_obs = nullptr; CU(cudaMalloc(&c->d_obs, N * O * sizeof(double))); c->cap_obs = N * O; }
    // join any copy still in flight, then decide how to stage the observations
    if (c->copy_inflight) { CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0)); c->copy_inflight = false; }
    c->stream_first_fvp = false;
    c->pieces_pending = false;
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, Observ) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned && N >= 65536 && fused_eligible(c->net) && c->path_req != TRPO_PATH_GEMM_CHAIN) {
        // pinned source: DMA in STAGE_CHUNKS pieces on the copy stream, bumping the device-side chunk counter after each
        // piece; the compute stream is not made to wait (the fused kernel polls the counter, see wait_samples)
        c->stage_chunk = ((N + STAGE_CHUNKS - 1) / STAGE_CHUNKS + 63) / 64 * 64;
        // The chunk counter is reset on the COMPUTE stream: the copy stream waits on ev_compute, so the reset is ordered
        // before every chunk bump, and the polling kernel is launched on the compute stream after it -- a reset issued on
        // the copy stream had no ordering against that kernel (it could still see the previous batch's final count).
        CU(cudaMemsetAsync(c->d_ready, 0, sizeof(int), c->stream));
        CU(cudaEventRecord(c->ev_compute, c->stream));
        CU(cudaStreamWaitEvent(c->copy_stream, c->ev_compute, 0));          // do not overwrite rows still being read
        int landed = 0;
        for (size_t s0 = 0; s0 < N; s0 += c->stage_chunk) {
            const size_t n = (N - s0 < c->stage_chunk) ? N - s0 : c->stage_chunk;
            CU(cudaMemcpyAsync(c->d_obs + s0 * O, Observ + s0 * O, n * O * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
            ++landed;
            CU(cudaMemcpyAsync(c->d_ready, &c->h_ready_vals[landed], sizeof(int), cudaMemcpyHostToDevice, c->copy_stream));
        }
        CU(cudaEventRecord(c->ev_copy, c->copy_stream));
        c->copy_inflight = true;
        c->stream_first_fvp = true;
    } else if (pinned && N >= 65536 && c->precision == TRPO_PRECISION_FP64 && N * O * sizeof(double) >= ((size_t)256 << 20)) {
        // GEMM-chain path, large pinned batch: pieces of about 256 MB on the copy stream, one event each
        size_t pr = ((((size_t)256 << 20) / (O * sizeof(double))) + 127) / 128 * 128;
        if ((N + pr - 1) / pr > 64) pr = ((N + 63) / 64 + 127) / 128 * 128;
        c->piece_rows = pr;
        c->n_pieces = (int)((N + pr - 1) / pr);
        CU(cudaEventRecord(c->ev_compute, c->stream));
        CU(cudaStreamWaitEvent(c->copy_stream, c->ev_compute, 0));          // do not overwrite rows still being read
        for (int i = 0; i < c->n_pieces; ++i) {
            const size_t s0 = (size_t)i * pr, n = (N - s0 < pr) ? N - s0 : pr;
            CU(cudaMemcpyAsync(c->d_obs + s0 * O, Observ + s0 * O, n * O * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
            CU(cudaEventRecord(c->ev_piece[i], c->copy_stream));
        }
        CU(cudaEventRecord(c->ev_copy, c->copy_stream));
        c->copy_inflight = true;
        c->pieces_pending = true;
    } else {
        CU(cudaMemcpyAsync(c->d_obs, Observ, N * O * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    if (Mean && Action && Advantage) {
        if (N * A > c->cap_mean) {
            cudaFree(c->d_mean); cudaFree(c->d_action); c->d_mean = c->d_action = nullptr;
            CU(cudaMalloc(&c->d_mean, N * A * sizeof(double)));
            CU(cudaMalloc(&c->d_action, N * A * sizeof(double)));
            c->cap_mean = N * A;
        }
        if (N > c->cap_adv) { cudaFree(c->d_adv); c->d_adv = nullptr; CU(cudaMalloc(&c->d_adv, N * sizeof(double))); c->cap_adv = N; }
        CU(cudaMemcpyAsync(c->d_mean, Mean, N * A * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_action, Action, N * A * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_adv, Advantage, N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        c->n_full = N;
    } else {
        c->n_full = 0;         // stale Mean / Action / Advantage of an earlier batch must not be paired with these rows
    }
    c->n_local = N;
    if (set_std(c, Std)) return -1;
    return update_global_samples(c);
}

extern "C" int trpo_ctx_set_batch_device(trpo_ctx *c, size_t N, const double *dObserv, const double *Std_host,
                                         const double *dMean, const double *dAction, const double *dAdvantage) {
    if (!c || !dObserv || !Std_host || N == 0) return fail("bad batch arguments");
    CU(cudaSetDevice(c->device));
    if (c->copy_inflight) { CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0)); c->copy_inflight = false; }
    c->stream_first_fvp = false;
    free_batch(c);
    c->own_batch = false;
    c->d_obs = (double *)dObserv; c->d_mean = (double *)dMean; c->d_action = (double *)dAction; c->d_adv = (double *)dAdvantage;
    c->n_local = N;
    c->n_full = (dMean && dAction && dAdvantage) ? N : 0;
    if (set_std(c, Std_host)) return -1;
    return update_global_samples(c);
}


// Stage a binary batch file (host/trpo_batch_file.c) section by section through two pinned 32 MB buffers: the read of
// one piece overlaps the host-to-device copy of the previous one.
extern "C" int trpo_ctx_set_batch_file(trpo_ctx *c, const char *path, size_t N) {
    if (!c || !path) return fail("null argument");
    trpo_batch_file_header h;
    if (trpo_batch_file_probe(path, &h)) {
        fprintf(stderr, "[ERROR] Cannot open Data File [%s]. \n", path);
        return fail("%s is not a binary batch file", path);
    }
    const size_t O = c->net.L[0], A = c->net.L[c->net.K], NF = (size_t)h.NumSamples;
    if (h.ObservSpaceDim != O || h.ActionSpaceDim != A) return fail("batch file is %llux%llu, the network needs %zux%zu", h.ObservSpaceDim, h.ActionSpaceDim, O, A);
    if (N == 0) N = NF;
    if (N > NF) return fail("batch file holds %zu samples, %zu requested", NF, N);
    CU(cudaSetDevice(c->device));
    if (c->copy_inflight) { CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0)); c->copy_inflight = false; }
    c->stream_first_fvp = false;
    if (!c->own_batch) free_batch(c);
    c->own_batch = true;
    const bool full = (h.flags & 1u) != 0;
    if (N * O > c->cap_obs) { cudaFree(c->d_obs); c->d_obs = nullptr; CU(cudaMalloc(&c->d_obs, N * O * sizeof(double))); c->cap_obs = N * O; }
    if (full) {
        if (N * A > c->cap_mean) {
            cudaFree(c->d_mean); cudaFree(c->d_action); c->d_mean = c->d_action = nullptr;
            CU(cudaMalloc(&c->d_mean, N * A * sizeof(double)));
            CU(cudaMalloc(&c->d_action, N * A * sizeof(double)));
            c->cap_mean = N * A;
        }
        if (N > c->cap_adv) { cudaFree(c->d_adv); c->d_adv = nullptr; CU(cudaMalloc(&c->d_adv, N * sizeof(double))); c->cap_adv = N; }
    }
    FILE *f = fopen(path, "rb");
    if (!f) return fail("cannot open %s", path);
    const size_t PIECE = (32u << 20) / sizeof(double);
    double *pin[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool used[2] = {false, false};
    int rc = 0, which = 0;
    for (int i = 0; i < 2 && !rc; ++i)
        if (cudaMallocHost(&pin[i], PIECE * sizeof(double)) != cudaSuccess || cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess)
            rc = fail("pinned staging allocation failed");
    double std_host[TRPO_MAX_LAYERS * 64];
    if (!rc && A > sizeof(std_host) / sizeof(std_host[0])) rc = fail("action dimension too large");
    if (!rc && (fseek(f, (long)sizeof(h), SEEK_SET) != 0 || fread(std_host, sizeof(double), A, f) != A)) rc = fail("short batch file");
    struct Section { double *dst; size_t count, file_count; };
    const Section secs[4] = {{c->d_obs, N * O, NF * O}, {c->d_mean, N * A, NF * A}, {c->d_action, N * A, NF * A}, {c->d_adv, N, NF}};
    long off = (long)(sizeof(h) + A * sizeof(double));
    for (int sidx = 0; sidx < (full ? 4 : 1) && !rc; ++sidx) {
        if (fseek(f, off, SEEK_SET) != 0) { rc = fail("short batch file"); break; }
        for (size_t done = 0; done < secs[sidx].count && !rc; done += PIECE) {
            const size_t n = secs[sidx].count - done < PIECE ? secs[sidx].count - done : PIECE;
            if (used[which] && cudaEventSynchronize(ev[which]) != cudaSuccess) { rc = fail("staging copy failed"); break; }
            if (fread(pin[which], sizeof(double), n, f) != n) { rc = fail("short batch file"); break; }
            if (cudaMemcpyAsync(secs[sidx].dst + done, pin[which], n * sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
                cudaEventRecord(ev[which], c->stream) != cudaSuccess) { rc = fail("staging copy failed"); break; }
            used[which] = true;
            which ^= 1;
        }
        off += (long)(secs[sidx].file_count * sizeof(double));
    }
    fclose(f);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 2; ++i) { if (pin[i]) cudaFreeHost(pin[i]); if (ev[i]) cudaEventDestroy(ev[i]); }
    if (rc) return rc;
    c->n_local = N;
    c->n_full = full ? N : 0;
    if (set_std(c, std_host)) return -1;
    return update_global_samples(c);
}

// --------------------------------------------------------------------------------------------------------------
static inline const P2PComm *active_p2p(const trpo_ctx *c) {
    return (c->p2p_on && c->p2p.world > 1 && c->precision == TRPO_PRECISION_FP64) ? &c->p2p : nullptr;
}
static inline bool active_p2p_ctx(const trpo_ctx *c) { return active_p2p(c) != nullptr; }

extern "C" int trpo_ctx_kernel_timing(trpo_ctx *c, int enable) {
    if (!c) return fail("null context");
    CU(cudaSetDevice(c->device));
    if (enable && !c->ktime_ev) {
        c->ktime_ev = (cudaEvent_t *)calloc(2 * KTIME_MAX, sizeof(cudaEvent_t));
        for (int i = 0; i < 2 * KTIME_MAX; ++i) CU(cudaEventCreate(&c->ktime_ev[i]));
    }
    c->ktime_on = enable != 0;
    c->ktime_n = 0;
    return 0;
}
extern "C" double trpo_ctx_kernel_time_ms(trpo_ctx *c, int *launches) {
    if (!c || !c->ktime_ev) { if (launches) *launches = 0; return 0.0; }
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    double total = 0.0;
    for (int i = 0; i < c->ktime_n; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ktime_ev[2 * i], c->ktime_ev[2 * i + 1]) == cudaSuccess) total += ms;
    }
    if (launches) *launches = c->ktime_n;
    c->ktime_n = 0;
    return total;
}

// un-normalised FVP sum of the local shard into d_zsum, then the cross-GPU sum
static int ensure_f32_scratch(trpo_ctx *c) {
    size_t maxL = 1, sumL = 0, maxH = 1;           // same sizing rules as ensure_chain_scratch
    for (int i = 1; i <= c->net.K; ++i) {
        sumL += c->net.L[i];
        if ((size_t)c->net.L[i] > maxL) maxL = c->net.L[i];
        if ((size_t)c->net.L[i] > maxH) maxH = c->net.L[i];
    }
    const size_t per_sample = 4 * (sumL + 4 * maxL);
    const size_t tiles_n = (maxH + 63) / 64;
    size_t gcd = 148, b = tiles_n;
    while (b) { size_t r = gcd % b; gcd = b; b = r; }
    const size_t unit = 128 * (148 / gcd);
    size_t chunk = (((size_t)2 << 30) / per_sample / unit) * unit;
    if (chunk < unit) chunk = unit;
    if (chunk > 64 * unit) chunk = 64 * unit;
    if (chunk > 1048576) chunk = (1048576 / unit) * unit;
    {
        size_t widest = c->net.L[0] + 1;
        for (int i = 1; i <= c->net.K; ++i) if ((size_t)c->net.L[i] + 1 > widest) widest = c->net.L[i] + 1;
        while (chunk > unit && chunk * widest >= ((size_t)1 << 31)) chunk -= unit;
    }
    if (c->chunk_override) chunk = ((c->chunk_override + 127) / 128) * 128;
    if (c->n_local && chunk > ((c->n_local + 127) / 128) * 128) chunk = ((c->n_local + 127) / 128) * 128;
    int nslices = 148;
    if ((size_t)nslices * 32 > chunk) nslices = (int)(chunk / 32);
    if (nslices < 1) nslices = 1;
    const size_t floats = chain_f32_scratch_floats(c->net, (int)chunk, nslices);
    if (c->scf_base && floats <= c->scf_floats && c->scf.chunk == (int)chunk && c->scf.nslices == nslices) return 0;
    if (c->scf_base) cudaFree(c->scf_base);
    c->scf_base = nullptr;
    CU(cudaMalloc(&c->scf_base, floats * sizeof(float)));
    CU(cudaMemsetAsync(c->scf_base, 0, floats * sizeof(float), c->stream));
    c->scf_floats = floats;
    float *p = c->scf_base;
    auto align64 = [](float *q) { return (float *)(((uintptr_t)q + 255) & ~(uintptr_t)255); };     // TMA bases: 16-byte aligned at least
    for (int i = 1; i <= c->net.K; ++i) { c->scf.Y[i] = p; p = align64(p + chunk * c->net.L[i]); }
    for (int j = 0; j < 2; ++j) { c->scf.RY[j] = p; p = align64(p + chunk * maxL); }
    for (int j = 0; j < 2; ++j) { c->scf.G[j] = p; p = align64(p + chunk * maxL); }
    for (int i = 1; i <= c->net.K; ++i) { c->scf.Ylo[i] = p; p = align64(p + chunk * c->net.L[i]); }
    for (int j = 0; j < 2; ++j) { c->scf.RYlo[j] = p; p = align64(p + chunk * maxL); }
    for (int i = 0; i + 1 < c->net.K; ++i) { c->scf.wt[i] = p; p = align64(p + 4 * (size_t)c->net.L[i] * c->net.L[i + 1]); }
    c->scf.partial = p;
    c->scf.chunk = (int)chunk;
    c->scf.nslices = nslices;
    return 0;
}

static int fvp_sum(trpo_ctx *c, const double *d_v, const int *d_done) {
    if (!c->d_obs || c->n_local == 0) return fail("no batch staged: call trpo_ctx_set_batch first");
    if (c->precision == TRPO_PRECISION_FP32) {
        c->path_used = TRPO_PATH_GEMM_CHAIN;
        if (c->copy_inflight) { CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0)); c->copy_inflight = false; }
        if (ensure_f32_scratch(c)) return -1;
        const bool timed32 = c->ktime_on && c->ktime_n < KTIME_MAX;
        if (timed32) cudaEventRecord(c->ktime_ev[2 * c->ktime_n], c->stream);
        chain_f32_convert(d_v, c->f_v, c->net.P, c->stream, &c->launches);
        c->scf.obs_lo = c->f_obs_lo;
        if (chain_f32_accumulate(c->net, c->scf, c->f_theta, c->f_v, c->f_inv_var, c->f_obs, c->n_local, c->d_zsum, d_done,
                                 nullptr, c->stream, &c->launches))
            return fail("FP32 gemm-chain FVP launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (timed32) { cudaEventRecord(c->ktime_ev[2 * c->ktime_n + 1], c->stream); ++c->ktime_n; }
        if (c->comm) NC(g_nccl.AllReduce(c->d_zsum, c->d_zsum, c->net.P, ncclFloat64_, ncclSum_, c->comm, c->stream));
        return 0;
    }
    int path = c->path_req;
    if (path == TRPO_PATH_AUTO) path = fused_eligible(c->net) ? TRPO_PATH_FUSED : TRPO_PATH_GEMM_CHAIN;
    c->path_used = path;
    const bool timed = c->ktime_on && c->ktime_n < KTIME_MAX;
    if (timed) cudaEventRecord(c->ktime_ev[2 * c->ktime_n], c->stream);
    if (path == TRPO_PATH_FUSED) {
        const bool streaming = c->stream_first_fvp && c->copy_inflight;
        if (fused_fvp_accumulate(c->net, c->d_theta, d_v, c->d_inv_var, c->d_obs, c->n_local, c->d_fused_partial,
                                 c->d_zsum, d_done, active_p2p(c), streaming ? c->d_ready : nullptr, c->stage_chunk,
                                 c->d_ready + 1, c->stream, &c->launches))
            return fail("fused FVP launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (c->copy_inflight) {      // everything enqueued after this FVP sees a fully resident batch
            CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
            c->copy_inflight = false;
        }
        c->stream_first_fvp = false;
    } else {
        const bool piecewise = c->copy_inflight && c->pieces_pending;
        if (c->copy_inflight && !piecewise) { CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0)); c->copy_inflight = false; }
        c->stream_first_fvp = false;
        if (ensure_chain_scratch(c)) return -1;
        c->sc.piece_events = piecewise ? c->ev_piece : nullptr;
        c->sc.n_pieces = piecewise ? c->n_pieces : 0;
        c->sc.piece_rows = piecewise ? c->piece_rows : 0;
        const int rc_chain = chain_accumulate(c->net, c->sc, CHAIN_FVP, c->d_theta, d_v, c->d_inv_var, c->d_obs, nullptr, nullptr, nullptr,
                                              c->n_local, c->d_zsum, d_done, active_p2p(c), c->stream, &c->launches);
        c->sc.piece_events = nullptr; c->sc.n_pieces = 0; c->sc.piece_rows = 0;
        if (piecewise) {             // everything enqueued after this FVP sees a fully resident batch
            CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
            c->copy_inflight = false;
            c->pieces_pending = false;
        }
        if (rc_chain)
            return fail("gemm-chain FVP launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (timed) { cudaEventRecord(c->ktime_ev[2 * c->ktime_n + 1], c->stream); ++c->ktime_n; }
    if (c->comm && !active_p2p(c)) NC(g_nccl.AllReduce(c->d_zsum, c->d_zsum, c->net.P, ncclFloat64_, ncclSum_, c->comm, c->stream));
    return 0;
}

extern "C" int trpo_ctx_fvp_device(trpo_ctx *c, const double *dInput, double *dResult, double damping) {
    if (!c || !dInput || !dResult) return fail("null argument");
    CU(cudaSetDevice(c->device));
    if (fvp_sum(c, dInput, nullptr)) return -1;
    launch_fvp_finalise(c->d_zsum, dInput, dResult, c->net.P, c->net.logstd_off, (double)c->n_total, damping, active_p2p(c), c->stream, &c->launches);
    CU(cudaGetLastError());
    return 0;
}

static int cg_enqueue(trpo_ctx *c, const double *db, double *dResult, size_t MaxIter, double ResidualTh, double damping) {
    launch_cg_init(db, c->d_x, c->d_r, c->d_p, c->net.P, ResidualTh, c->d_state, c->d_trace, c->trace_cap, c->stream, &c->launches);
    for (size_t it = 0; it < MaxIter; ++it) {
        if (fvp_sum(c, c->d_p, &c->d_state->done)) return -1;
        launch_cg_update(c->d_zsum, c->d_x, c->d_r, c->d_p, c->d_z, c->net.P, c->net.logstd_off, (double)c->n_total, damping,
                         ResidualTh, c->d_state, c->d_trace, c->trace_cap, active_p2p(c), c->stream, &c->launches);
    }
    CU(cudaMemcpyAsync(dResult, c->d_x, c->net.P * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

extern "C" int trpo_ctx_cg_device(trpo_ctx *c, const double *db, double *dResult, size_t MaxIter, double ResidualTh, double damping) {
    if (!c || !db || !dResult) return fail("null argument");
    if (MaxIter > (size_t)1 << 20) return fail("MaxIter %zu is not a sensible iteration count", MaxIter);
    CU(cudaSetDevice(c->device));
    if ((size_t)c->trace_cap < MaxIter + 2) {
        // the reference places no bound on MaxIter (TRPO_CG.c:45): grow the per-iteration trace buffers
        int cap = c->trace_cap;
        while ((size_t)cap < MaxIter + 2) cap *= 2;
        CU(cudaStreamSynchronize(c->stream));
        if (c->cg_exec) { cudaGraphExecDestroy(c->cg_exec); c->cg_exec = nullptr; c->cg_key_seen = 0; memset(&c->cg_key, 0, sizeof(c->cg_key)); }
        cudaFree(c->d_trace); cudaFreeHost(c->h_trace);
        c->d_trace = c->h_trace = nullptr;
        CU(cudaMalloc(&c->d_trace, 2 * (size_t)cap * sizeof(double)));
        CU(cudaMallocHost(&c->h_trace, 2 * (size_t)cap * sizeof(double)));
        c->trace_cap = cap;
    }
    // Shapes the fused kernels take, on one GPU or with the peer-memory exchange attached: the whole solve is ONE persistent
    // cooperative kernel (fvp_fused.cu, k_cg_solve). NCCL cannot be called from inside a kernel, so a communicator without
    // peer buffers keeps the per-iteration launches below.
    {
        const int path0 = (c->path_req == TRPO_PATH_AUTO) ? (fused_eligible(c->net) ? TRPO_PATH_FUSED : TRPO_PATH_GEMM_CHAIN) : c->path_req;
        const bool solo = !c->comm || active_p2p(c) != nullptr;
        if (path0 == TRPO_PATH_FUSED && c->precision == TRPO_PRECISION_FP64 && solo && c->d_obs && c->n_local) {
            const bool streaming = c->stream_first_fvp && c->copy_inflight;
            const bool timed = c->ktime_on && c->ktime_n < KTIME_MAX;
            if (timed) cudaEventRecord(c->ktime_ev[2 * c->ktime_n], c->stream);
            const int rc = fused_cg_solve(c->net, c->d_theta, c->d_inv_var, c->d_obs, c->n_local, (double)c->n_total, c->d_fused_partial,
                                          db, c->d_x, c->d_r, c->d_p, c->d_z, c->d_zsum, c->d_dots, c->d_gbar, c->d_state, c->d_trace,
                                          c->trace_cap, MaxIter, ResidualTh, damping, active_p2p(c), streaming ? c->d_ready : nullptr,
                                          c->stage_chunk, c->d_ready + 1,
                                          (c->d_timeline && MaxIter <= c->timeline_iters) ? c->d_timeline : nullptr, c->stream, &c->launches);
            if (rc < 0) return fail("fused CG solve launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            if (rc == 0) {
                if (timed) { cudaEventRecord(c->ktime_ev[2 * c->ktime_n + 1], c->stream); ++c->ktime_n; }
                if (c->copy_inflight) { CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0)); c->copy_inflight = false; }
                c->stream_first_fvp = false;
                c->path_used = TRPO_PATH_FUSED;
                c->solve_kernel_used = true;
                CU(cudaMemcpyAsync(dResult, c->d_x, c->net.P * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
                return 0;
            }
        }
        c->solve_kernel_used = false;
    }
    // A solve is 3 launches per iteration; for small batches (armDOF_0 x 50 k states: a whole FVP is ~30 us) the launch
    // gaps are a fifth of the solve, so an identical repeated solve is captured once into a CUDA graph and replayed.
    static const bool no_graph = getenv("TRPO_NO_GRAPH") != nullptr;
    const int path = (c->path_req == TRPO_PATH_AUTO) ? (fused_eligible(c->net) ? TRPO_PATH_FUSED : TRPO_PATH_GEMM_CHAIN) : c->path_req;
    const bool eligible = !no_graph && !c->comm && !c->ktime_on && !c->copy_inflight && path == TRPO_PATH_FUSED &&
                          c->precision == TRPO_PRECISION_FP64 && c->d_obs && c->n_local;
    if (eligible) {
        trpo_ctx::CgKey key;
        memset(&key, 0, sizeof(key));            // padding bytes take part in the memcmp below
        key.db = db; key.dres = dResult; key.iters = MaxIter; key.th = ResidualTh; key.damping = damping;
        key.obs = c->d_obs; key.n = c->n_local; key.st = c->stream; key.path = path;
        const bool same = memcmp(&key, &c->cg_key, sizeof(key)) == 0;
        if (same && c->cg_exec) {
            CU(cudaGraphLaunch(c->cg_exec, c->stream));
            c->launches += c->cg_graph_launches;
            c->path_used = path;
            return 0;
        }
        if (same && c->cg_key_seen >= 1) {
            if (c->cg_exec) { cudaGraphExecDestroy(c->cg_exec); c->cg_exec = nullptr; }
            const long long l0 = c->launches;
            cudaGraph_t graph = nullptr;
            CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            const int rc = cg_enqueue(c, db, dResult, MaxIter, ResidualTh, damping);
            const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            c->launches = l0;
            if (rc == 0 && e == cudaSuccess && graph && cudaGraphInstantiate(&c->cg_exec, graph, 0) == cudaSuccess) {
                c->cg_graph_launches = 3 * (long long)MaxIter + 1;
                cudaGraphDestroy(graph);
                CU(cudaGraphLaunch(c->cg_exec, c->stream));
                c->launches += c->cg_graph_launches;
                return 0;
            }
            if (graph) cudaGraphDestroy(graph);
            c->cg_exec = nullptr;
            cudaGetLastError();
            c->cg_key_seen = -1000000;           // capture is not possible here: stay on direct launches
        }
        if (!same) { if (c->cg_exec) { cudaGraphExecDestroy(c->cg_exec); c->cg_exec = nullptr; } c->cg_key = key; c->cg_key_seen = 0; }
        ++c->cg_key_seen;
    }
    if (cg_enqueue(c, db, dResult, MaxIter, ResidualTh, damping)) return -1;
    CU(cudaGetLastError());
    return 0;
}

// Synchronise the context stream and fail the call if a device-side wait timed out since the last check: a peer that
// never pushed its FVP sum (peer-memory all-reduce) or a staged chunk that never landed. The flags are reset, the results
// of the call are not to be used (the kernels poison them with NaN / stop the solve).
static int sync_and_check(trpo_ctx *c) {
    c->h_flags[0] = c->h_flags[1] = 0;
    if (c->p2p_buf) CU(cudaMemcpyAsync(&c->h_flags[0], c->p2p_buf + 268, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(&c->h_flags[1], c->d_ready + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->h_flags[0] || c->h_flags[1]) {
        const int e0 = c->h_flags[0], e1 = c->h_flags[1];
        if (e0) cudaMemsetAsync(c->p2p_buf + 268, 0, sizeof(int), c->stream);
        if (e1) cudaMemsetAsync(c->d_ready + 1, 0, sizeof(int), c->stream);
        cudaStreamSynchronize(c->stream);
        return fail(e0 ? "peer-memory all-reduce timed out: a rank never delivered its FVP sum (results discarded)"
                       : "streamed batch staging timed out: a chunk of the observation matrix never landed (results discarded)");
    }
    return 0;
}

static int fetch_cg_info(trpo_ctx *c) {
    CU(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_trace, c->d_trace, 2 * (size_t)c->trace_cap * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (sync_and_check(c)) return -1;
    c->info.cg_iters = c->h_state->iters;
    const int keep = (int)(sizeof(c->info.cg_rdotr) / sizeof(double));
    for (int i = 0; i < keep; ++i) {
        const bool have = i <= c->h_state->iters && i < c->trace_cap;
        c->info.cg_rdotr[i] = have ? c->h_trace[i] : 0.0;
        c->info.cg_xnorm[i] = have ? c->h_trace[c->trace_cap + i] : 0.0;
    }
    return 0;
}

extern "C" int trpo_ctx_get_cg_trace(const trpo_ctx *c, double *rdotr_out, double *xnorm_out, size_t max_entries) {
    if (!c) return -1;
    const size_t have = (size_t)c->info.cg_iters + 1 < (size_t)c->trace_cap ? (size_t)c->info.cg_iters + 1 : (size_t)c->trace_cap;
    const size_t n = have < max_entries ? have : max_entries;
    for (size_t i = 0; i < n; ++i) {
        if (rdotr_out) rdotr_out[i] = c->h_trace[i];
        if (xnorm_out) xnorm_out[i] = c->h_trace[c->trace_cap + i];
    }
    return (int)n;
}

// Host-buffer entry points with the peer-memory exchange: one tiny NCCL all-reduce first, so that ranks which enter
// seconds apart (rollouts, L-BFGS, first-use module loads) are aligned before any kernel starts its bounded spin.
static int p2p_align_ranks(trpo_ctx *c) {
    if (!c->comm || !active_p2p_ctx(c)) return 0;
    NC(g_nccl.AllReduce(c->d_scal + 15, c->d_scal + 15, 1, ncclFloat64_, ncclSum_, c->comm, c->stream));
    return 0;
}

// P-length host vector -> device. While a streamed batch copy is in flight the host-to-device copy engine is busy with it for
// tens of milliseconds, and a cudaMemcpyAsync of the direction / right-hand side queues behind it: the solve would start only
// when the whole batch has landed (measured: first Humanoid-size FVP 100.7 ms = 54 ms copy + 46 ms compute, no overlap).
// So the vector goes through a pinned buffer that a kernel reads over PCIe (UVA: page-locked host memory is device-visible).
__global__ void k_pull_vector(double *__restrict__ dst, const double *__restrict__ src_host, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src_host[i];
}
static int upload_vector(trpo_ctx *c, double *d_dst, const double *src_host) {
    const size_t bytes = c->net.P * sizeof(double);
    if (!c->copy_inflight) {
        CU(cudaMemcpyAsync(d_dst, src_host, bytes, cudaMemcpyHostToDevice, c->stream));
        return 0;
    }
    memcpy(c->h_vec, src_host, bytes);
    k_pull_vector<<<(c->net.P + 255) / 256, 256, 0, c->stream>>>(d_dst, c->h_vec, c->net.P);
    ++c->launches;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int trpo_ctx_fvp(trpo_ctx *c, const double *Input, double *Result, double damping) {
    if (!c || !Input || !Result) return fail("null argument");
    CU(cudaSetDevice(c->device));
    const size_t bytes = c->net.P * sizeof(double);
    if (upload_vector(c, c->d_in, Input)) return -1;
    if (p2p_align_ranks(c)) return -1;
    if (trpo_ctx_fvp_device(c, c->d_in, c->d_out, damping)) return -1;
    CU(cudaMemcpyAsync(Result, c->d_out, bytes, cudaMemcpyDeviceToHost, c->stream));
    return sync_and_check(c);
}

extern "C" int trpo_ctx_cg(trpo_ctx *c, const double *b, double *Result, size_t MaxIter, double ResidualTh, double damping) {
    if (!c || !b || !Result) return fail("null argument");
    CU(cudaSetDevice(c->device));
    const size_t bytes = c->net.P * sizeof(double);
    if (upload_vector(c, c->d_b, b)) return -1;
    if (p2p_align_ranks(c)) return -1;
    if (trpo_ctx_cg_device(c, c->d_b, c->d_out, MaxIter, ResidualTh, damping)) return -1;
    CU(cudaMemcpyAsync(Result, c->d_out, bytes, cudaMemcpyDeviceToHost, c->stream));
    return fetch_cg_info(c);
}

extern "C" int trpo_ctx_get_info(const trpo_ctx *c, trpo_info *info) {
    if (!c || !info) return fail("null argument");
    *info = c->info;
    return 0;
}

// policy gradient b = (1/N) sum_n grad (TRPO_Update.c:254-378) into c->d_b.
// own_mean_out != NULL (fused path only): the seed uses the network's own output instead of the batch's Mean column and
// the output is stored there -- prediction and gradient of the baseline objective in one pass.
static bool fused_pg_eligible(const trpo_ctx *c) {
    return fused_eligible(c->net) && c->path_req != TRPO_PATH_GEMM_CHAIN && c->precision == TRPO_PRECISION_FP64;
}
static int policy_gradient_device(trpo_ctx *c, double *own_mean_out = nullptr) {
    if (!c->d_obs || !c->d_mean || !c->d_action || !c->d_adv || c->n_full != c->n_local || c->n_local == 0)
        return fail("policy gradient needs Mean/Action/Advantage in the batch (the current batch was staged without them)");
    if (c->copy_inflight) { CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0)); c->copy_inflight = false; }
    c->stream_first_fvp = false;
    if (ensure_chain_scratch(c)) return -1;      // the line search's forward pass uses it in any case
    const bool fused_pg = fused_pg_eligible(c);
    if (own_mean_out && !fused_pg) return fail("internal: own-mean policy gradient needs the fused kernel");
    if (fused_pg) {
        if (fused_pg_accumulate(c->net, c->d_theta, c->d_inv_std_model, c->d_obs, own_mean_out ? nullptr : c->d_mean, c->d_action,
                                c->d_adv, c->n_local, c->d_fused_partial, c->d_zsum, own_mean_out, c->stream, &c->launches))
            return fail("fused policy-gradient launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    } else if (chain_accumulate(c->net, c->sc, CHAIN_PG, c->d_theta, nullptr, nullptr, c->d_obs, c->d_mean, c->d_action, c->d_adv,
                                c->n_local, c->d_zsum, nullptr, nullptr, c->stream, &c->launches))
        return fail("policy-gradient launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (c->comm) NC(g_nccl.AllReduce(c->d_zsum, c->d_zsum, c->net.P, ncclFloat64_, ncclSum_, c->comm, c->stream));
    // b = zsum / N: same kernel as the FVP finalise with no damping and no LogStd special case
    launch_fvp_finalise(c->d_zsum, c->d_zsum, c->d_b, c->net.P, c->net.P, (double)c->n_total, 0.0, nullptr, c->stream, &c->launches);
    return 0;
}

extern "C" int trpo_ctx_forward(trpo_ctx *c, double *Mean_out) {
    if (!c || !Mean_out) return fail("null argument");
    if (!c->d_obs || c->n_local == 0) return fail("no batch staged: call trpo_ctx_set_batch first");
    CU(cudaSetDevice(c->device));
    if (c->copy_inflight) { CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0)); c->copy_inflight = false; }
    c->stream_first_fvp = false;
    if (ensure_chain_scratch(c)) return -1;
    const size_t n = c->n_local * (size_t)c->net.L[c->net.K];
    if (n > c->cap_mean_new) {
        if (c->d_mean_new) cudaFree(c->d_mean_new);
        c->d_mean_new = nullptr;
        CU(cudaMalloc(&c->d_mean_new, n * sizeof(double)));
        c->cap_mean_new = n;
    }
    if (chain_forward(c->net, c->sc, c->d_theta, c->d_obs, c->n_local, c->d_mean_new, c->stream, &c->launches))
        return fail("forward launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    CU(cudaMemcpyAsync(Mean_out, c->d_mean_new, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int trpo_ctx_policy_gradient(trpo_ctx *c, double *b_out) {
    if (!c || !b_out) return fail("null argument");
    CU(cudaSetDevice(c->device));
    if (policy_gradient_device(c)) return -1;
    CU(cudaMemcpyAsync(b_out, c->d_b, c->net.P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

__global__ void k_linesearch_point(double *__restrict__ out, const double *__restrict__ theta, const double *__restrict__ x,
                                   double lm, double stepfrac, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = theta[i] + stepfrac * (x[i] / lm);     // fullstep = x / lm (TRPO_Update.c:836-838,899-902)
}

static int allreduce_scalars(trpo_ctx *c, double *d, int n) {
    if (c->comm) NC(g_nccl.AllReduce(d, d, n, ncclFloat64_, ncclSum_, c->comm, c->stream));
    return 0;
}

extern "C" int trpo_ctx_update(trpo_ctx *c, double *Result, double damping) {
    if (!c || !Result) return fail("null argument");
    CU(cudaSetDevice(c->device));
    // constants of TRPO_Update.c:29-33
    const double ResidualTh = 1e-10, MaxKL = 0.01, AcceptRatio = 0.1;
    const size_t MaxIter = 10, MaxBackTracks = 10;
    const int P = c->net.P, A = c->net.L[c->net.K];
    memset(&c->info, 0, sizeof(c->info));
    if (p2p_align_ranks(c)) return -1;
    if (policy_gradient_device(c)) return -1;
    if (trpo_ctx_cg_device(c, c->d_b, c->d_out, MaxIter, ResidualTh, damping)) return -1;   // d_x = stepdir
    if (trpo_ctx_fvp_device(c, c->d_x, c->d_z, damping)) return -1;                          // z = F x + damping x
    launch_dot(c->d_z, c->d_x, P, c->d_scal + 0, c->stream, &c->launches);                   // 2*shs
    launch_dot(c->d_b, c->d_b, P, c->d_scal + 1, c->stream, &c->launches);                   // gnorm^2
    launch_dot(c->d_b, c->d_x, P, c->d_scal + 2, c->stream, &c->launches);                   // -g . stepdir
    launch_sum(c->d_adv, c->n_local, c->d_blockpart, c->d_scal + 3, c->stream, &c->launches);
    if (allreduce_scalars(c, c->d_scal + 3, 1)) return -1;
    CU(cudaMemcpyAsync(c->h_scal, c->d_scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (fetch_cg_info(c)) return -1;
    const double shs = 0.5 * c->h_scal[0];
    const double lm = sqrt(shs / MaxKL);
    const double gnorm = sqrt(c->h_scal[1]);
    const double rate = c->h_scal[2] / lm;
    const double fval = -c->h_scal[3] / (double)c->n_total;
    c->info.shs = shs; c->info.lm = lm; c->info.gnorm = gnorm; c->info.fval = fval;
    if (c->n_local * A > c->cap_mean_new) {
        if (c->d_mean_new) cudaFree(c->d_mean_new);
        c->d_mean_new = nullptr;
        CU(cudaMalloc(&c->d_mean_new, c->n_local * A * sizeof(double)));
        c->cap_mean_new = c->n_local * A;
    }
    // fall-back result = step direction, the reference's quirk (TRPO_Update.c:852)
    const double *d_result = c->d_x;
    for (size_t t = 0; t < MaxBackTracks; ++t) {
        const double stepfrac = pow(0.5, (double)t);
        k_linesearch_point<<<(P + 255) / 256, 256, 0, c->stream>>>(c->d_xnew, c->d_theta, c->d_x, lm, stepfrac, P);
        ++c->launches;
        if (chain_forward(c->net, c->sc, c->d_xnew, c->d_obs, c->n_local, c->d_mean_new, c->stream, &c->launches))
            return fail("line-search forward launch failed");
        launch_surrogate(c->d_mean_new, c->d_mean, c->d_action, c->d_adv, c->d_std, c->d_xnew + c->net.logstd_off, A,
                         c->n_local, c->d_blockpart, c->d_scal + 4, c->stream, &c->launches);
        if (allreduce_scalars(c, c->d_scal + 4, 1)) return -1;
        CU(cudaMemcpyAsync(c->h_scal + 4, c->d_scal + 4, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        const double newfval = -c->h_scal[4] / (double)c->n_total;
        const double actual = fval - newfval;
        const double expected = rate * stepfrac;
        const double ratio = actual / expected;
        c->info.ls_actual[t] = actual; c->info.ls_expected[t] = expected; c->info.ls_ratio[t] = ratio;
        c->info.ls_steps = (int)t + 1;
        if (ratio > AcceptRatio && actual > 0) { c->info.ls_accepted = 1; d_result = c->d_xnew; break; }
    }
    CU(cudaMemcpyAsync(Result, d_result, P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return sync_and_check(c);
}


// --------------------------------------------------------------------------------------------------------------
// Rows f-3 / f-4: rollout staging, advantage estimation and the baseline ("value function") objective.
extern "C" int trpo_ctx_set_rollout(trpo_ctx *c, size_t NumEpBatch, size_t EpLen, const double *Observ, const double *Std,
                                    const double *Mean, const double *Action, const double *Reward) {
    if (!c || !Observ || !Std || !Mean || !Action || !Reward || NumEpBatch == 0 || EpLen == 0) return fail("bad rollout arguments");
    const size_t N = NumEpBatch * EpLen;
    // Reward is staged twice: once as the (not yet valid) Advantage so that all batch buffers exist, once on its own
    if (trpo_ctx_set_batch(c, N, Observ, Std, Mean, Action, Reward)) return -1;
    if (N > c->cap_reward) {
        if (c->d_reward) cudaFree(c->d_reward);
        c->d_reward = nullptr;
        CU(cudaMalloc(&c->d_reward, N * sizeof(double)));
        c->cap_reward = N;
    }
    CU(cudaMemcpyAsync(c->d_reward, c->d_adv, N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->ep_len = EpLen;
    return 0;
}

// Rollout producer on the device (the role TRPO_RunLightweight plays for the FPGA build, TRPO_Lightweight_FPGA.c:548-556):
// the lightweight arm simulator driven by the CURRENT model, writing Observ / Mean / Action / Reward straight into the
// context's batch buffers -- nothing but the random draws crosses PCIe.
extern "C" int trpo_ctx_rollout_arm(trpo_ctx *c, size_t NumEpBatch, size_t EpLen, const int *RandDraws, unsigned long long Seed) {
    if (!c || NumEpBatch == 0 || EpLen == 0) return fail("bad rollout arguments");
    if (EpLen > 0x7fffffff) return fail("episode too long");
    CU(cudaSetDevice(c->device));
    const size_t N = NumEpBatch * EpLen, O = c->net.L[0], A = c->net.L[c->net.K];
    if (c->copy_inflight) { CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0)); c->copy_inflight = false; }
    c->stream_first_fvp = false;
    if (!c->own_batch) free_batch(c);
    c->own_batch = true;
    if (N * O > c->cap_obs) { cudaFree(c->d_obs); c->d_obs = nullptr; CU(cudaMalloc(&c->d_obs, N * O * sizeof(double))); c->cap_obs = N * O; }
    if (N * A > c->cap_mean) {
        cudaFree(c->d_mean); cudaFree(c->d_action); c->d_mean = c->d_action = nullptr;
        CU(cudaMalloc(&c->d_mean, N * A * sizeof(double)));
        CU(cudaMalloc(&c->d_action, N * A * sizeof(double)));
        c->cap_mean = N * A;
    }
    if (N > c->cap_adv) { cudaFree(c->d_adv); c->d_adv = nullptr; CU(cudaMalloc(&c->d_adv, N * sizeof(double))); c->cap_adv = N; }
    if (N > c->cap_reward) { cudaFree(c->d_reward); c->d_reward = nullptr; CU(cudaMalloc(&c->d_reward, N * sizeof(double))); c->cap_reward = N; }
    const int *d_draws = nullptr;
    if (RandDraws) {
        const size_t nd = NumEpBatch * (3 + 2 * A * EpLen);
        if (nd > c->cap_draws) { cudaFree(c->d_draws); c->d_draws = nullptr; CU(cudaMalloc(&c->d_draws, nd * sizeof(int))); c->cap_draws = nd; }
        CU(cudaMemcpyAsync(c->d_draws, RandDraws, nd * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        d_draws = c->d_draws;
    }
    const int rc = launch_arm_rollout(c->net, c->d_theta, d_draws, Seed, NumEpBatch, (int)EpLen, c->d_obs, c->d_mean, c->d_action,
                                      c->d_reward, c->stream, &c->launches);
    if (rc > 0) return fail("the arm simulator needs a 15-...-3 policy with hidden widths <= 32");
    if (rc < 0) return fail("rollout launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    c->n_local = N;
    c->n_full = N;
    c->ep_len = EpLen;
    double std_host[TRPO_MAX_LAYERS * 64];
    for (size_t j = 0; j < A; ++j) std_host[j] = exp(c->h_logstd[j]);              // TRPO_Lightweight.c:460
    if (set_std(c, std_host)) return -1;                                           // synchronises: RandDraws may be reused
    return update_global_samples(c);
}

extern "C" int trpo_ctx_get_rollout(trpo_ctx *c, double *Observ, double *Mean, double *Action, double *Reward) {
    if (!c) return fail("null context");
    if (!c->d_obs || !c->d_reward || c->n_local == 0) return fail("no rollout staged");
    CU(cudaSetDevice(c->device));
    if (c->copy_inflight) { CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0)); c->copy_inflight = false; }
    const size_t N = c->n_local, O = c->net.L[0], A = c->net.L[c->net.K];
    if (Observ) CU(cudaMemcpyAsync(Observ, c->d_obs, N * O * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (Mean) CU(cudaMemcpyAsync(Mean, c->d_mean, N * A * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (Action) CU(cudaMemcpyAsync(Action, c->d_action, N * A * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (Reward) CU(cudaMemcpyAsync(Reward, c->d_reward, N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

struct trpo_vf {
    trpo_ctx *pol, *net;       // policy context (borrowed) and the baseline network's own context
    size_t npar;               // parameters of the baseline network (no LogStd tail)
    size_t n, ep_len;          // bound batch
    double *d_aug;             // [n x (O+1)] observation + step/EpLen
    double *d_pred, *d_target, *d_coef;   // [n] prediction, regression target, constant gradient seed factor
    size_t cap_aug, cap_n;
    double *h_theta, *h_g;     // host staging (npar + 1)
    bool target_set;
    int failed;                // sticky: an objective evaluation failed (L-BFGS cannot tell -1.0 from a low objective value)
};

extern "C" trpo_vf *trpo_vf_create(trpo_ctx *pol, const size_t *LayerSizeBase, const char *AcFunc, size_t NumLayers) {
    if (!pol || !LayerSizeBase || !AcFunc || NumLayers < 2) { fail("bad value-function description"); return nullptr; }
    if (LayerSizeBase[0] != (size_t)pol->net.L[0] + 1) { fail("LayerSizeBase[0] must be ObservSpaceDim + 1 (observation, step/EpLen)"); return nullptr; }
    if (LayerSizeBase[NumLayers - 1] != 1) { fail("the value-function network must have one output"); return nullptr; }
    for (size_t i = 1; i < NumLayers; ++i)
        if (AcFunc[i] != 'l' && AcFunc[i] != 't') {
            // TRPO_Baseline.c:126-129,152-156: only linear and tanh layers exist for the baseline
            fprintf(stderr, "[ERROR] Activation Function for Layer [%zu] is %c. Unsupported.\n", i, AcFunc[i]);
            fail("unsupported baseline activation '%c'", AcFunc[i]);
            return nullptr;
        }
    trpo_ctx *net = trpo_ctx_create(LayerSizeBase, AcFunc, NumLayers, pol->device, TRPO_PRECISION_FP64);
    if (!net) return nullptr;
    trpo_vf *vf = (trpo_vf *)calloc(1, sizeof(trpo_vf));
    vf->pol = pol; vf->net = net;
    vf->npar = (size_t)net->net.logstd_off;
    vf->h_theta = (double *)calloc(vf->npar + 1, sizeof(double));
    vf->h_g = (double *)calloc(vf->npar + 1, sizeof(double));
    return vf;
}

extern "C" void trpo_vf_destroy(trpo_vf *vf) {
    if (!vf) return;
    cudaSetDevice(vf->net->device);
    cudaStreamSynchronize(vf->net->stream);
    trpo_ctx_destroy(vf->net);                   // adopted batch pointers are not freed by the context
    double *bufs[] = {vf->d_aug, vf->d_pred, vf->d_target, vf->d_coef};
    for (double *b : bufs) if (b) cudaFree(b);
    free(vf->h_theta); free(vf->h_g);
    free(vf);
}

extern "C" size_t trpo_vf_num_params(const trpo_vf *vf) { return vf ? vf->npar : 0; }

extern "C" int trpo_vf_bind_batch(trpo_vf *vf, size_t EpLen) {
    if (!vf) return fail("null value function");
    trpo_ctx *pol = vf->pol, *net = vf->net;
    if (!pol->d_obs || pol->n_local == 0) return fail("no batch staged in the policy context");
    if (EpLen == 0 || pol->n_local % EpLen) return fail("NumSamples (%zu) is not a multiple of EpLen (%zu)", pol->n_local, EpLen);
    CU(cudaSetDevice(pol->device));
    if (pol->copy_inflight) { CU(cudaStreamWaitEvent(pol->stream, pol->ev_copy, 0)); pol->copy_inflight = false; }
    pol->stream_first_fvp = false;
    if (net->stream != pol->stream && trpo_ctx_set_stream(net, pol->stream)) return -1;
    net->comm = pol->comm; net->borrowed_comm = true; net->rank = pol->rank; net->world = pol->world;
    net->path_req = (pol->path_req == TRPO_PATH_GEMM_CHAIN) ? TRPO_PATH_GEMM_CHAIN : TRPO_PATH_AUTO;   // follow the policy's choice
    const size_t n = pol->n_local, O = pol->net.L[0];
    if (n * (O + 1) > vf->cap_aug) {
        if (vf->d_aug) cudaFree(vf->d_aug);
        vf->d_aug = nullptr;
        CU(cudaMalloc(&vf->d_aug, n * (O + 1) * sizeof(double)));
        vf->cap_aug = n * (O + 1);
    }
    if (n > vf->cap_n) {
        double **bufs[] = {&vf->d_pred, &vf->d_target, &vf->d_coef};
        for (double **b : bufs) { if (*b) cudaFree(*b); *b = nullptr; CU(cudaMalloc(b, n * sizeof(double))); }
        vf->cap_n = n;
        vf->target_set = false;
    }
    launch_vf_augment(pol->d_obs, n, (int)O, EpLen, vf->d_aug, pol->stream, &net->launches);
    // gradient seed 0.02 * (Predict - Target) (TRPO_Baseline.c:141) == Advantage * (Action - Mean) / sigma^2 of the policy
    // gradient (TRPO_Update.c:298-300) with Action := Target, Mean := Predict, Advantage := -0.02, sigma := 1
    launch_fill(vf->d_coef, -0.02, n, pol->stream, &net->launches);
    const double one = 1.0;
    if (trpo_ctx_set_batch_device(net, n, vf->d_aug, &one, vf->d_pred, vf->d_target, vf->d_coef)) return -1;
    vf->n = n; vf->ep_len = EpLen;
    return 0;
}

extern "C" int trpo_vf_set_target(trpo_vf *vf, const double *Target) {
    if (!vf || !Target) return fail("null argument");
    if (vf->n == 0) return fail("trpo_vf_bind_batch first");
    CU(cudaSetDevice(vf->net->device));
    CU(cudaMemcpyAsync(vf->d_target, Target, vf->n * sizeof(double), cudaMemcpyHostToDevice, vf->net->stream));
    CU(cudaStreamSynchronize(vf->net->stream));
    vf->target_set = true;
    return 0;
}

// forward pass of the baseline network for parameters x into vf->d_pred (stays on the stream)
static int vf_forward(trpo_vf *vf, const double *x) {
    trpo_ctx *net = vf->net;
    if (vf->n == 0) return fail("trpo_vf_bind_batch first");
    memcpy(vf->h_theta, x, vf->npar * sizeof(double));
    vf->h_theta[vf->npar] = 0.0;                 // LogStd slot of the embedding context: sigma = 1
    if (trpo_ctx_set_model(net, vf->h_theta)) return -1;
    if (ensure_chain_scratch(net)) return -1;
    if (chain_forward(net->net, net->sc, net->d_theta, vf->d_aug, vf->n, vf->d_pred, net->stream, &net->launches))
        return fail("baseline forward launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

extern "C" int trpo_vf_predict(trpo_vf *vf, const double *x, double *Predict_out) {
    if (!vf || !x || !Predict_out) return fail("null argument");
    CU(cudaSetDevice(vf->net->device));
    if (vf_forward(vf, x)) return -1;
    CU(cudaMemcpyAsync(Predict_out, vf->d_pred, vf->n * sizeof(double), cudaMemcpyDeviceToHost, vf->net->stream));
    CU(cudaStreamSynchronize(vf->net->stream));
    return 0;
}

extern "C" int trpo_vf_advantage(trpo_vf *vf, const double *x, double gamma, double lam, double *Return_out, double *Advantage_out) {
    if (!vf || !x) return fail("null argument");
    trpo_ctx *pol = vf->pol;
    if (!pol->d_reward || pol->ep_len == 0) return fail("no rollout staged: call trpo_ctx_set_rollout first");
    CU(cudaSetDevice(pol->device));
    if (trpo_vf_bind_batch(vf, pol->ep_len)) return -1;
    if (vf_forward(vf, x)) return -1;
    const size_t n = vf->n;
    cudaStream_t st = pol->stream;
    launch_gae(pol->d_reward, vf->d_pred, n / pol->ep_len, (int)pol->ep_len, gamma, lam, vf->d_target, pol->d_adv, st, &pol->launches);
    vf->target_set = true;
    // standardise over the whole (global) batch: mean, then the squared deviations, both fixed-order
    launch_sum(pol->d_adv, n, pol->d_blockpart, pol->d_scal + 8, st, &pol->launches);
    if (allreduce_scalars(pol, pol->d_scal + 8, 1)) return -1;
    launch_sqdiff(pol->d_adv, nullptr, pol->d_scal + 8, (double)pol->n_total, n, pol->d_blockpart, pol->d_scal + 9, st, &pol->launches);
    if (allreduce_scalars(pol, pol->d_scal + 9, 1)) return -1;
    launch_standardise(pol->d_adv, n, pol->d_scal + 8, pol->d_scal + 9, (double)pol->n_total, st, &pol->launches);
    if (Return_out) CU(cudaMemcpyAsync(Return_out, vf->d_target, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (Advantage_out) CU(cudaMemcpyAsync(Advantage_out, pol->d_adv, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int trpo_vf_failed(trpo_vf *vf) {
    if (!vf) return 1;
    const int f = vf->failed;
    vf->failed = 0;
    return f;
}

static double vf_evaluate_impl(trpo_vf *vf, const double *x, double *g, const int n);
extern "C" double trpo_vf_evaluate(void *instance, const double *x, double *g, const int n, const double step) {
    (void)step;
    trpo_vf *vf = (trpo_vf *)instance;
    const double fx = vf_evaluate_impl(vf, x, g, n);
    if (vf && fx < 0.0) {
        // the objective 0.01*MSE + 0.001*|x|^2 is never negative: -1 is the failure value. Make it sticky and hand L-BFGS
        // an infinite objective so that the line search backs off instead of accepting a "lower" value.
        vf->failed = 1;
        if (g) for (int i = 0; i < n; ++i) g[i] = 0.0;
        return HUGE_VAL;
    }
    return fx;
}
static double vf_evaluate_impl(trpo_vf *vf, const double *x, double *g, const int n) {
    if (!vf || !x || !g) { fail("null argument"); return -1.0; }
    if ((size_t)n < vf->npar) { fail("n (%d) is smaller than the number of baseline parameters (%zu)", n, vf->npar); return -1.0; }
    if (!vf->target_set) { fail("no regression target: trpo_vf_set_target or trpo_vf_advantage first"); return -1.0; }
    trpo_ctx *net = vf->net;
    if (cudaSetDevice(net->device) != cudaSuccess) { fail("cudaSetDevice failed"); return -1.0; }
    // net->d_b = (1/N) sum_n dLoss_n/dx. Shapes the fused kernel takes: prediction and gradient in one pass over the
    // batch; otherwise a forward GEMM chain for the prediction, then the chain's gradient pass.
    if (fused_pg_eligible(net)) {
        if (vf->n == 0) { fail("trpo_vf_bind_batch first"); return -1.0; }
        memcpy(vf->h_theta, x, vf->npar * sizeof(double));
        vf->h_theta[vf->npar] = 0.0;
        if (trpo_ctx_set_model(net, vf->h_theta)) return -1.0;
        if (policy_gradient_device(net, vf->d_pred)) return -1.0;
    } else {
        if (vf_forward(vf, x)) return -1.0;
        if (policy_gradient_device(net)) return -1.0;
    }
    launch_sqdiff(vf->d_pred, vf->d_target, nullptr, 1.0, vf->n, net->d_blockpart, net->d_scal, net->stream, &net->launches);
    if (allreduce_scalars(net, net->d_scal, 1)) return -1.0;
    if (cudaMemcpyAsync(vf->h_g, net->d_b, vf->npar * sizeof(double), cudaMemcpyDeviceToHost, net->stream) != cudaSuccess ||
        cudaMemcpyAsync(net->h_scal, net->d_scal, sizeof(double), cudaMemcpyDeviceToHost, net->stream) != cudaSuccess ||
        sync_and_check(net) != 0) {
        fail("baseline objective failed: %s", cudaGetErrorString(cudaGetLastError()));
        return -1.0;
    }
    double l2 = 0.0;
    for (size_t i = 0; i < vf->npar; ++i) {
        g[i] = vf->h_g[i] + 0.002 * x[i];                                // TRPO_Baseline.c:205-214
        l2 += x[i] * x[i];
    }
    for (int i = (int)vf->npar; i < n; ++i) g[i] = 0.0;
    return 0.01 * net->h_scal[0] / (double)net->n_total + 0.001 * l2;    // TRPO_Baseline.c:220-235
}

// --------------------------------------------------------------------------------------------------------------
extern "C" double *trpo_device_alloc(size_t n) {
    double *p = nullptr;
    if (cudaMalloc(&p, n * sizeof(double)) != cudaSuccess) { fail("cudaMalloc failed"); return nullptr; }
    return p;
}
extern "C" void trpo_device_free(double *p) { if (p) cudaFree(p); }
// page-locked host arrays for C callers (a pinned Observ source lets trpo_ctx_set_batch stream the copy under the first FVP)
extern "C" double *trpo_host_alloc_pinned(size_t n) {
    double *p = nullptr;
    if (cudaMallocHost(&p, (n ? n : 1) * sizeof(double)) != cudaSuccess) { cudaGetLastError(); fail("cudaMallocHost failed"); return nullptr; }
    return p;
}
extern "C" void trpo_host_free_pinned(double *p) { if (p) cudaFreeHost(p); }
extern "C" int trpo_memcpy_h2d(double *dst, const double *src, size_t n) { CU(cudaMemcpy(dst, src, n * sizeof(double), cudaMemcpyHostToDevice)); return 0; }
extern "C" int trpo_memcpy_d2h(double *dst, const double *src, size_t n) { CU(cudaMemcpy(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost)); return 0; }

extern "C" int trpo_nccl_unique_id(char id_out[128]) {
    if (nccl_load()) return -1;
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id_out, id.internal, 128);
    return 0;
}

extern "C" int trpo_ctx_init_comm(trpo_ctx *c, const char id[128], int rank, int world) {
    if (!c || !id) return fail("null argument");
    if (world <= 1) { c->rank = 0; c->world = 1; return 0; }
    if (nccl_load()) return -1;
    CU(cudaSetDevice(c->device));
    ncclUniqueId uid;
    memcpy(uid.internal, id, 128);
    NC(g_nccl.CommInitRank(&c->comm, world, uid, rank));
    c->rank = rank; c->world = world;
    if (c->n_local) return update_global_samples(c);
    return 0;
}

// ---- peer-memory all-reduce plumbing ------------------------------------------------------------------------
// buffer layout: [0,256) flags[2][8] u64 | 256: seq_dev u64 | 264: block_counter u32 | 268: error i32 | 1024: slots[2][world][P]
#define P2P_HDR 1024
extern "C" int trpo_ctx_p2p_export(trpo_ctx *c, char handle_out[64]) {
    if (!c || !handle_out) return fail("null argument");
    if (c->world < 2) return fail("trpo_ctx_init_comm with world_size >= 2 first");
    if (c->world > TRPO_MAX_RANKS) return fail("peer-memory all-reduce supports at most %d ranks", TRPO_MAX_RANKS);
    CU(cudaSetDevice(c->device));
    if (!c->p2p_buf) {
        // header | slots[2][world][P] doubles | (fused shapes) tagged words [2][world][P][2] of the persistent solve kernel
        const size_t slot_bytes = sizeof(double) * 2 * (size_t)c->world * c->net.P;
        const size_t bytes = P2P_HDR + slot_bytes + (fused_eligible(c->net) ? 2 * slot_bytes : 0);
        CU(cudaMalloc(&c->p2p_buf, bytes));
        CU(cudaMemset(c->p2p_buf, 0, bytes));
        CU(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->p2p_buf));
    static_assert(sizeof(h) == 64, "CUDA IPC handle size");
    memcpy(handle_out, &h, 64);
    return 0;
}

extern "C" int trpo_ctx_p2p_attach(trpo_ctx *c, const char *handles) {
    if (!c || !handles) return fail("null argument");
    if (!c->p2p_buf) return fail("call trpo_ctx_p2p_export first");
    CU(cudaSetDevice(c->device));
    memset(&c->p2p, 0, sizeof(c->p2p));
    for (int r = 0; r < c->world; ++r) {
        char *base;
        if (r == c->rank) base = c->p2p_buf;
        else {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles + 64 * r, 64);
            void *ptr = nullptr;
            CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            c->p2p_peer[r] = ptr;
            base = (char *)ptr;
        }
        c->p2p.flags[r] = (unsigned long long *)base;
        c->p2p.ll[r] = fused_eligible(c->net) ? (unsigned long long *)(base + P2P_HDR + sizeof(double) * 2 * (size_t)c->world * c->net.P) : nullptr;
        c->p2p.slots[r] = (double *)(base + P2P_HDR);
    }
    c->p2p.seq_dev = (unsigned long long *)(c->p2p_buf + 256);
    c->p2p.block_counter = (unsigned int *)(c->p2p_buf + 264);
    c->p2p.error = (int *)(c->p2p_buf + 268);
    c->p2p.world = c->world;
    c->p2p.rank = c->rank;
    c->p2p.P = c->net.P;
    c->p2p_on = true;
    return 0;
}

extern "C" int trpo_ctx_set_comm_mode(trpo_ctx *c, int mode) {
    if (!c) return fail("null context");
    if (mode == TRPO_COMM_P2P && c->p2p.world < 2) return fail("peer buffers are not attached");
    c->p2p_on = (mode == TRPO_COMM_P2P);
    return 0;
}

extern "C" int trpo_ctx_comm_error(trpo_ctx *c) {
    if (!c) return 0;
    int e = 0, e2 = 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->p2p_buf) cudaMemcpy(&e, c->p2p_buf + 268, sizeof(int), cudaMemcpyDeviceToHost);
    if (c->d_ready) cudaMemcpy(&e2, c->d_ready + 1, sizeof(int), cudaMemcpyDeviceToHost);   // streamed-staging wait timed out
    return e | e2;
}

extern "C" size_t trpo_ctx_global_samples(const trpo_ctx *c) { return c ? c->n_total : 0; }
extern "C" int trpo_ctx_solve_kernel_used(const trpo_ctx *c) { return c && c->solve_kernel_used ? 1 : 0; }
// Phase stamps of the persistent solve kernel (CTA 0, %globaltimer, nanoseconds): enable with max_iters > 0, then after a solve
// and a sync read max_iters x 8 values: [0] pass start, [1] own pass done, [2] all CTAs' passes done, [3] slice sums formed,
// [4] peers' slices arrived, [5] p.z complete, [6] r.r complete, [7] new direction published.
extern "C" int trpo_ctx_solve_timeline(trpo_ctx *c, size_t max_iters, unsigned long long *out) {
    if (!c) return fail("null context");
    CU(cudaSetDevice(c->device));
    if (out && c->d_timeline) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaMemcpy(out, c->d_timeline, (max_iters < c->timeline_iters ? max_iters : c->timeline_iters) * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        return 0;
    }
    if (c->d_timeline) { cudaFree(c->d_timeline); c->d_timeline = nullptr; }
    c->timeline_iters = 0;
    if (max_iters) {
        CU(cudaMalloc(&c->d_timeline, max_iters * 8 * sizeof(unsigned long long)));
        CU(cudaMemset(c->d_timeline, 0, max_iters * 8 * sizeof(unsigned long long)));
        c->timeline_iters = max_iters;
    }
    return 0;
}
