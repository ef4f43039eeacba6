// fvp_fused.cu -- fused per-sample Fisher-vector-product kernel for 4-layer policies whose weights fit in shared memory.
//
// Computes the un-normalised sum over samples of the reference's FVPFast loop (TRPO_FVP.c:771-924) on the FP64 tensor
// pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; tcgen05 has no f64 kind).  One persistent CTA per SM, NW warps, tiles of
// S = 8*NW samples:
//   phase A (per warp, 8 samples, no block sync): combined forward + R-forward (:783-836), R-gradient seed (:852-854)
//            and R-backward (:857-900) chained through REGISTERS: the m8n8 accumulator layout of one layer is used
//            directly as the A fragment of the next by permuting the contraction index, which only changes which
//            weight rows a B fragment holds -- and the weights sit in shared memory in exactly that fragment order.
//            W1/W2 are stored once, in an XOR-swizzled 8x8-block layout that serves both the forward (W) and the
//            backward (W^T) fragment reads without bank conflicts.
//   phase B (block-wide): the parameter-gradient outer products [Y_{i-1},1]^T * RG_i (:890-921), contraction over the
//            S samples of the tile; each warp owns a fixed set of 8x8 output tiles whose accumulators stay in registers
//            for the whole kernel.  Activations are exchanged through shared memory [sample][neuron] (+4 padding:
//            conflict-free transposed fragment reads).
// At the end every CTA writes its P-length partial row; rows are summed in fixed order by k_reduce_partials, so the
// result is bitwise deterministic.
#include <stdlib.h>

#include "trpo_internal.cuh"
#include "dmma_common.cuh"
#include "tmem_scratch.cuh"

namespace {

// y = f(x) and d = f'(x) written through y (TRPO_FVP.c:806-834,869-882). ACT == 0: runtime switch on `a`.
template <char ACT>
__device__ __forceinline__ void act_fwd(char a, double x, double &y, double &d) {
    const char k = ACT ? ACT : a;
    if (k == 't') { y = tanh(x); d = 1.0 - y * y; }
    else if (k == 's') { y = 1.0 / (1.0 + exp(-x)); d = y * (1.0 - y); }
    else if (k == 'o') { y = 0.1 * x; d = 0.1; }
    else { y = x; d = 1.0; }
}
template <char ACT>
__device__ __forceinline__ double act_deriv_y(char a, double y) {
    const char k = ACT ? ACT : a;
    if (k == 't') return 1.0 - y * y;
    if (k == 's') return y * (1.0 - y);
    if (k == 'o') return 0.1;
    return 1.0;
}

// Activation of CNT accumulator tiles starting at tile C0: x <- f(x), rx <- rx * f'(x)
template <char ACT, int C0, int CNT, int NTILES>
__device__ __forceinline__ void activate_tiles(char a, double (&x)[NTILES][2], double (&rx)[NTILES][2], const double *tab) {
    if constexpr (ACT == 't') {
        double xv[2 * CNT], dv[2 * CNT];
#pragma unroll
        for (int c = 0; c < CNT; ++c) { xv[2 * c] = x[C0 + c][0]; xv[2 * c + 1] = x[C0 + c][1]; }
        tanh_vec<2 * CNT>(xv, dv, tab);
#pragma unroll
        for (int c = 0; c < CNT; ++c) {
            x[C0 + c][0] = xv[2 * c]; x[C0 + c][1] = xv[2 * c + 1];
            rx[C0 + c][0] *= dv[2 * c]; rx[C0 + c][1] *= dv[2 * c + 1];
        }
    } else {
#pragma unroll
        for (int c = 0; c < CNT; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                double d;
                act_fwd<ACT>(a, x[C0 + c][r], x[C0 + c][r], d);
                rx[C0 + c][r] *= d;
            }
    }
}

__host__ __device__ constexpr int pad_rs(int k) {   // smallest row stride >= k with stride % 16 in {4, 12}
    int r = k;
    while ((r % 16) != 4 && (r % 16) != 12) ++r;
    return r;
}
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }

// position of element (rr, cc) inside a swizzled 8x8 block
__host__ __device__ __forceinline__ int swz(int rr, int cc) {
    const int a = rr >> 1, rp = rr & 1, cq = cc >> 1, cp = cc & 1;
    return (a & 1) + 2 * (cq & 1) + 4 * ((a >> 1) ^ rp) + 8 * (cp ^ (cq >> 1)) + 16 * rp + 32 * (cq >> 1);
}

// NG = number of independent warp groups inside the CTA. NG = 1: the whole CTA works on one tile of S = 8 * NW samples and
// its phases are separated by block barriers. NG = 2: two groups of WG = NW / 2 warps (one warp of each group per SM
// sub-partition) share the staged weights but own separate activation buffers, tiles of SG = 8 * WG samples and named
// barriers, so the two warps of a scheduler are at DIFFERENT places of the per-tile instruction stream (with NG = 1 they
// run in lock step and stall together: the same net is 18 % faster as two independent 4-warp CTAs, profiles/r02_summary.md).
// Each group then needs the full set of parameter-gradient accumulators over half as many warps; they are parked in
// Tensor Memory while phase A runs (tmem_scratch.cuh).
template <int K0_, int H1_, int H2_, int AP_, int NW_, int CTAS_PER_SM_ = 1, int NG_ = 1>
struct Cfg {
    static constexpr int K0 = K0_, H1 = H1_, H2 = H2_, AP = AP_, NW = NW_, CTAS_PER_SM = CTAS_PER_SM_, NG = NG_;
    static constexpr int S = 8 * NW, NTHREADS = 32 * NW;
    static constexpr int WG = NW / NG, SG = 8 * WG;                       // warps / samples per group tile
    static constexpr int Q0 = K0 / 4, MT0 = (K0 + 7) / 8;
    static constexpr int NT1 = H1 / 8, NT2 = H2 / 8, NT3 = AP / 8;
    static constexpr int R1 = (NT1 + WG - 1) / WG, R2 = (NT2 + WG - 1) / WG;   // output tile rows a warp owns in phase B
    static constexpr int RS0 = pad_rs(K0), RS1 = H1 + 4, RS2 = H2 + 4, RS3 = AP + 4, RSB = cmax(RS1, RS2);
    // shared memory carve-up (in doubles): weights and direction (shared by the groups), then per group the observation
    // double buffer and the activation exchange buffers, then the exp2 table
    static constexpr int oW0 = 0, oVW0 = oW0 + K0 * H1, oW1 = oVW0 + K0 * H1, oVW1 = oW1 + H1 * H2,
                         oW2 = oVW1 + H1 * H2, oVW2 = oW2 + H2 * AP, oB0 = oVW2 + H2 * AP, oVB0 = oB0 + H1,
                         oB1 = oVB0 + H1, oVB1 = oB1 + H2, oVB2 = oVB1 + H2, oIV = oVB2 + AP,
                         oY0 = oIV + AP, Y0SZ = SG * RS0 + 8,
                         gA = 2 * Y0SZ, gC = gA + SG * RS1, gB = gC + SG * RS2, gD = gB + SG * RSB, GSZ = gD + SG * RS3,
                         oTab = oY0 + NG * GSZ, TOTAL = oTab + 64;
    static constexpr size_t SMEM_BYTES = sizeof(double) * TOTAL;
    // accumulator doubles per thread (phase B), padded to a multiple of 4 for the TMEM parking area
    static constexpr int NACC = 2 * (MT0 * R1 + R1 + R1 * NT2 + R2 + R2 * NT3 + NT3), PKN = (NACC + 3) / 4 * 4;
    static_assert(K0 % 4 == 0 && H1 % 8 == 0 && H2 % 8 == 0 && AP % 8 == 0, "padded sizes");
    static_assert(NW % NG == 0 && NT3 <= WG, "group shape");
    static_assert(NG == 1 || (2 * PKN * (NW / 4) <= 512), "TMEM columns: NW/4 warps share a 32-lane quadrant");
};

struct FusedArgs {
    const double *theta, *v, *inv_var, *obs;
    double *partial;
    const int *done;
    long long nsamples;
    int L0, L1, L2, L3;
    int w_off0, w_off1, w_off2;
    int P;
    char act1, act2, act3;
    // streamed staging: the observation rows arrive by DMA in chunks of chunk_samples while this kernel already runs;
    // *ready counts the chunks that have landed (written by the copy engine after each chunk). NULL = all resident.
    const int *ready;
    long long chunk_samples;
    int *error;
    // policy-gradient mode (TRPO_Update.c:254-378): ordinary forward / backward with the surrogate-loss seed
    const double *mean, *action, *adv;       // [N x A], [N x A], [N]
    int logstd_off;
    // mean == NULL: the kernel forms the network output itself (one more DMMA step per sample) and, when mean_out is
    // set, stores it -- the baseline objective's Predict (TRPO_Baseline.c:134) and its gradient in ONE pass
    double *mean_out;
    // persistent solve kernel: the grid barrier that publishes the direction v is split -- the CTA arrived before calling the
    // pass and waits (thread 0 on the generation word, then a block barrier) only right before it reads v, with the first
    // observation tile already on its way
    unsigned int *wait_bar;
    unsigned int wait_gen;
};
__device__ __forceinline__ void pass_grid_wait(const FusedArgs &p) {
    if (p.wait_bar == nullptr) return;
    if (threadIdx.x == 0) {
        unsigned int g;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(p.wait_bar + 1) : "memory");
            if (g == p.wait_gen && clock64() - t0 > 20000000000LL) __trap();
        } while (g == p.wait_gen);
    }
    __syncthreads();
}

__device__ __forceinline__ int ld_volatile_i32(const int *p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Block until the chunk holding sample `last_sample` has been copied in (bounded spin: ~4 s, then flag an error; the
// host entry points read the flag after their synchronisation and fail the call).
__device__ __forceinline__ void wait_samples(const FusedArgs &p, long long last_sample) {
    if (p.ready == nullptr) return;
    const int need = (int)(last_sample / p.chunk_samples) + 1;
    if (ld_volatile_i32(p.ready) >= need) return;
    const long long t0 = clock64();
    while (ld_volatile_i32(p.ready) < need) {
        if (clock64() - t0 > 8000000000LL) { *(volatile int *)p.error = 2; break; }
    }
}

// PG = true turns the same kernel into the policy-gradient kernel: no R{} quantities (a third of the forward DMMAs, no
// layer-2 forward at all: the seed uses the rollout's own Mean), seed g3 = A (a - mu) / sigma^2 * f', and the LogStd
// gradient sum_n A (((a - mu)/sigma)^2 - 1) accumulated per thread; backward pass and outer products are shared.
// In PG mode p.inv_var holds 1/sigma = exp(-LogStd) of the CURRENT parameters (TRPO_Update.c:298-300), p.v is unused.
// stage: bit 0 = the parts that depend on the model and the batch header (W, B, 1/sigma^2, tables, zeroed buffers), bit 1 = the
// parts that depend on the direction v (VW, VB). The stand-alone kernel stages both; the persistent solve kernel stages the
// model once and the direction once per CG iteration (the direction is read with ld.global.cg: another CTA wrote it).
template <typename C, char ACT1, char ACT2, bool PG>
__device__ __forceinline__ void fused_pass(const FusedArgs &p, double *sm, const int stage) {
    constexpr int K0 = C::K0, H1 = C::H1, H2 = C::H2, AP = C::AP, NT = C::NTHREADS;
    constexpr int NG = C::NG, WG = C::WG, SG = C::SG, GT = 32 * WG, R1 = C::R1, R2 = C::R2;
    constexpr int Q0 = C::Q0, MT0 = C::MT0, NT1 = C::NT1, NT2 = C::NT2, NT3 = C::NT3;
    constexpr int RS0 = C::RS0, RS1 = C::RS1, RS2 = C::RS2, RS3 = C::RS3, RSB = C::RSB;
    constexpr bool PARK = NG > 1;
    double *W0f = sm + C::oW0, *VW0f = sm + C::oVW0, *W1s = sm + C::oW1, *VW1s = sm + C::oVW1;
    double *W2s = sm + C::oW2, *VW2s = sm + C::oVW2, *B0s = sm + C::oB0, *VB0s = sm + C::oVB0;
    double *B1s = sm + C::oB1, *VB1s = sm + C::oVB1, *VB2s = sm + C::oVB2, *IVs = sm + C::oIV;
    double *Tab = sm + C::oTab;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int grp = w / WG, wg = w % WG, tg = tid - grp * GT;      // warp group, warp and thread index inside the group
    double *Y0s = sm + C::oY0 + grp * C::GSZ;                     // this group's observation double buffer + exchange buffers
    double *BufA = Y0s + C::gA, *BufC = Y0s + C::gC, *BufB = Y0s + C::gB, *BufD = Y0s + C::gD;
    const int L0 = p.L0, L1 = p.L1, L2 = p.L2, L3 = p.L3;
    // barrier over this group's warps only (named barrier 1 + grp); the whole CTA when there is one group
    auto group_sync = [&]() {
        if constexpr (NG == 1) __syncthreads();
        else asm volatile("bar.sync %0, %1;" :: "r"(1 + grp), "n"(GT) : "memory");
    };

    // ---------------- prologue: weights and direction into shared memory, in fragment order ----------------
    // A padded input column (K0 > L0) carries the layer-0 bias for free: the observation tile holds 1.0 in column L0 and
    // row L0 of the staged weights is the bias row -- which is simply the next row of the flat [W0;B0] matrix. The bias
    // then needs neither an accumulator init nor its own gradient tile (row L0 of Y0^T * RG1 is the bias gradient).
    const bool free_col = L0 < K0;
    const int rows0 = L0 + (free_col ? 1 : 0);
    const bool st_m = (stage & 1) != 0, st_v = (stage & 2) != 0;
    // Work split over the Wk = gridDim.x * NG warp groups: F full rounds of round-robin tiles (tile wid + r * Wk: the batch is
    // consumed front to back, which is what the streamed staging needs), then the remaining < Wk * SG samples are dealt out in
    // 8-sample units, contiguous and as evenly as possible, ONE partial tile per group. A partial tile costs the active warps'
    // phase A plus a phase B over its rows only, so 13.2 tiles per group take 13.4 tile times instead of 14 (the 8-GPU shard of
    // the headline batch: 125 k states over 148 CTAs).
    const long long Wk = (long long)gridDim.x * NG, wid = (long long)blockIdx.x * NG + grp;
    const long long F = p.nsamples / (Wk * SG), sT = F * Wk * SG;
    const long long Ut = (p.nsamples - sT + 7) / 8, ubase = Ut / Wk, uextra = Ut % Wk;
    const long long tail_lo = sT + 8 * (wid * ubase + (wid < uextra ? wid : uextra));
    const long long tail_hi0 = tail_lo + 8 * (ubase + (wid < uextra ? 1 : 0));
    const long long tail_hi = tail_hi0 < p.nsamples ? tail_hi0 : p.nsamples;
    const int nk = (int)F + (tail_hi > tail_lo ? 1 : 0);            // tiles of this group
    auto tile_begin = [&](int k) { return k < F ? (wid + (long long)k * Wk) * SG : tail_lo; };
    auto tile_end = [&](int k) { return k < F ? (wid + (long long)k * Wk) * SG + SG : tail_hi; };
    // observation tile [SG][K0] (zero padded / zero past the end of the tile), staged asynchronously one tile ahead
    auto stage_obs = [&](int k, int buf) {
        double *dst = Y0s + buf * C::Y0SZ;
        const long long s0n = tile_begin(k), s1n = tile_end(k);
        wait_samples(p, s1n - 1);
        for (int idx = tg; idx < SG * K0; idx += GT) {
            const int row = idx / K0, col = idx % K0;
            const long long gs = s0n + row;
            const bool in = gs < s1n && col < L0;
            if (free_col && col == L0) dst[row * RS0 + col] = (gs < s1n) ? 1.0 : 0.0;
            else cp_async8(&dst[row * RS0 + col], in ? &p.obs[gs * L0 + col] : p.obs, in ? 8 : 0);
        }
    };

    // later passes of the persistent solve: the buffers are already clean, so the first observation tile is requested before
    // anything else and lands while the CTA waits for the new direction and stages it
    if (!st_m && nk > 0) stage_obs(0, 0);
    pass_grid_wait(p);
    for (int idx = tid; idx < K0 * H1; idx += NT) {
        const int k = idx / H1, n = idx % H1;
        const bool in = k < rows0 && n < L1;
        const int dst = ((k >> 2) * NT1 + (n >> 3)) * 32 + (n & 7) * 4 + (k & 3);
        if (st_m) W0f[dst]  = in ? p.theta[p.w_off0 + k * L1 + n] : 0.0;
        if (!PG && st_v) VW0f[dst] = in ? __ldcg(&p.v[p.w_off0 + k * L1 + n]) : 0.0;
    }
    for (int idx = tid; idx < H1 * H2; idx += NT) {
        const int j = idx / H2, n = idx % H2;
        const bool in = j < L1 && n < L2;
        const int dst = ((j >> 3) * NT2 + (n >> 3)) * 64 + swz(j & 7, n & 7);
        if (st_m) W1s[dst]  = in ? p.theta[p.w_off1 + j * L2 + n] : 0.0;
        if (!PG && st_v) VW1s[dst] = in ? __ldcg(&p.v[p.w_off1 + j * L2 + n]) : 0.0;
    }
    for (int idx = tid; idx < H2 * AP; idx += NT) {
        const int j = idx / AP, n = idx % AP;
        const bool in = j < L2 && n < L3;
        const int dst = ((j >> 3) * NT3 + (n >> 3)) * 64 + swz(j & 7, n & 7);
        if (st_m) W2s[dst]  = in ? p.theta[p.w_off2 + j * L3 + n] : 0.0;
        if (!PG && st_v) VW2s[dst] = in ? __ldcg(&p.v[p.w_off2 + j * L3 + n]) : 0.0;
    }
    for (int n = tid; n < H1; n += NT) {
        if (st_m) B0s[n]  = n < L1 ? p.theta[p.w_off0 + L0 * L1 + n] : 0.0;
        if (st_v) VB0s[n] = (!PG && n < L1) ? __ldcg(&p.v[p.w_off0 + L0 * L1 + n]) : 0.0;
    }
    for (int n = tid; n < H2; n += NT) {
        if (st_m) B1s[n]  = n < L2 ? p.theta[p.w_off1 + L1 * L2 + n] : 0.0;
        if (st_v) VB1s[n] = (!PG && n < L2) ? __ldcg(&p.v[p.w_off1 + L1 * L2 + n]) : 0.0;
    }
    for (int n = tid; n < AP; n += NT) {
        if (PG ? st_m : st_v) VB2s[n] = n >= L3 ? 0.0 : (PG ? p.theta[p.w_off2 + L2 * L3 + n] : __ldcg(&p.v[p.w_off2 + L2 * L3 + n]));   // PG: plain bias B2
        if (st_m) IVs[n]  = n < L3 ? p.inv_var[n] : 0.0;
    }
    if (st_m) {
        load_exp2_table(Tab);
        for (int i = tid; i < NG * C::GSZ; i += NT) sm[C::oY0 + i] = 0.0;
    }
    // Tensor Memory for the parked accumulators: 2 * PKN columns per warp, the NW / 4 warps of a quadrant side by side
    __shared__ uint32_t tmem_slot;
    constexpr uint32_t TM_COLS = (2 * C::PKN * (C::NW / 4) <= 32) ? 32 : (2 * C::PKN * (C::NW / 4) <= 64) ? 64
                               : (2 * C::PKN * (C::NW / 4) <= 128) ? 128 : (2 * C::PKN * (C::NW / 4) <= 256) ? 256 : 512;
    if constexpr (PARK) {
        if (w == 0) tmem_alloc(&tmem_slot, TM_COLS);
        tmem_fence_before_sync();
    }
    __syncthreads();
    uint32_t tm = 0;
    if constexpr (PARK) {
        tmem_fence_after_sync();
        tm = tmem_warp_addr(tmem_slot, w, (w >> 2) * 2 * C::PKN);
    }
    // per-lane offsets into a swizzled block: forward fragment (row 2t+r, col g), transposed fragment (row g, col 2t+r)
    int sf[2], sb[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) { sf[r] = swz(2 * t + r, g); sb[r] = swz(g, 2 * t + r); }
    const double d3 = (p.act3 == 'o') ? 0.1 : 1.0;      // last layer is linear or 0.1x on this path
    const double ones = (g == 0) ? 1.0 : 0.0;            // A fragment whose row 0 is all ones: column sums (bias gradients)

    // ---------------- persistent parameter-gradient accumulators (C-fragment layout) ----------------
    // Warp wg of a group owns R1 = ceil(NT1 / WG) row blocks of the layer-1 outer product and the same column blocks of the
    // layer-0 one, R2 row blocks of the layer-2 one and R2 bias tiles of layer 1; the group's last warp the layer-2 bias.
    double acc0[MT0][R1][2], accb0[R1][2];       // [Y0;1]^T * G1, output column blocks R1*wg + jj
    double acc1[R1][NT2][2], accb1[R2][2];       // rows 8*(R1*wg+rr).. of Y1^T * G2; bias tiles R2*wg + jj of layer 1
    double acc2[R2][NT3][2], accb2[NT3][2];      // rows 8*(R2*wg+rr).. of Y2^T * G3; last warp: bias of layer 2
#pragma unroll
    for (int i = 0; i < MT0; ++i)
#pragma unroll
        for (int j = 0; j < R1; ++j) acc0[i][j][0] = acc0[i][j][1] = 0.0;
#pragma unroll
    for (int i = 0; i < R1; ++i) {
        accb0[i][0] = accb0[i][1] = 0.0;
#pragma unroll
        for (int j = 0; j < NT2; ++j) acc1[i][j][0] = acc1[i][j][1] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < R2; ++i) {
        accb1[i][0] = accb1[i][1] = 0.0;
#pragma unroll
        for (int j = 0; j < NT3; ++j) acc2[i][j][0] = acc2[i][j][1] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < NT3; ++i) accb2[i][0] = accb2[i][1] = 0.0;
    double gl[NT3][2];                  // PG: this thread's share of the LogStd gradient (its row, its two columns per tile)
#pragma unroll
    for (int i = 0; i < NT3; ++i) gl[i][0] = gl[i][1] = 0.0;

    // NG > 1: the accumulators live in Tensor Memory while phase A runs (tcgen05.st / tcgen05.ld, SASS STTM / LDTM)
    auto park = [&]() {
        double pk[C::PKN];
        int k = 0;
#pragma unroll
        for (int i = 0; i < MT0; ++i)
#pragma unroll
            for (int j = 0; j < R1; ++j) { pk[k++] = acc0[i][j][0]; pk[k++] = acc0[i][j][1]; }
#pragma unroll
        for (int i = 0; i < R1; ++i) {
            pk[k++] = accb0[i][0]; pk[k++] = accb0[i][1];
#pragma unroll
            for (int j = 0; j < NT2; ++j) { pk[k++] = acc1[i][j][0]; pk[k++] = acc1[i][j][1]; }
        }
#pragma unroll
        for (int i = 0; i < R2; ++i) {
            pk[k++] = accb1[i][0]; pk[k++] = accb1[i][1];
#pragma unroll
            for (int j = 0; j < NT3; ++j) { pk[k++] = acc2[i][j][0]; pk[k++] = acc2[i][j][1]; }
        }
#pragma unroll
        for (int i = 0; i < NT3; ++i) { pk[k++] = accb2[i][0]; pk[k++] = accb2[i][1]; }
#pragma unroll
        for (; k < C::PKN; ++k) pk[k] = 0.0;
        tmem_park<C::PKN>(tm, pk);
    };
    auto unpark = [&]() {
        double pk[C::PKN];
        tmem_wait_st();
        tmem_fetch<C::PKN>(tm, pk);
        int k = 0;
#pragma unroll
        for (int i = 0; i < MT0; ++i)
#pragma unroll
            for (int j = 0; j < R1; ++j) { acc0[i][j][0] = pk[k++]; acc0[i][j][1] = pk[k++]; }
#pragma unroll
        for (int i = 0; i < R1; ++i) {
            accb0[i][0] = pk[k++]; accb0[i][1] = pk[k++];
#pragma unroll
            for (int j = 0; j < NT2; ++j) { acc1[i][j][0] = pk[k++]; acc1[i][j][1] = pk[k++]; }
        }
#pragma unroll
        for (int i = 0; i < R2; ++i) {
            accb1[i][0] = pk[k++]; accb1[i][1] = pk[k++];
#pragma unroll
            for (int j = 0; j < NT3; ++j) { acc2[i][j][0] = pk[k++]; acc2[i][j][1] = pk[k++]; }
        }
#pragma unroll
        for (int i = 0; i < NT3; ++i) { accb2[i][0] = pk[k++]; accb2[i][1] = pk[k++]; }
    };

    const int rowA = 8 * wg + g;                         // this lane's sample row inside the group's tile (phase A)
    if (st_m && nk > 0) stage_obs(0, 0);
    cp_async_wait_all();
    group_sync();
    int buf = 0;
    for (int k = 0; k < nk; ++k, buf ^= 1) {
        const long long s0 = tile_begin(k), s1 = tile_end(k);
        const int hmax = (int)((s1 - s0 + 7) / 8) * 2;   // phase-B k-steps (4 rows each) that hold samples; SG / 4 for a full tile
        const bool warp_active = s0 + 8 * wg < s1;       // a partial tile leaves the upper warps without samples
        const double *Y0c = Y0s + buf * C::Y0SZ;
        if (k + 1 < nk) stage_obs(k + 1, buf ^ 1);       // lands during this tile's math
        if constexpr (PARK) park();

        // ======================= phase A: this warp's 8 samples =======================
        double g1[NT1][2];
        if (warp_active) {
        double y1[NT1][2], ry1[NT1][2];
        {   // layer 0: x1 = [y0,1]*[W0;B0], Rx1 = [y0,1]*[VW0;VB0]   (Ry0 = 0)
#pragma unroll
            for (int c = 0; c < NT1; ++c)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    y1[c][r] = free_col ? 0.0 : B0s[8 * c + 2 * t + r];
                    ry1[c][r] = free_col ? 0.0 : VB0s[8 * c + 2 * t + r];
                }
#pragma unroll
            for (int q = 0; q < Q0; ++q) {
                const double a = Y0c[rowA * RS0 + 4 * q + t];
#pragma unroll
                for (int c = 0; c < NT1; ++c) {
                    dmma(y1[c], a, W0f[(q * NT1 + c) * 32 + lane]);
                    if (!PG) dmma(ry1[c], a, VW0f[(q * NT1 + c) * 32 + lane]);
                }
            }
            constexpr int ACH = NT1 < 4 ? NT1 : 4;       // activation chunk: 8 values in flight per thread
            static_assert(NT1 % ACH == 0, "layer-1 width must split into equal activation chunks");
            if constexpr (NT1 >= 1 * ACH) activate_tiles<ACT1, 0 * ACH, ACH>(p.act1, y1, ry1, Tab);
            if constexpr (NT1 >= 2 * ACH) activate_tiles<ACT1, 1 * ACH, ACH>(p.act1, y1, ry1, Tab);
            if constexpr (NT1 >= 3 * ACH) activate_tiles<ACT1, 2 * ACH, ACH>(p.act1, y1, ry1, Tab);
            if constexpr (NT1 >= 4 * ACH) activate_tiles<ACT1, 3 * ACH, ACH>(p.act1, y1, ry1, Tab);
            static_assert(NT1 <= 4 * ACH, "add activation chunks");
#pragma unroll
            for (int c = 0; c < NT1; ++c)
                *reinterpret_cast<double2 *>(&BufA[rowA * RS1 + 8 * c + 2 * t]) = make_double2(y1[c][0], y1[c][1]);
        }
        // layer 1 in column groups of <= 4 output tiles (keeps the live register set small); each finished group is
        // consumed immediately by layer 2 (only Rx3 is needed: y3 does not enter the FVP when the last layer is linear).
        constexpr int GRP = NT2 < 4 ? NT2 : 4;
        static_assert(NT2 % GRP == 0, "layer-2 width must split into equal column groups");
        double rx3p[GRP][NT3][2];            // GRP independent partial sums of Rx3
#pragma unroll
        for (int cc = 0; cc < GRP; ++cc)
#pragma unroll
            for (int c = 0; c < NT3; ++c)
#pragma unroll
                for (int r = 0; r < 2; ++r) rx3p[cc][c][r] = (cc == 0) ? VB2s[8 * c + 2 * t + r] : 0.0;
        const bool self_mean = PG && p.mean == nullptr;      // uniform: x3 = y2*W2 + B2 accumulates in rx3p
#pragma unroll
        for (int c0 = 0; c0 < NT2; c0 += GRP) {
            double x2[GRP][2], rx2[GRP][2];
#pragma unroll
            for (int cc = 0; cc < GRP; ++cc)
#pragma unroll
                for (int r = 0; r < 2; ++r) { x2[cc][r] = B1s[8 * (c0 + cc) + 2 * t + r]; rx2[cc][r] = VB1s[8 * (c0 + cc) + 2 * t + r]; }
#pragma unroll
            for (int b = 0; b < NT1; ++b)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    double bw[GRP], bv[GRP];
#pragma unroll
                    for (int cc = 0; cc < GRP; ++cc) {
                        bw[cc] = W1s[(b * NT2 + c0 + cc) * 64 + sf[r]];
                        if (!PG) bv[cc] = VW1s[(b * NT2 + c0 + cc) * 64 + sf[r]];
                    }
#pragma unroll
                    for (int cc = 0; cc < GRP; ++cc) {
                        if (!PG) dmma(rx2[cc], ry1[b][r], bw[cc]);
                        dmma(x2[cc], y1[b][r], bw[cc]);
                        if (!PG) dmma(rx2[cc], y1[b][r], bv[cc]);
                    }
                }
            activate_tiles<ACT2, 0, GRP>(p.act2, x2, rx2, Tab);
#pragma unroll
            for (int cc = 0; cc < GRP; ++cc)
                *reinterpret_cast<double2 *>(&BufC[rowA * RS2 + 8 * (c0 + cc) + 2 * t]) = make_double2(x2[cc][0], x2[cc][1]);
            // layer 2 contribution of these k-blocks: Rx3 += Ry2*W2 + y2*VW2 (the policy gradient needs no layer-2 forward)
            if (!PG) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < NT3; ++c) {
#pragma unroll
                    for (int cc = 0; cc < GRP; ++cc) dmma(rx3p[cc][c], rx2[cc][r], W2s[((c0 + cc) * NT3 + c) * 64 + sf[r]]);
#pragma unroll
                    for (int cc = 0; cc < GRP; ++cc) dmma(rx3p[cc][c], x2[cc][r], VW2s[((c0 + cc) * NT3 + c) * 64 + sf[r]]);
                }
            } else if (self_mean) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < NT3; ++c)
#pragma unroll
                    for (int cc = 0; cc < GRP; ++cc) dmma(rx3p[cc][c], x2[cc][r], W2s[((c0 + cc) * NT3 + c) * 64 + sf[r]]);
            }
        }
        double rx3[NT3][2];
#pragma unroll
        for (int c = 0; c < NT3; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                double sum = rx3p[0][c][r];
#pragma unroll
                for (int cc = 1; cc < GRP; ++cc) sum += rx3p[cc][c][r];
                rx3[c][r] = sum;
            }
        // R-gradient seed: RG3 = Ry3 / sigma^2 (times the constant f' of the last layer, twice: Ry3 = f' Rx3, RG3 *= f')
        const bool valid = (s0 + rowA) < s1;             // rows past the end of the tile contribute nothing
        double g3[NT3][2];
        if (PG) {
            // surrogate-loss seed (TRPO_Update.c:297-301,310-324): t = (a - mu)/sigma, g = A t / sigma * f', dLogStd = A (t^2 - 1)
            const long long gs = s0 + rowA;
            const double adv = valid ? p.adv[gs] : 0.0;
#pragma unroll
            for (int c = 0; c < NT3; ++c)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int col = 8 * c + 2 * t + r;
                    double gv = 0.0;
                    if (valid && col < L3) {
                        const double is = IVs[col];
                        double mu;
                        if (self_mean) {
                            mu = rx3[c][r] * d3;             // last layer is 'l' (y = x) or 'o' (y = 0.1 x)
                            if (p.mean_out) p.mean_out[gs * L3 + col] = mu;
                        } else mu = p.mean[gs * L3 + col];
                        const double tt = (p.action[gs * L3 + col] - mu) * is;
                        gv = adv * tt * is * d3;
                        gl[c][r] += adv * (tt * tt - 1.0);
                    }
                    g3[c][r] = gv;
                }
        } else {
#pragma unroll
            for (int c = 0; c < NT3; ++c)
#pragma unroll
                for (int r = 0; r < 2; ++r) g3[c][r] = valid ? rx3[c][r] * d3 * IVs[8 * c + 2 * t + r] * d3 : 0.0;
        }
#pragma unroll
        for (int c = 0; c < NT3; ++c)
            *reinterpret_cast<double2 *>(&BufD[rowA * RS3 + 8 * c + 2 * t]) = make_double2(g3[c][0], g3[c][1]);
        // backward through layer 2: RG2 = (RG3 * W2^T) .* f'(y2)
        double g2[NT2][2];
#pragma unroll
        for (int b = 0; b < NT2; ++b) g2[b][0] = g2[b][1] = 0.0;
#pragma unroll
        for (int c = 0; c < NT3; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int b = 0; b < NT2; ++b) dmma(g2[b], g3[c][r], W2s[(b * NT3 + c) * 64 + sb[r]]);
#pragma unroll
        for (int b = 0; b < NT2; ++b) {
            const double2 y = *reinterpret_cast<const double2 *>(&BufC[rowA * RS2 + 8 * b + 2 * t]);
            g2[b][0] *= act_deriv_y<ACT2>(p.act2, y.x);
            g2[b][1] *= act_deriv_y<ACT2>(p.act2, y.y);
            *reinterpret_cast<double2 *>(&BufB[rowA * RSB + 8 * b + 2 * t]) = make_double2(g2[b][0], g2[b][1]);
        }
        // backward through layer 1: RG1 = (RG2 * W1^T) .* f'(y1); stays in registers until BufB is free again
#pragma unroll
        for (int b = 0; b < NT1; ++b) g1[b][0] = g1[b][1] = 0.0;
#pragma unroll
        for (int c = 0; c < NT2; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int b = 0; b < NT1; ++b) dmma(g1[b], g2[c][r], W1s[(b * NT2 + c) * 64 + sb[r]]);
#pragma unroll
        for (int b = 0; b < NT1; ++b) {
            const double2 y = *reinterpret_cast<const double2 *>(&BufA[rowA * RS1 + 8 * b + 2 * t]);
            g1[b][0] *= act_deriv_y<ACT1>(p.act1, y.x);
            g1[b][1] *= act_deriv_y<ACT1>(p.act1, y.y);
        }
        } else {
#pragma unroll
            for (int b = 0; b < NT1; ++b) g1[b][0] = g1[b][1] = 0.0;
        }
        group_sync();
        if constexpr (PARK) unpark();

        // ======================= phase B1: W1 / B1 and W2 / B2 gradients over the group's tile =======================
        if (R1 * wg < NT1 || R2 * wg < NT2 || wg == WG - 1) {
#pragma unroll 4
            for (int h = 0; h < hmax; ++h) {
                const int srow = 4 * h + t;
                if (R1 * wg < NT1) {            // row blocks R1*wg + rr of Y1^T * G2
                    double bg[NT2];
#pragma unroll
                    for (int j = 0; j < NT2; ++j) bg[j] = BufB[srow * RSB + 8 * j + g];
#pragma unroll
                    for (int rr = 0; rr < R1; ++rr) {
                        if (R1 * wg + rr < NT1) {
                            const double a = BufA[srow * RS1 + 8 * (R1 * wg + rr) + g];
#pragma unroll
                            for (int j = 0; j < NT2; ++j) dmma(acc1[rr][j], a, bg[j]);
                        }
                    }
                }
                if (R2 * wg < NT2) {            // bias gradient tiles of layer 1, row blocks R2*wg + rr of Y2^T * G3
                    double bd[NT3];
#pragma unroll
                    for (int c = 0; c < NT3; ++c) bd[c] = BufD[srow * RS3 + 8 * c + g];
#pragma unroll
                    for (int rr = 0; rr < R2; ++rr) {
                        if (R2 * wg + rr < NT2) {
                            dmma(accb1[rr], ones, BufB[srow * RSB + 8 * (R2 * wg + rr) + g]);
                            const double a2 = BufC[srow * RS2 + 8 * (R2 * wg + rr) + g];
#pragma unroll
                            for (int c = 0; c < NT3; ++c) dmma(acc2[rr][c], a2, bd[c]);
                        }
                    }
                }
                if (wg == WG - 1) {
#pragma unroll
                    for (int c = 0; c < NT3; ++c) dmma(accb2[c], ones, BufD[srow * RS3 + 8 * c + g]);
                }
            }
        }
        group_sync();
        // RG1 takes over BufB
#pragma unroll
        for (int b = 0; b < NT1; ++b)
            *reinterpret_cast<double2 *>(&BufB[rowA * RSB + 8 * b + 2 * t]) = make_double2(g1[b][0], g1[b][1]);
        group_sync();
        // ======================= phase B2: W0 / B0 gradients =======================
        if (R1 * wg < NT1) {
#pragma unroll 4
            for (int h = 0; h < hmax; ++h) {
                const int srow = 4 * h + t;
                double ya[MT0];
#pragma unroll
                for (int m = 0; m < MT0; ++m) ya[m] = Y0c[srow * RS0 + 8 * m + g];
#pragma unroll
                for (int jj = 0; jj < R1; ++jj) {
                    if (R1 * wg + jj < NT1) {
                        const double bg = BufB[srow * RSB + 8 * (R1 * wg + jj) + g];
#pragma unroll
                        for (int m = 0; m < MT0; ++m) dmma(acc0[m][jj], ya[m], bg);
                        if (!free_col) dmma(accb0[jj], ones, bg);
                    }
                }
            }
        }
        cp_async_wait_all();
        group_sync();
    }

    if constexpr (PARK) {
        tmem_fence_before_sync();
        __syncthreads();                                     // every warp has fetched its accumulators for the last time
        if (w == 0) tmem_dealloc(tmem_slot, TM_COLS);
    }
    // ---------------- epilogue: this group's partial row (only real, un-padded entries) ----------------
    double *out = p.partial + ((size_t)blockIdx.x * NG + grp) * p.P;
#pragma unroll
    for (int jj = 0; jj < R1; ++jj) {
        const int nb = R1 * wg + jj;                         // column block of layer 0 / row block of layer 1
        if (nb >= NT1) continue;
#pragma unroll
        for (int m = 0; m < MT0; ++m)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = 8 * m + g, col = 8 * nb + 2 * t + r;
                if (row < rows0 && col < L1) out[p.w_off0 + row * L1 + col] = acc0[m][jj][r];
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int col = 8 * nb + 2 * t + r;
            if (!free_col && g == 0 && col < L1) out[p.w_off0 + L0 * L1 + col] = accb0[jj][r];
        }
#pragma unroll
        for (int j = 0; j < NT2; ++j)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = 8 * nb + g, col = 8 * j + 2 * t + r;
                if (row < L1 && col < L2) out[p.w_off1 + row * L2 + col] = acc1[jj][j][r];
            }
    }
#pragma unroll
    for (int jj = 0; jj < R2; ++jj) {
        const int nb = R2 * wg + jj;
        if (nb >= NT2) continue;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int col = 8 * nb + 2 * t + r;
            if (g == 0 && col < L2) out[p.w_off1 + L1 * L2 + col] = accb1[jj][r];
        }
#pragma unroll
        for (int c = 0; c < NT3; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = 8 * nb + g, col = 8 * c + 2 * t + r;
                if (row < L2 && col < L3) out[p.w_off2 + row * L3 + col] = acc2[jj][c][r];
            }
    }
    if (wg == WG - 1) {
#pragma unroll
        for (int c = 0; c < NT3; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int col = 8 * c + 2 * t + r;
                if (g == 0 && col < L3) out[p.w_off2 + L2 * L3 + col] = accb2[c][r];
            }
    }
    if (PG) {
        // LogStd gradient: fixed-order sum over the 8 rows of a warp (shuffle tree), then over the group's warps
        group_sync();
        double *red = BufD;                                  // WG * AP doubles
#pragma unroll
        for (int c = 0; c < NT3; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                double vsum = gl[c][r];
                vsum += __shfl_xor_sync(0xffffffffu, vsum, 4);
                vsum += __shfl_xor_sync(0xffffffffu, vsum, 8);
                vsum += __shfl_xor_sync(0xffffffffu, vsum, 16);
                if (g == 0) red[wg * AP + 8 * c + 2 * t + r] = vsum;
            }
        group_sync();
        if (tg < L3) {
            double vsum = red[tg];
#pragma unroll
            for (int ww = 1; ww < WG; ++ww) vsum += red[ww * AP + tg];
            out[p.logstd_off + tg] = vsum;
        }
    }
}

template <typename C, char ACT1, char ACT2, bool PG>
__global__ void __launch_bounds__(C::NTHREADS, C::CTAS_PER_SM) k_fvp_fused(const FusedArgs p) {
    if (p.done && *p.done) return;
    extern __shared__ __align__(16) double sm[];
    fused_pass<C, ACT1, ACT2, PG>(p, sm, 3);
}


// ---------------------------------------------------------------------------------------------------------------
// Warp-private variant for very small nets (hidden <= 16, e.g. the reference's armDOF_0 policy 15-16-16-3): all
// parameter-gradient accumulators of the whole net fit in one warp's registers (30 doubles per thread), so a warp runs
// phase A AND phase B on its own 8 samples and there is no block barrier inside the sample loop at all. Warps drift
// apart, which keeps the FP64 pipe busy while others sit in tanh or wait for loads, and the work splits in 8-sample
// units (50 k states over 1184 warps: 5.3 units each instead of 5.3 tiles of 64 per CTA with 2 of 8 warps doing the
// outer products). The per-warp sums are added in fixed warp order at the end of the kernel.
template <typename C, char ACT1, char ACT2>
__device__ __forceinline__ void warp_pass(const FusedArgs &p, double *sm, const int stage) {
    constexpr int K0 = C::K0, H1 = C::H1, H2 = C::H2, AP = C::AP, NW = C::NW, NT = C::NTHREADS;
    constexpr int Q0 = C::Q0, MT0 = C::MT0, NT1 = C::NT1, NT2 = C::NT2, NT3 = C::NT3;
    constexpr int RS0 = C::RS0, RS1 = C::RS1, RS2 = C::RS2, RS3 = C::RS3, RSB = C::RSB;
    constexpr int WS = 2 * (8 * RS0 + 8) + 8 * RS1 + 8 * RS2 + 8 * RSB + 8 * RS3;      // per-warp scratch (doubles)
    double *W0f = sm + C::oW0, *VW0f = sm + C::oVW0, *W1s = sm + C::oW1, *VW1s = sm + C::oVW1;
    double *W2s = sm + C::oW2, *VW2s = sm + C::oVW2, *B0s = sm + C::oB0, *VB0s = sm + C::oVB0;
    double *B1s = sm + C::oB1, *VB1s = sm + C::oVB1, *VB2s = sm + C::oVB2, *IVs = sm + C::oIV;
    double *Tab = sm + C::oY0;                              // the region after the weights is laid out locally
    double *scratch = Tab + 64;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int L0 = p.L0, L1 = p.L1, L2 = p.L2, L3 = p.L3;
    const bool free_col = L0 < K0;
    const int rows0 = L0 + (free_col ? 1 : 0);
    const bool st_m = (stage & 1) != 0, st_v = (stage & 2) != 0;
    for (int idx = tid; idx < K0 * H1; idx += NT) {
        const int k = idx / H1, n = idx % H1;
        const bool in = k < rows0 && n < L1;
        const int dst = ((k >> 2) * NT1 + (n >> 3)) * 32 + (n & 7) * 4 + (k & 3);
        if (st_m) W0f[dst]  = in ? p.theta[p.w_off0 + k * L1 + n] : 0.0;
        if (st_v) VW0f[dst] = in ? __ldcg(&p.v[p.w_off0 + k * L1 + n]) : 0.0;
    }
    for (int idx = tid; idx < H1 * H2; idx += NT) {
        const int j = idx / H2, n = idx % H2;
        const bool in = j < L1 && n < L2;
        const int dst = ((j >> 3) * NT2 + (n >> 3)) * 64 + swz(j & 7, n & 7);
        if (st_m) W1s[dst]  = in ? p.theta[p.w_off1 + j * L2 + n] : 0.0;
        if (st_v) VW1s[dst] = in ? __ldcg(&p.v[p.w_off1 + j * L2 + n]) : 0.0;
    }
    for (int idx = tid; idx < H2 * AP; idx += NT) {
        const int j = idx / AP, n = idx % AP;
        const bool in = j < L2 && n < L3;
        const int dst = ((j >> 3) * NT3 + (n >> 3)) * 64 + swz(j & 7, n & 7);
        if (st_m) W2s[dst]  = in ? p.theta[p.w_off2 + j * L3 + n] : 0.0;
        if (st_v) VW2s[dst] = in ? __ldcg(&p.v[p.w_off2 + j * L3 + n]) : 0.0;
    }
    for (int n = tid; n < H1; n += NT) {
        if (st_m) B0s[n]  = n < L1 ? p.theta[p.w_off0 + L0 * L1 + n] : 0.0;
        if (st_v) VB0s[n] = n < L1 ? __ldcg(&p.v[p.w_off0 + L0 * L1 + n]) : 0.0;
    }
    for (int n = tid; n < H2; n += NT) {
        if (st_m) B1s[n]  = n < L2 ? p.theta[p.w_off1 + L1 * L2 + n] : 0.0;
        if (st_v) VB1s[n] = n < L2 ? __ldcg(&p.v[p.w_off1 + L1 * L2 + n]) : 0.0;
    }
    for (int n = tid; n < AP; n += NT) {
        if (st_v) VB2s[n] = n < L3 ? __ldcg(&p.v[p.w_off2 + L2 * L3 + n]) : 0.0;
        if (st_m) IVs[n]  = n < L3 ? p.inv_var[n] : 0.0;
    }
    if (st_m) load_exp2_table(Tab);
    for (int i = tid; i < NW * WS; i += NT) scratch[i] = 0.0;      // every pass: the per-warp sums below reuse this area
    __syncthreads();

    double *Y0w = scratch + w * WS;                         // [2][8*RS0 + 8]
    double *Y1w = Y0w + 2 * (8 * RS0 + 8), *Y2w = Y1w + 8 * RS1, *Gw = Y2w + 8 * RS2, *G3w = Gw + 8 * RSB;
    int sf[2], sb[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) { sf[r] = swz(2 * t + r, g); sb[r] = swz(g, 2 * t + r); }
    const double d3 = (p.act3 == 'o') ? 0.1 : 1.0;
    const double ones = (g == 0) ? 1.0 : 0.0;

    double acc0[MT0][NT1][2], accb0[NT1][2], acc1[NT1][NT2][2], accb1[NT2][2], acc2[NT2][NT3][2], accb2[NT3][2];
#pragma unroll
    for (int m = 0; m < MT0; ++m)
#pragma unroll
        for (int j = 0; j < NT1; ++j) acc0[m][j][0] = acc0[m][j][1] = 0.0;
#pragma unroll
    for (int j = 0; j < NT1; ++j) accb0[j][0] = accb0[j][1] = 0.0;
#pragma unroll
    for (int m = 0; m < NT1; ++m)
#pragma unroll
        for (int j = 0; j < NT2; ++j) acc1[m][j][0] = acc1[m][j][1] = 0.0;
#pragma unroll
    for (int j = 0; j < NT2; ++j) accb1[j][0] = accb1[j][1] = 0.0;
#pragma unroll
    for (int m = 0; m < NT2; ++m)
#pragma unroll
        for (int c = 0; c < NT3; ++c) acc2[m][c][0] = acc2[m][c][1] = 0.0;
#pragma unroll
    for (int c = 0; c < NT3; ++c) accb2[c][0] = accb2[c][1] = 0.0;

    // this warp's 8-sample units: u = blockIdx.x * NW + w, stride gridDim.x * NW
    const long long nunits = (p.nsamples + 7) / 8;
    const long long ustride = (long long)gridDim.x * NW;
    auto stage_obs = [&](long long unit, int buf) {         // the warp stages its own 8 x K0 tile
        double *dst = Y0w + buf * (8 * RS0 + 8);
        const long long s0n = unit * 8;
        wait_samples(p, (s0n + 8 < p.nsamples ? s0n + 8 : p.nsamples) - 1);
        for (int idx = lane; idx < 8 * K0; idx += 32) {
            const int row = idx / K0, col = idx % K0;
            const long long gs = s0n + row;
            const bool in = gs < p.nsamples && col < L0;
            if (free_col && col == L0) dst[row * RS0 + col] = (gs < p.nsamples) ? 1.0 : 0.0;
            else cp_async8(&dst[row * RS0 + col], in ? &p.obs[gs * L0 + col] : p.obs, in ? 8 : 0);
        }
    };
    long long unit = (long long)blockIdx.x * NW + w;
    if (unit < nunits) stage_obs(unit, 0);
    cp_async_wait_all();
    __syncwarp();
    int buf = 0;
    for (; unit < nunits; unit += ustride, buf ^= 1) {
        const double *Y0c = Y0w + buf * (8 * RS0 + 8);
        if (unit + ustride < nunits) stage_obs(unit + ustride, buf ^ 1);
        // ---- phase A (as in k_fvp_fused, rows = g) ----
        double y1[NT1][2], ry1[NT1][2];
#pragma unroll
        for (int c = 0; c < NT1; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                y1[c][r] = free_col ? 0.0 : B0s[8 * c + 2 * t + r];
                ry1[c][r] = free_col ? 0.0 : VB0s[8 * c + 2 * t + r];
            }
#pragma unroll
        for (int q = 0; q < Q0; ++q) {
            const double a = Y0c[g * RS0 + 4 * q + t];
#pragma unroll
            for (int c = 0; c < NT1; ++c) {
                dmma(y1[c], a, W0f[(q * NT1 + c) * 32 + lane]);
                dmma(ry1[c], a, VW0f[(q * NT1 + c) * 32 + lane]);
            }
        }
        activate_tiles<ACT1, 0, NT1>(p.act1, y1, ry1, Tab);
#pragma unroll
        for (int c = 0; c < NT1; ++c)
            *reinterpret_cast<double2 *>(&Y1w[g * RS1 + 8 * c + 2 * t]) = make_double2(y1[c][0], y1[c][1]);
        double x2[NT2][2], rx2[NT2][2];
#pragma unroll
        for (int c = 0; c < NT2; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r) { x2[c][r] = B1s[8 * c + 2 * t + r]; rx2[c][r] = VB1s[8 * c + 2 * t + r]; }
#pragma unroll
        for (int b = 0; b < NT1; ++b)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < NT2; ++c) {
                    const double bw = W1s[(b * NT2 + c) * 64 + sf[r]], bv = VW1s[(b * NT2 + c) * 64 + sf[r]];
                    dmma(rx2[c], ry1[b][r], bw);
                    dmma(x2[c], y1[b][r], bw);
                    dmma(rx2[c], y1[b][r], bv);
                }
        activate_tiles<ACT2, 0, NT2>(p.act2, x2, rx2, Tab);
#pragma unroll
        for (int c = 0; c < NT2; ++c)
            *reinterpret_cast<double2 *>(&Y2w[g * RS2 + 8 * c + 2 * t]) = make_double2(x2[c][0], x2[c][1]);
        double rx3[NT3][2];
#pragma unroll
        for (int c = 0; c < NT3; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r) rx3[c][r] = VB2s[8 * c + 2 * t + r];
#pragma unroll
        for (int b = 0; b < NT2; ++b)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < NT3; ++c) {
                    dmma(rx3[c], rx2[b][r], W2s[(b * NT3 + c) * 64 + sf[r]]);
                    dmma(rx3[c], x2[b][r], VW2s[(b * NT3 + c) * 64 + sf[r]]);
                }
        const bool valid = (unit * 8 + g) < p.nsamples;
        double g3[NT3][2];
#pragma unroll
        for (int c = 0; c < NT3; ++c) {
#pragma unroll
            for (int r = 0; r < 2; ++r) g3[c][r] = valid ? rx3[c][r] * d3 * IVs[8 * c + 2 * t + r] * d3 : 0.0;
            *reinterpret_cast<double2 *>(&G3w[g * RS3 + 8 * c + 2 * t]) = make_double2(g3[c][0], g3[c][1]);
        }
        double g2[NT2][2];
#pragma unroll
        for (int b = 0; b < NT2; ++b) g2[b][0] = g2[b][1] = 0.0;
#pragma unroll
        for (int c = 0; c < NT3; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int b = 0; b < NT2; ++b) dmma(g2[b], g3[c][r], W2s[(b * NT3 + c) * 64 + sb[r]]);
#pragma unroll
        for (int b = 0; b < NT2; ++b) {
            g2[b][0] *= act_deriv_y<ACT2>(p.act2, x2[b][0]);
            g2[b][1] *= act_deriv_y<ACT2>(p.act2, x2[b][1]);
            *reinterpret_cast<double2 *>(&Gw[g * RSB + 8 * b + 2 * t]) = make_double2(g2[b][0], g2[b][1]);
        }
        double g1[NT1][2];
#pragma unroll
        for (int b = 0; b < NT1; ++b) g1[b][0] = g1[b][1] = 0.0;
#pragma unroll
        for (int c = 0; c < NT2; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int b = 0; b < NT1; ++b) dmma(g1[b], g2[c][r], W1s[(b * NT2 + c) * 64 + sb[r]]);
#pragma unroll
        for (int b = 0; b < NT1; ++b) {
            g1[b][0] *= act_deriv_y<ACT1>(p.act1, y1[b][0]);
            g1[b][1] *= act_deriv_y<ACT1>(p.act1, y1[b][1]);
        }
        __syncwarp();
        // ---- phase B on this warp's own 8 samples (2 k-steps) ----
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int srow = 4 * h + t;
            double bg2[NT2], bg3[NT3];
#pragma unroll
            for (int j = 0; j < NT2; ++j) bg2[j] = Gw[srow * RSB + 8 * j + g];
#pragma unroll
            for (int c = 0; c < NT3; ++c) bg3[c] = G3w[srow * RS3 + 8 * c + g];
#pragma unroll
            for (int m = 0; m < NT1; ++m) {
                const double a = Y1w[srow * RS1 + 8 * m + g];
#pragma unroll
                for (int j = 0; j < NT2; ++j) dmma(acc1[m][j], a, bg2[j]);
            }
#pragma unroll
            for (int j = 0; j < NT2; ++j) dmma(accb1[j], ones, bg2[j]);
#pragma unroll
            for (int m = 0; m < NT2; ++m) {
                const double a = Y2w[srow * RS2 + 8 * m + g];
#pragma unroll
                for (int c = 0; c < NT3; ++c) dmma(acc2[m][c], a, bg3[c]);
            }
#pragma unroll
            for (int c = 0; c < NT3; ++c) dmma(accb2[c], ones, bg3[c]);
        }
        __syncwarp();
#pragma unroll
        for (int b = 0; b < NT1; ++b)
            *reinterpret_cast<double2 *>(&Gw[g * RSB + 8 * b + 2 * t]) = make_double2(g1[b][0], g1[b][1]);
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int srow = 4 * h + t;
            double bg1[NT1];
#pragma unroll
            for (int j = 0; j < NT1; ++j) bg1[j] = Gw[srow * RSB + 8 * j + g];
#pragma unroll
            for (int m = 0; m < MT0; ++m) {
                const double a = Y0c[srow * RS0 + 8 * m + g];
#pragma unroll
                for (int j = 0; j < NT1; ++j) dmma(acc0[m][j], a, bg1[j]);
            }
            if (!free_col) {
#pragma unroll
                for (int j = 0; j < NT1; ++j) dmma(accb0[j], ones, bg1[j]);
            }
        }
        cp_async_wait_all();
        __syncwarp();
    }

    // ---- per-warp sums -> shared [NW][P] -> fixed-order sum over the warps -> this CTA's partial row ----
    __syncthreads();
    double *red = scratch;                                  // NW * P doubles (fits: P <= 1024 on this path)
    double *mine = red + (size_t)w * p.P;
    for (int i = lane; i < p.P; i += 32) mine[i] = 0.0;
    __syncwarp();
#pragma unroll
    for (int m = 0; m < MT0; ++m)
#pragma unroll
        for (int j = 0; j < NT1; ++j)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = 8 * m + g, col = 8 * j + 2 * t + r;
                if (row < rows0 && col < L1) mine[p.w_off0 + row * L1 + col] = acc0[m][j][r];
            }
#pragma unroll
    for (int j = 0; j < NT1; ++j)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int col = 8 * j + 2 * t + r;
            if (!free_col && g == 0 && col < L1) mine[p.w_off0 + L0 * L1 + col] = accb0[j][r];
        }
#pragma unroll
    for (int m = 0; m < NT1; ++m)
#pragma unroll
        for (int j = 0; j < NT2; ++j)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = 8 * m + g, col = 8 * j + 2 * t + r;
                if (row < L1 && col < L2) mine[p.w_off1 + row * L2 + col] = acc1[m][j][r];
            }
#pragma unroll
    for (int j = 0; j < NT2; ++j)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int col = 8 * j + 2 * t + r;
            if (g == 0 && col < L2) mine[p.w_off1 + L1 * L2 + col] = accb1[j][r];
        }
#pragma unroll
    for (int m = 0; m < NT2; ++m)
#pragma unroll
        for (int c = 0; c < NT3; ++c)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = 8 * m + g, col = 8 * c + 2 * t + r;
                if (row < L2 && col < L3) mine[p.w_off2 + row * L3 + col] = acc2[m][c][r];
            }
#pragma unroll
    for (int c = 0; c < NT3; ++c)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int col = 8 * c + 2 * t + r;
            if (g == 0 && col < L3) mine[p.w_off2 + L2 * L3 + col] = accb2[c][r];
        }
    __syncthreads();
    double *out = p.partial + (size_t)blockIdx.x * p.P;
    for (int i = tid; i < p.P; i += NT) {
        double sum = red[i];
#pragma unroll
        for (int ww = 1; ww < NW; ++ww) sum += red[(size_t)ww * p.P + i];
        out[i] = sum;
    }
}

template <typename C, char ACT1, char ACT2>
__global__ void __launch_bounds__(C::NTHREADS, 1) k_fvp_warp(const FusedArgs p) {
    if (p.done && *p.done) return;
    extern __shared__ __align__(16) double sm[];
    warp_pass<C, ACT1, ACT2>(p, sm, 3);
}

// ---------------------------------------------------------------------------------------------------------------
// The whole conjugate-gradient solve (TRPO_CG.c:25-107) as ONE persistent cooperative kernel: one CTA per SM, the model
// staged once, per iteration the FVP pass above (direction = the current p, re-staged from global memory), then
//   grid barrier -> fixed-order sum of the per-CTA partial rows, every CTA a slice of columns
//                -> multi-GPU: the slice is pushed into every peer's memory over NVLink and the peers' slices are summed
//                   in rank order as they arrive (per-CTA flags: CTA b only waits for CTA b of each peer)
//                -> z = sum / N + damping p, partial p.z        -> grid barrier -> v = r.r / p.z, x += v p, r -= v z
//                -> partial r.r, x.x                            -> grid barrier -> mu, p = r + mu p -> grid barrier.
// Every CTA forms the global scalars itself from the per-CTA partials in the same fixed order, so all CTAs (and, with the
// fixed rank order of the exchange, all GPUs) take identical decisions and end with bitwise identical state: no broadcast.
// Round 1 ran 3 - 4 launches per iteration (FVP, row reduction, all-reduce, single-CTA update): 44 us of serial tail per
// iteration at 8 GPUs, a third of a 50 k-state solve of the arm policy.
struct SolveArgs {
    const double *b;             // right-hand side
    double *x, *r, *pv, *z;      // CG vectors in global memory (P each); pv is the direction the FVP pass stages
    double *zsum;                // reduced, un-normalised FVP sum (scratch, P)
    double *dots;                // [4][DOT_STRIDE] per-CTA partials of b.b / p.z / r.r / x.x
    unsigned int *gbar;          // grid barrier: [0] arrivals, [1] generation
    CgState *st;
    double *trace;
    int trace_cap, max_iter, rows, logstd_off;
    double residual_th, damping, n_total;
    P2PComm comm;                // world <= 1: single GPU
    unsigned long long *timeline;   // optional [max_iter][8] globaltimer stamps of CTA 0 (nullptr: off), see trpo_ctx_solve_timeline
};
constexpr int DOT_STRIDE = 160;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// fixed-order block sum of one value per thread (NT threads, NT / 32 <= 32 warps); result in every thread
__device__ __forceinline__ double solve_block_sum(double v, double *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (w == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
// All CTAs of the (co-resident: cooperative launch) grid. One release-acquire chain, no stand-alone fences: the block barrier
// orders every thread's writes before thread 0's acq_rel arrival, the last arriver's release store of the new generation
// publishes them, the acquire loads of the spinning CTAs (followed by their block barrier) make them visible. 2.5 - 3.5 us on
// 148 SMs. (A flag-per-CTA variant in which every CTA polls all 148 flags and picks up the partial sums in the same round
// trip measured SLOWER -- 4.6 - 7.5 us: 148 x 148 polling acquire loads -- and was dropped, profiles/r02_summary.md.)
__device__ __forceinline__ void grid_sync(unsigned int *bar, unsigned int nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int gen, prev;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");    // cannot change before we arrive
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(bar) : "memory");
        if (prev == nblocks - 1) {
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(bar), "r"(0u) : "memory");
            st_release_gpu_u32(bar + 1, gen + 1);
        } else {
            const long long t0 = clock64();
            while (ld_acquire_gpu_u32(bar + 1) == gen) {
                if (clock64() - t0 > 20000000000LL) __trap();      // ~10 s: a CTA is missing (not co-resident?): abort, do not hang
            }
        }
    }
    __syncthreads();
}
// sum of the per-CTA partials of one dot product, identical in every thread of every CTA
__device__ __forceinline__ double solve_global_sum(const double *part, int n, double *red) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += __ldcg(&part[i]);
    return solve_block_sum(v, red);
}

template <typename C, char ACT1, char ACT2, bool WARP>
__global__ void __launch_bounds__(C::NTHREADS, 1) k_cg_solve(FusedArgs p, const SolveArgs s) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[40];
    __shared__ double colsh[8][32];
    const int tid = threadIdx.x, NT = C::NTHREADS, G = gridDim.x, P = p.P;
    const int cpb = (P + G - 1) / G, lo = blockIdx.x * cpb, hi = min(P, lo + cpb);
    const bool multi = s.comm.world > 1;
    unsigned long long seq = multi ? *s.comm.seq_dev : 0ull;
    // ---- x = 0, r = p = b, r.r (TRPO_CG.c:25-42) ----
    double acc = 0.0;
    for (int e = lo + tid; e < hi; e += NT) {
        const double bi = s.b[e];
        s.x[e] = 0.0; s.r[e] = bi; s.pv[e] = bi;
        acc += bi * bi;
    }
    acc = solve_block_sum(acc, red);
    if (tid == 0) s.dots[0 * DOT_STRIDE + blockIdx.x] = acc;
    grid_sync(s.gbar, G);
    double rdotr = solve_global_sum(s.dots + 0 * DOT_STRIDE, G, red);
    if (blockIdx.x == 0 && tid == 0) { s.trace[0] = rdotr; s.trace[s.trace_cap] = 0.0; }
    int iters = 0, done = rdotr < s.residual_th ? 1 : 0;
    double pdotz = 0.0, xnorm = 0.0;
    p.v = s.pv;
    int stage = 3;
    for (int it = 0; it < s.max_iter && !done; ++it) {
        // ---- FVP pass: this CTA's partial row of sum_n F_n p ----
        const bool stamp = s.timeline != nullptr && blockIdx.x == 0 && tid == 0;
        unsigned long long *tl = s.timeline + (size_t)it * 8;
        if (stamp) tl[0] = globaltimer_ns();
        if (WARP) { pass_grid_wait(p); warp_pass<C, ACT1, ACT2>(p, sm, stage); }
        else fused_pass<C, ACT1, ACT2, false>(p, sm, stage);
        stage = 2;                                            // the model stays staged; only the direction changes
        if (stamp) tl[1] = globaltimer_ns();                  // this CTA's pass done
        grid_sync(s.gbar, G);
        if (stamp) tl[2] = globaltimer_ns();                  // every CTA's pass done
        // ---- column sums of this CTA's slice (fixed row order), exchange, z and p.z ----
        const unsigned long long nseq = seq + 1;
        const size_t slot = multi ? ((size_t)(nseq & 1) * s.comm.world + s.comm.rank) * P : 0;
        for (int c0 = lo; c0 < hi; c0 += 32) {
            const int e = c0 + (tid & 31), ry = tid >> 5;
            const bool valid = e < hi && e < s.logstd_off;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            if (valid && ry < 8) {
                // all of this thread's rows in flight at once (the loads are independent; four at a time cost five L2 round
                // trips per chunk), summed afterwards in the order of k_reduce_partials
                for (int rb = ry; rb < s.rows; rb += 8 * 20) {
                    double v[20];
#pragma unroll
                    for (int k = 0; k < 20; ++k) v[k] = rb + 8 * k < s.rows ? __ldcg(&p.partial[(size_t)(rb + 8 * k) * P + e]) : 0.0;
#pragma unroll
                    for (int k = 0; k < 20; k += 4) { s0 += v[k]; s1 += v[k + 1]; s2 += v[k + 2]; s3 += v[k + 3]; }
                }
            }
            __syncthreads();
            if (ry < 8) colsh[ry][tid & 31] = (s0 + s1) + (s2 + s3);
            __syncthreads();
            if (ry == 0) {
                double tsum = colsh[0][tid];
#pragma unroll
                for (int k = 1; k < 8; ++k) tsum += colsh[k][tid];
                if (valid) s.zsum[e] = tsum;
                colsh[0][tid] = tsum;
            }
            if (multi) {
                // low-latency exchange (the NCCL "LL" idea): every 8-byte word carries 4 bytes of payload and the 32-bit sequence
                // number, so the receiver needs no flag and the sender no fence -- a word is valid when its tag matches. Warp w
                // stores this chunk's sums into rank w's memory (two tagged words per double, one 16-byte NVLink store).
                __syncthreads();
                if (ry < s.comm.world && valid) {
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(colsh[0][tid & 31]);
                    const unsigned long long tag = (nseq & 0xffffffffull) << 32;
                    unsigned long long *dst = s.comm.ll[ry] + 2 * (slot + e);
                    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" :: "l"(dst), "l"(tag | (bits & 0xffffffffull)), "l"(tag | (bits >> 32)) : "memory");
                }
            }
        }
        __syncthreads();                                      // zsum of this slice was written by warp 0, every warp reads it below
        if (stamp) tl[3] = globaltimer_ns();                  // slice column sums formed (and pushed)
        if (multi) {
            // receive: thread (rank r, column c) of each 32-column chunk spins on its two tagged words, then the ranks are summed
            // in fixed order (bounded spin: a peer that never arrives poisons the solve instead of hanging the GPU)
            const unsigned long long want = nseq & 0xffffffffull;
            const unsigned long long *mine = s.comm.ll[s.comm.rank] + 2 * ((size_t)(nseq & 1) * s.comm.world * P);
            for (int c0 = lo; c0 < hi; c0 += 32) {
                const int e = c0 + (tid & 31), rk = tid >> 5;
                const bool valid = e < hi && e < s.logstd_off;
                double val = 0.0;
                if (valid && rk < s.comm.world) {
                    const unsigned long long *src = mine + 2 * ((size_t)rk * P + e);
                    unsigned long long w0, w1;
                    const long long t0 = clock64();
                    for (;;) {
                        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
                        if ((w0 >> 32) == want && (w1 >> 32) == want) break;
                        if (clock64() - t0 > 40000000000LL) { *(volatile int *)s.comm.error = 1; break; }
                    }
                    val = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
                }
                __syncthreads();
                if (rk < 8) colsh[rk][tid & 31] = val;
                __syncthreads();
                if (rk == 0 && valid) {
                    double t2 = colsh[0][tid];
                    for (int r2 = 1; r2 < s.comm.world; ++r2) t2 += colsh[r2][tid];
                    s.zsum[e] = t2;
                }
            }
            __syncthreads();
            seq = nseq;                                       // a timed-out wait is handled after the next grid barrier
        }
        if (stamp) tl[4] = globaltimer_ns();                  // peers' slices have arrived
        acc = 0.0;
        for (int e = lo + tid; e < hi; e += NT) {
            const double pi = __ldcg(&s.pv[e]);
            double mean;
            if (e >= s.logstd_off) mean = 2.0 * pi;          // LogStd block: sum_n 2 v / N exactly (TRPO_FVP.c:918-921)
            else mean = s.zsum[e] / s.n_total;          // single GPU: this CTA's column sums; several: the sum over the ranks
            const double zi = mean + s.damping * pi;
            s.z[e] = zi;
            acc += pi * zi;
        }
        acc = solve_block_sum(acc, red);
        if (tid == 0) s.dots[1 * DOT_STRIDE + blockIdx.x] = acc;
        grid_sync(s.gbar, G);
        // a peer never arrived (flag set before the barrier, so every CTA sees it now): stop, poison, the host call fails
        if (multi && *(volatile int *)s.comm.error) { done = 2; break; }
        if (stamp) tl[5] = globaltimer_ns();
        pdotz = solve_global_sum(s.dots + 1 * DOT_STRIDE, G, red);
        const double v = rdotr / pdotz;
        double acc_r = 0.0, acc_x = 0.0;
        for (int e = lo + tid; e < hi; e += NT) {
            const double xi = s.x[e] + v * s.pv[e];
            const double ri = s.r[e] - v * s.z[e];
            s.x[e] = xi; s.r[e] = ri;
            acc_r += ri * ri;
            acc_x += xi * xi;
        }
        acc_r = solve_block_sum(acc_r, red);
        acc_x = solve_block_sum(acc_x, red);
        if (tid == 0) { s.dots[2 * DOT_STRIDE + blockIdx.x] = acc_r; s.dots[3 * DOT_STRIDE + blockIdx.x] = acc_x; }
        grid_sync(s.gbar, G);
        if (stamp) tl[6] = globaltimer_ns();
        const double newrdotr = solve_global_sum(s.dots + 2 * DOT_STRIDE, G, red);
        const double xx = solve_global_sum(s.dots + 3 * DOT_STRIDE, G, red);
        const double mu = newrdotr / rdotr;
        for (int e = lo + tid; e < hi; e += NT) s.pv[e] = s.r[e] + mu * s.pv[e];
        rdotr = newrdotr;
        xnorm = sqrt(xx);
        iters = it + 1;
        if (blockIdx.x == 0 && tid == 0 && iters < s.trace_cap) { s.trace[iters] = newrdotr; s.trace[s.trace_cap + iters] = xnorm; }
        if (newrdotr < s.residual_th) done = 1;
        // the new direction must be complete before anyone stages it: arrive here, wait inside the next pass (pass_grid_wait)
        {
            __syncthreads();
            unsigned int gen = 0;
            if (tid == 0) {
                unsigned int prev;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(s.gbar + 1) : "memory");
                asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(s.gbar) : "memory");
                if (prev == (unsigned int)G - 1) {
                    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(s.gbar), "r"(0u) : "memory");
                    st_release_gpu_u32(s.gbar + 1, gen + 1);
                }
            }
            p.wait_bar = s.gbar;
            p.wait_gen = gen;                                 // meaningful in thread 0, the only reader
        }
        if (stamp) tl[7] = globaltimer_ns();
    }
    if (blockIdx.x == 0 && tid == 0) {
        s.st->rdotr = done == 2 ? __longlong_as_double(0x7ff8000000000000LL) : rdotr;
        s.st->pdotz = pdotz; s.st->xnorm = xnorm; s.st->iters = iters; s.st->done = done ? 1 : 0;
        if (multi) *s.comm.seq_dev = seq;
    }
}

template <typename C>
constexpr size_t warp_smem_bytes() {
    constexpr int WS = 2 * (8 * C::RS0 + 8) + 8 * C::RS1 + 8 * C::RS2 + 8 * C::RSB + 8 * C::RS3;
    constexpr int PMAX = (C::K0 + 1) * C::H1 + (C::H1 + 1) * C::H2 + (C::H2 + 1) * C::AP + C::AP;
    constexpr int SCR = C::NW * (WS > PMAX ? WS : PMAX);
    return sizeof(double) * (C::oY0 + 64 + SCR);
}

// ---------------------------------------------------------------------------------------------------------------
constexpr int FUSED_SMS = 148;
constexpr int FUSED_MAX_ROWS = 4 * FUSED_SMS;

// armDOF_0 (15-16-16-3). Measured at 50 k states (profiles/r01_summary.md): one 8-warp CTA per SM gives a 30.7 us kernel +
// a 148-row partial reduction; four 4-warp CTAs per SM give 26.0 us but 592 partial rows, and lose end to end
// (0.57 vs 0.51 ms per 10-iteration solve). The small-CTA variant stays selectable with TRPO_FUSED_ARM_VARIANT=4.
using CfgArm  = Cfg<16, 16, 16, 8, 8, 1>;
using CfgArm4 = Cfg<16, 16, 16, 8, 4, 4>;
using CfgH32  = Cfg<32, 32, 32, 8, 8>;     // hidden <= 32, up to 32 inputs (e.g. 11-32-32-3)
using CfgP64  = Cfg<4, 64, 64, 8, 8>;      // InvertedPendulum-size: 4-64-64-1
using CfgM64  = Cfg<20, 64, 64, 8, 8>;     // 17-64-64-6 (and anything with L0 <= 20, hidden <= 64, A <= 8)
// two independent 4-warp groups per CTA (see Cfg): tiles of 32 samples, accumulators parked in Tensor Memory
using CfgH32g = Cfg<32, 32, 32, 8, 8, 1, 2>;
using CfgP64g = Cfg<4, 64, 64, 8, 8, 1, 2>;
using CfgM64g = Cfg<20, 64, 64, 8, 8, 1, 2>;

enum FusedShape { SHAPE_NONE = 0, SHAPE_ARM, SHAPE_H32, SHAPE_P64, SHAPE_M64 };

FusedShape pick_shape(const NetDesc &net) {
    if (net.K != 3) return SHAPE_NONE;
    if (net.ac[3] != 'l' && net.ac[3] != 'o') return SHAPE_NONE;
    const int L0 = net.L[0], L1 = net.L[1], L2 = net.L[2], L3 = net.L[3];
    if (L3 > 8) return SHAPE_NONE;
    if (L0 <= 16 && L1 <= 16 && L2 <= 16) return SHAPE_ARM;
    if (L0 <= 32 && L1 <= 32 && L2 <= 32) return SHAPE_H32;
    if (L0 <= 4 && L1 <= 64 && L2 <= 64) return SHAPE_P64;
    if (L0 <= 20 && L1 <= 64 && L2 <= 64) return SHAPE_M64;
    return SHAPE_NONE;
}

// Warp groups per CTA of the block-wide kernel. Measured at 1 M states (profiles/r02_summary.md): 17-32-32-6 0.963 -> 0.830 ms
// with two groups, 17-64-64-6 2.082 -> 2.079 and 4-64-64-1 1.797 -> 1.794 (no gain: there every warp already owns a full row of
// outer-product tiles and the kernel is bound by the LDS-fed DMMA stream itself). Default: two groups for hidden widths <= 32,
// one for the 64-wide shapes (half as many partial rows to reduce); TRPO_FUSED_GROUPS = 1 | 2 overrides (read once).
int fused_groups(int dflt) {
    static const int forced = getenv("TRPO_FUSED_GROUPS") ? atoi(getenv("TRPO_FUSED_GROUPS")) : 0;
    return forced == 1 || forced == 2 ? forced : dflt;
}

template <typename C, char A1, char A2, bool PG>
int launch_cfg(const FusedArgs &a, int grid, cudaStream_t st) {
    static DeviceOnce configured;
    if (configured.pending()) {
        if (cudaFuncSetAttribute(k_fvp_fused<C, A1, A2, PG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES) != cudaSuccess)
            return -1;
        configured.mark();
    }
    k_fvp_fused<C, A1, A2, PG><<<grid, C::NTHREADS, C::SMEM_BYTES, st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template <typename C, char A1, char A2>
int launch_warp_cfg(const FusedArgs &a, int grid, cudaStream_t st) {
    static DeviceOnce configured;
    if (configured.pending()) {
        if (cudaFuncSetAttribute(k_fvp_warp<C, A1, A2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)warp_smem_bytes<C>()) != cudaSuccess)
            return -1;
        configured.mark();
    }
    k_fvp_warp<C, A1, A2><<<grid, C::NTHREADS, warp_smem_bytes<C>(), st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// warp-private kernel: one 8-warp CTA per SM, every warp works through its own 8-sample units
template <typename C>
int launch_warp_shape(const FusedArgs &a, cudaStream_t st, int *rows) {
    const long long nunits = (a.nsamples + 7) / 8;
    const long long want = (nunits + C::NW - 1) / C::NW;
    // one 8-warp CTA per SM: a 128-register build with two CTAs per SM measured 7 % slower at 50 k states and 2 % slower at 1 M
    // (profiles/r02_summary.md) -- the kernel is not short of warps
    const int grid = (int)(want < FUSED_SMS ? want : FUSED_SMS);
    *rows = grid;
    if (a.act1 == 't' && a.act2 == 't') return launch_warp_cfg<C, 't', 't'>(a, grid, st);
    return launch_warp_cfg<C, 0, 0>(a, grid, st);
}

template <typename C, bool PG = false>
int launch_shape(const FusedArgs &a, cudaStream_t st, int *rows) {
    const long long ntiles = (a.nsamples + C::SG - 1) / C::SG, nctas = (ntiles + C::NG - 1) / C::NG;
    const int max_grid = FUSED_SMS * C::CTAS_PER_SM;
    const int grid = (int)(nctas < max_grid ? nctas : max_grid);
    *rows = grid * C::NG;                                  // every warp group writes its own partial row
    if (a.act1 == 't' && a.act2 == 't') return launch_cfg<C, 't', 't', PG>(a, grid, st);
    return launch_cfg<C, 0, 0, PG>(a, grid, st);
}

}  // namespace

bool fused_eligible(const NetDesc &net) { return pick_shape(net) != SHAPE_NONE; }
int fused_partial_rows() { return FUSED_MAX_ROWS; }

int fused_fvp_accumulate(const NetDesc &net, const double *d_theta, const double *d_v, const double *d_inv_var,
                         const double *d_obs, size_t nsamples, double *d_partial, double *d_zsum,
                         const int *d_done, const P2PComm *p2p, const int *stream_ready, size_t stream_chunk,
                         int *stream_error, cudaStream_t st, long long *launches) {
    const FusedShape shape = pick_shape(net);
    if (shape == SHAPE_NONE) return 1;
    FusedArgs a;
    a.theta = d_theta; a.v = d_v; a.inv_var = d_inv_var; a.obs = d_obs; a.partial = d_partial; a.done = d_done;
    a.nsamples = (long long)nsamples;
    a.L0 = net.L[0]; a.L1 = net.L[1]; a.L2 = net.L[2]; a.L3 = net.L[3];
    a.w_off0 = net.w_off[0]; a.w_off1 = net.w_off[1]; a.w_off2 = net.w_off[2];
    a.P = net.P;
    a.act1 = net.ac[1]; a.act2 = net.ac[2]; a.act3 = net.ac[3];
    a.ready = stream_ready; a.chunk_samples = (long long)(stream_chunk ? stream_chunk : 1); a.error = stream_error;
    a.mean = a.action = a.adv = nullptr; a.logstd_off = net.logstd_off; a.mean_out = nullptr;
    a.wait_bar = nullptr; a.wait_gen = 0;
    int rows = 0, rc = -1;
    switch (shape) {
        case SHAPE_ARM: {
            // default: warp-private kernel (no block barriers); TRPO_FUSED_ARM_VARIANT=8 / 4 select the block-wide variants
            static const int variant = getenv("TRPO_FUSED_ARM_VARIANT") ? atoi(getenv("TRPO_FUSED_ARM_VARIANT")) : 0;
            rc = variant == 4 ? launch_shape<CfgArm4>(a, st, &rows)
               : variant == 8 ? launch_shape<CfgArm>(a, st, &rows)
                              : launch_warp_shape<CfgArm>(a, st, &rows);
            break;
        }
        case SHAPE_H32: rc = fused_groups(2) == 2 ? launch_shape<CfgH32g>(a, st, &rows) : launch_shape<CfgH32>(a, st, &rows); break;
        case SHAPE_P64: rc = fused_groups(1) == 2 ? launch_shape<CfgP64g>(a, st, &rows) : launch_shape<CfgP64>(a, st, &rows); break;
        case SHAPE_M64: rc = fused_groups(1) == 2 ? launch_shape<CfgM64g>(a, st, &rows) : launch_shape<CfgM64>(a, st, &rows); break;
        default: return 1;
    }
    if (rc) return -1;
    ++*launches;
    launch_reduce_partials(d_partial, rows, net.P, d_zsum, d_done, p2p, st, launches);
    return 0;
}

namespace {
template <typename C, char A1, char A2, bool WARP>
int launch_solve_cfg(const FusedArgs &a, const SolveArgs &sa, int grid, cudaStream_t st) {
    const size_t smem = WARP ? warp_smem_bytes<C>() : C::SMEM_BYTES;
    static DeviceOnce configured;
    if (configured.pending()) {
        if (cudaFuncSetAttribute(k_cg_solve<C, A1, A2, WARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -1;
        configured.mark();
    }
    void *args[] = {(void *)&a, (void *)&sa};
    // cooperative launch: the grid barriers need every CTA resident (one CTA per SM, grid <= number of SMs)
    return cudaLaunchCooperativeKernel((const void *)k_cg_solve<C, A1, A2, WARP>, dim3(grid), dim3(C::NTHREADS), args, smem, st) == cudaSuccess ? 0 : -1;
}
template <typename C, bool WARP>
int launch_solve_shape(FusedArgs &a, SolveArgs &sa, cudaStream_t st) {
    int grid;
    if (WARP) {
        const long long nunits = (a.nsamples + 7) / 8, want = (nunits + C::NW - 1) / C::NW;
        grid = (int)(want < FUSED_SMS ? want : FUSED_SMS);
    } else {
        const long long ntiles = (a.nsamples + C::SG - 1) / C::SG;
        grid = (int)(ntiles < FUSED_SMS ? ntiles : FUSED_SMS);
    }
    sa.rows = grid;
    if (a.act1 == 't' && a.act2 == 't') return launch_solve_cfg<C, 't', 't', WARP>(a, sa, grid, st);
    return launch_solve_cfg<C, 0, 0, WARP>(a, sa, grid, st);
}
}  // namespace

int fused_cg_solve(const NetDesc &net, const double *d_theta, const double *d_inv_var, const double *d_obs, size_t nsamples,
                   double n_total, double *d_partial, const double *d_b, double *d_x, double *d_r, double *d_p, double *d_z,
                   double *d_zsum, double *d_dots, unsigned int *d_gbar, CgState *d_state, double *d_trace, int trace_cap,
                   size_t max_iter, double residual_th, double damping, const P2PComm *p2p, const int *stream_ready,
                   size_t stream_chunk, int *stream_error, unsigned long long *d_timeline, cudaStream_t st, long long *launches) {
    const FusedShape shape = pick_shape(net);
    if (shape == SHAPE_NONE) return 1;
    static const bool disabled = getenv("TRPO_NO_FUSED_SOLVE") != nullptr;
    static const bool forced = getenv("TRPO_FUSED_SOLVE") != nullptr;
    if (disabled || max_iter > 0x7fffffff) return 1;
    // Single GPU, tiny batch (armDOF_0 x 50 k: a pass is 20 us): the 4 grid synchronisations per iteration (13 us) cost more than
    // the three short launches of the per-iteration path replayed from a CUDA graph (0.355 against 0.31 ms per solve, measured);
    // the persistent kernel wins from about 2 GFLOP per pass on, and always when the FVP sums cross GPUs.
    {
        double flops = 6.0 * net.L[0] * net.L[1];
        for (int i = 1; i < net.K; ++i) flops += 10.0 * net.L[i] * net.L[i + 1];
        if (!forced && !(p2p && p2p->world > 1) && flops * (double)nsamples < 2e9) return 1;
        // 4-64-64-1 (one input k-step, one action): on one GPU the pass compiled into the solve kernel is 2.4 % slower than the
        // stand-alone kernel replayed from a graph (18.44 against 18.01 ms per 1 M-state solve, tools/time_solve.py); 17-64-64-6 gains 1.5 %
        if (!forced && !(p2p && p2p->world > 1) && shape == SHAPE_P64) return 1;
    }
    FusedArgs a;
    a.theta = d_theta; a.v = d_p; a.inv_var = d_inv_var; a.obs = d_obs; a.partial = d_partial; a.done = nullptr;
    a.nsamples = (long long)nsamples;
    a.L0 = net.L[0]; a.L1 = net.L[1]; a.L2 = net.L[2]; a.L3 = net.L[3];
    a.w_off0 = net.w_off[0]; a.w_off1 = net.w_off[1]; a.w_off2 = net.w_off[2];
    a.P = net.P;
    a.act1 = net.ac[1]; a.act2 = net.ac[2]; a.act3 = net.ac[3];
    a.ready = stream_ready; a.chunk_samples = (long long)(stream_chunk ? stream_chunk : 1); a.error = stream_error;
    a.mean = a.action = a.adv = nullptr; a.logstd_off = net.logstd_off; a.mean_out = nullptr;
    a.wait_bar = nullptr; a.wait_gen = 0;
    SolveArgs sa;
    sa.b = d_b; sa.x = d_x; sa.r = d_r; sa.pv = d_p; sa.z = d_z; sa.zsum = d_zsum; sa.dots = d_dots; sa.gbar = d_gbar;
    sa.st = d_state; sa.trace = d_trace; sa.trace_cap = trace_cap; sa.max_iter = (int)max_iter; sa.rows = 0;
    sa.logstd_off = net.logstd_off; sa.residual_th = residual_th; sa.damping = damping; sa.n_total = n_total;
    if (p2p && p2p->world > 1) sa.comm = *p2p; else { sa.comm = P2PComm{}; }
    sa.timeline = d_timeline;
    int rc;
    switch (shape) {
        case SHAPE_ARM: rc = launch_solve_shape<CfgArm, true>(a, sa, st); break;
        case SHAPE_H32: rc = launch_solve_shape<CfgH32, false>(a, sa, st); break;
        case SHAPE_P64: rc = launch_solve_shape<CfgP64, false>(a, sa, st); break;
        case SHAPE_M64: rc = launch_solve_shape<CfgM64, false>(a, sa, st); break;
        default: return 1;
    }
    if (rc) return -1;
    ++*launches;
    return 0;
}

// Policy-gradient sum (TRPO_Update.c:254-371) on the fused kernel: zsum = sum_n [GW, GB, ..., GLogStd].
// d_inv_std = exp(-LogStd) of the current parameters.
int fused_pg_accumulate(const NetDesc &net, const double *d_theta, const double *d_inv_std, const double *d_obs,
                        const double *d_mean, const double *d_action, const double *d_adv, size_t nsamples,
                        double *d_partial, double *d_zsum, double *d_mean_out, cudaStream_t st, long long *launches) {
    const FusedShape shape = pick_shape(net);
    if (shape == SHAPE_NONE) return 1;
    FusedArgs a;
    a.theta = d_theta; a.v = nullptr; a.inv_var = d_inv_std; a.obs = d_obs; a.partial = d_partial; a.done = nullptr;
    a.nsamples = (long long)nsamples;
    a.L0 = net.L[0]; a.L1 = net.L[1]; a.L2 = net.L[2]; a.L3 = net.L[3];
    a.w_off0 = net.w_off[0]; a.w_off1 = net.w_off[1]; a.w_off2 = net.w_off[2];
    a.P = net.P;
    a.act1 = net.ac[1]; a.act2 = net.ac[2]; a.act3 = net.ac[3];
    a.ready = nullptr; a.chunk_samples = 1; a.error = nullptr;
    a.mean = d_mean; a.action = d_action; a.adv = d_adv; a.logstd_off = net.logstd_off; a.mean_out = d_mean_out;
    a.wait_bar = nullptr; a.wait_gen = 0;
    int rows = 0, rc = -1;
    switch (shape) {
        case SHAPE_ARM: rc = launch_shape<CfgArm, true>(a, st, &rows); break;
        case SHAPE_H32: rc = fused_groups(2) == 2 ? launch_shape<CfgH32g, true>(a, st, &rows) : launch_shape<CfgH32, true>(a, st, &rows); break;
        case SHAPE_P64: rc = fused_groups(1) == 2 ? launch_shape<CfgP64g, true>(a, st, &rows) : launch_shape<CfgP64, true>(a, st, &rows); break;
        case SHAPE_M64: rc = fused_groups(1) == 2 ? launch_shape<CfgM64g, true>(a, st, &rows) : launch_shape<CfgM64, true>(a, st, &rows); break;
        default: return 1;
    }
    if (rc) return -1;
    ++*launches;
    launch_reduce_partials(d_partial, rows, net.P, d_zsum, nullptr, nullptr, st, launches);
    return 0;
}
