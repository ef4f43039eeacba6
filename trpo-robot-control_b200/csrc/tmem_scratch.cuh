// tmem_scratch.cuh -- Tensor Memory (TMEM, 256 KB per SM on sm_100a) used as a per-thread register parking area.
//
// The FP64 kernels cannot use tcgen05.mma (no f64 kind), but TMEM itself is reachable with tcgen05.st / tcgen05.ld (SASS
// STTM / LDTM): shape 32x32b moves register j of lane i of a warp to/from TMEM lane (quadrant base + i), column (base + j) --
// a thread gets back exactly what it stored, so the layout needs no thought. A warp may only touch the 32-lane quadrant
// (warp_id % 4); warps w and w + 4 therefore use different column ranges of the same quadrant.
#pragma once
#include <stdint.h>

// one warp allocates `cols` columns (power of two >= 32) for the CTA and publishes the base address through shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot, uint32_t cols) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// address of this warp's quadrant at column `col` of an allocation
__device__ __forceinline__ uint32_t tmem_warp_addr(uint32_t base, int warp, int col) {
    return base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col;
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8]; tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16]; tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32]; tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
}

// park / fetch N doubles (N a multiple of 4) of a thread at column `col` of its warp's quadrant. Stores are asynchronous:
// follow a group of them with tmem_wait_st() before the data is fetched again. A load and its tcgen05.wait::ld sit in ONE asm
// statement, so the compiler cannot schedule a use of the destination registers before the wait.
template <int N>
__device__ __forceinline__ void tmem_park(uint32_t taddr, const double (&v)[N]) {
    static_assert(N % 4 == 0, "pad to a multiple of 4 doubles");
    int done = 0;
#pragma unroll
    for (int chunk = 16; chunk >= 4; chunk >>= 1) {           // 16 / 8 / 4 doubles = x32 / x16 / x8
#pragma unroll
        for (; done + chunk <= N; done += chunk) {
            if (chunk == 16) {
                uint32_t r[32];
#pragma unroll
                for (int i = 0; i < 16; ++i) { r[2 * i] = (uint32_t)__double2loint(v[done + i]); r[2 * i + 1] = (uint32_t)__double2hiint(v[done + i]); }
                tmem_st32(taddr + 2 * done, r);
            } else if (chunk == 8) {
                uint32_t r[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) { r[2 * i] = (uint32_t)__double2loint(v[done + i]); r[2 * i + 1] = (uint32_t)__double2hiint(v[done + i]); }
                tmem_st16(taddr + 2 * done, r);
            } else {
                uint32_t r[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) { r[2 * i] = (uint32_t)__double2loint(v[done + i]); r[2 * i + 1] = (uint32_t)__double2hiint(v[done + i]); }
                tmem_st8(taddr + 2 * done, r);
            }
        }
    }
}
template <int N>
__device__ __forceinline__ void tmem_fetch(uint32_t taddr, double (&v)[N]) {
    static_assert(N % 4 == 0, "pad to a multiple of 4 doubles");
    int done = 0;
#pragma unroll
    for (int chunk = 16; chunk >= 4; chunk >>= 1) {
#pragma unroll
        for (; done + chunk <= N; done += chunk) {
            if (chunk == 16) {
                uint32_t r[32];
                tmem_ld32(taddr + 2 * done, r);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[done + i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
            } else if (chunk == 8) {
                uint32_t r[16];
                tmem_ld16(taddr + 2 * done, r);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[done + i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
            } else {
                uint32_t r[8];
                tmem_ld8(taddr + 2 * done, r);
#pragma unroll
                for (int i = 0; i < 4; ++i) v[done + i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
            }
        }
    }
}
