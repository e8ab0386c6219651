// tcgen05 (5th-gen tensor core) + TMEM primitives for sm_100a, inline PTX.
//
// Operand convention used throughout this library ("swap-AB" for short token counts):
//   D^T[128 features x N tokens] (TMEM, fp32)  +=  W[128 x 64] (smem A operand, K-major)
//                                                  * X[N x 64]^T (smem B operand, K-major)
// Both operands live in shared memory in the canonical 128-byte-swizzled K-major layout:
//   row r, 16-byte chunk c (8 fp16 along K) at byte  r*128 + ((c ^ (r & 7)) << 4),
//   8-row groups 1024 B apart (SBO), one 64-element K block per [rows x 128 B] slab.
// One tcgen05.mma consumes K = 16 (32 bytes): the K-step inside a 64-wide block is selected by
// advancing the descriptor start address by 32 B.
#pragma once
#include "common.cuh"

namespace sfb {

// ---- shared-memory matrix descriptor (SWIZZLE_128B, K-major) -------------------------------
// bits [0,14) start>>4 | [16,30) LBO>>4 (ignored for swizzled K-major, set 1) | [32,46) SBO>>4
// | [46,48) version = 1 (sm_100) | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// ---- instruction descriptor: kind::f16, A = B = fp16 (K-major), D = fp32, M x N ---------------
// bits [4,6) D format (1 = f32) | [7,10) A format (0 = f16) | [10,13) B format | bit 15/16 A/B major
// (0 = K) | [17,23) N >> 3 | [24,29) M >> 4
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

// one lane of the (converged) warp; the surrounding code stays warp-uniform so that ptxas keeps the
// descriptors in uniform registers (a lane-0 branch makes it emit an R2UR waterfall per MMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp executes these) -----------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32-bit, 8 consecutive columns per thread ------------------
// A warp may only touch the TMEM lanes of its quadrant: lanes 32*(warp_id % 4) .. +31.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// ---- TMEM -> registers, 16 lanes x 256 bit: the m16n8 accumulator-fragment distribution -----------
// Lanes [lane0, lane0+16) of the warp's quadrant x 8 (x1) or 16 (x2) consecutive columns.  Thread T holds
//   v[0], v[1] = (lane0 + T/4,     col + 2(T%4) + {0,1})      v[2], v[3] = (lane0 + T/4 + 8, same columns)
// and, for .x2, v[4..7] the same rows at columns + 8.  With features on TMEM lanes and tokens on columns this is
// exactly the fragment stmatrix.trans stores as token rows of 8 consecutive features.
__device__ __forceinline__ void tmem_ld_16x256b_x1(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// four / two 8x8 b16 matrices, transposed fragments -> rows of 16 bytes; lanes 8i..8i+7 give matrix i's row addresses
__device__ __forceinline__ void stsm_x4_t(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};"
                 :: "r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void stsm_x2_t(uint32_t addr, uint32_t r0, uint32_t r1) {
    asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1,%2};"
                 :: "r"(addr), "r"(r0), "r"(r1) : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// byte offset of element (row r, k) inside a K-major swizzled operand whose 64-wide K blocks are
// `kblock_bytes` apart (= rows * 128)
__device__ __forceinline__ uint32_t swz_off(int r, int k, int kblock_bytes) {
    return (uint32_t)((k >> 6) * kblock_bytes + r * 128 + ((((k & 63) >> 3) ^ (r & 7)) << 4) + (k & 7) * 2);
}

}  // namespace sfb
