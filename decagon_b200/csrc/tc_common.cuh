// tcgen05 / TMEM / mbarrier primitives shared by the tensor-core kernels (sm_100a only).
// Operand tiles are K-major SWIZZLE_128B tiles written by the threads themselves: rows of 128 bytes (32 tf32 of K),
// 8-row groups of 1024 bytes, the 16-byte chunk index XORed with the row index inside the group.
#pragma once

#include <stdint.h>

namespace dgn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t sw128(int row, int chunk) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in [0,14), leading byte offset >> 4
// in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
// K-major operand (row = one M / N index, 128 B = 32 tf32 of K): LBO unused (1), SBO = 1024 B between 8-row groups;
// one k-step of 8 tf32 = 32 bytes = + 2 in the encoded start address.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes = 16, uint32_t sbo_bytes = 1024) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major tf32 operand: layout type SWIZZLE_128B_BASE32B = 1 (plain SWIZZLE_128B reads as zeros for transposed
// 32-bit operands; measured with tools/umma_probe.cu on a B200).  An atom is 4 K-rows of 128 bytes, a row = 32
// consecutive M / N indices at one K; the 32-byte chunk index is XORed with the row index inside the atom
// (cute: Swizzle<2,5,2> o ((T,8,m),(4,k)):((1,T,LBO),(4T,SBO))).  Measured address map:
//   byte(mn, k) = (mn >> 5) LBO + (k >> 2) SBO + (k & 3) 128 + ((((mn & 31) >> 3) ^ (k & 3)) << 5) + (mn & 7) 4
// With SBO = 512 one k-step (8 K-rows) of one 32-wide M / N block is 1024 contiguous bytes.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes = 512) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46) | (1ull << 61);
}
// byte offset of 16-byte chunk `chunk` (4 tf32) of K-row `k_row` inside its 1024-byte k-step block (SBO = 512)
__device__ __forceinline__ uint32_t mn_off(int k_row, int chunk) {
    return (uint32_t)(((k_row >> 2) & 1) * 512 + (k_row & 3) * 128 + ((((chunk >> 1) ^ (k_row & 3))) << 5) + (chunk & 1) * 16);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 at [4,6)), A = B = TF32 (2 at [7,10), [10,13)),
// A / B major at bits 15 / 16 (0 = K-major, 1 = MN-major), N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t idesc_tf32(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The same instruction issued from a CONVERGENT warp: every lane carries the (warp-uniform) operands and one elected
// lane issues.  Inside a divergent `if (tid == 0)` the compiler has to move each descriptor into the uniform
// register file through a per-instruction loop, ~90 cycles per MMA; from a convergent warp the operands live in
// uniform registers to begin with.
__device__ __forceinline__ void mma_tf32_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in tensor memory (TS form): lane = row, one 32-bit column per K element, so a k-step of 8 tf32 is 8 columns
__device__ __forceinline__ void mma_tf32_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// this warp's 32 lanes x 32 consecutive columns <- v[j] (lane = row)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_commit_elect(uint64_t *bar) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    for (long long spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (spin > (1ll << 26)) __trap();  // a lost MMA must fault, not hang the device
    }
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the tensor core's async-proxy reads
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *slot) {  // one whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem) {  // one whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(COLS));
}
// 32 lanes (this warp's quarter of the 128 TMEM lanes) x 32 consecutive columns; lane = row, v[j] = column j
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

// 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// 3 x TF32 split: kind::tf32 ignores the 13 low mantissa bits, so hi is exact and lo carries the next bits
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void split4(const float4 &x, float4 &hi, float4 &lo) {
    hi.x = tf32_hi(x.x), hi.y = tf32_hi(x.y), hi.z = tf32_hi(x.z), hi.w = tf32_hi(x.w);
    lo.x = x.x - hi.x, lo.y = x.y - hi.y, lo.z = x.z - hi.z, lo.w = x.w - hi.w;
}
__device__ __forceinline__ void st128(unsigned char *p, const float4 &v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ void st32(unsigned char *p, float v) { *reinterpret_cast<float *>(p) = v; }

}  // namespace tc
}  // namespace dgn
