// All-pairs scoring on the 5th-generation tensor cores (reference: DecagonOptimizer.predict,
// decagon/deep/optimizer.py:87-106, evaluated for every relation by DecagonAccuracyEvaluator.py:57-91):
//     P_r = Z_i M_r Z_j^T,  M_r = loc_r glb_r loc_r (32 x 32),  P_r: [n_i, n_j] fp32
// for a run of relations of one group -- BASELINE config #4 (964 / 1928 relations x 645 x 645).
//
// tcgen05.mma has no fp32 input kind; the fp32 contract (rel-err 1e-5) is met with the 3 x TF32 split
//     a b ~= a_hi b_hi + a_lo b_hi + a_hi b_lo,   x_hi = x with the 13 low mantissa bits cleared, x_lo = x - x_hi
// (kind::tf32 ignores those 13 bits, so hi is exact and lo carries the next 11 bits: error ~2^-22).
//
// CTA = (column tile of 128 nodes of Z_j, worker), two CTAs per SM; a worker loops over relations.  Per relation
// the threads form B_r = Z_j M_r^T for the CTA's 128 columns on the CUDA cores (thread = column node, its
// Z_j row lives in registers for the whole kernel), split it and write hi / lo as K-major SWIZZLE_128B
// operand tiles into shared memory.  Per 128-row tile of Z_i the rows are split and written the same
// way, one elected thread issues the 12 tcgen05.mma (3 passes x 4 k-steps of 8) that accumulate the
// [128, 128] fp32 tile in tensor memory, tcgen05.commit signals an mbarrier, and the 4 warps read the
// accumulator back with tcgen05.ld (32 lanes x 32 columns per instruction), transpose 32 x 32 blocks
// through shared memory and store full 128-byte row pieces.  The kernel is bound by the output write
// (1.66 MB per relation); the tensor pipe is busy for a few per cent of the time by construction.
#include <algorithm>

#include "dgn_internal.cuh"

namespace dgn {
namespace {

constexpr int D = 32;            // hidden2 = K of the GEMM
constexpr int kTileM = 128;      // rows of Z_i per MMA
constexpr int kTileN = 128;      // columns (nodes of Z_j) per CTA = threads; two CTAs per SM overlap their phases
constexpr int kThreads = 128;
constexpr uint32_t kTmemCols = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of the 16-byte chunk `chunk` (4 floats) of row `row` in a K-major SWIZZLE_128B tile whose rows
// are 128 bytes (= 32 tf32 = the whole K): 8-row groups of 1024 bytes, chunk index XOR row-in-group
__device__ __forceinline__ uint32_t sw128(int row, int chunk) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in [0,14), leading byte
// offset >> 4 in [16,30) (unused for swizzled K-major: 1), stride byte offset >> 4 in [32,46) (1024 B between
// 8-row groups), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 at [4,6)), A = B = TF32 (2 at [7,10), [10,13)),
// both K-major (0 at bits 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__device__ __forceinline__ float relation_entry(int decoder, const float *glb, const float *loc, int p, int q) {
    switch (decoder) {
        case DGN_DEC_INNERPRODUCT: return p == q ? 1.f : 0.f;
        case DGN_DEC_DISTMULT: return p == q ? loc[p] : 0.f;
        case DGN_DEC_BILINEAR: return loc[p * D + q];
        default: return loc[p] * glb[p * D + q] * loc[q];  // dedicom
    }
}

__global__ void __launch_bounds__(kThreads, 2) predict_tc_kernel(const PredictArgs a, int n_workers) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *Bhi = smem;                         // [128 rows][128 B]
    unsigned char *Blo = smem + 16384;
    unsigned char *Ahi = smem + 32768;                 // [128 rows][128 B]
    unsigned char *Alo = smem + 49152;
    float *stage = reinterpret_cast<float *>(smem + 65536);  // [4 warps][32][33]
    __shared__ __align__(16) float Ms[D][D];
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ uint32_t tmem_base_slot;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nt = blockIdx.x / n_workers, worker = blockIdx.x % n_workers;
    const int v0 = nt * kTileN;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(&mma_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_slot;
    if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B operand tiles need 1024-byte alignment

    // this thread's column node: its embedding stays in registers
    float zj[D];
    {
        const int v = v0 + tid;
#pragma unroll
        for (int q = 0; q < D; q += 4) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < a.n_j) x = *reinterpret_cast<const float4 *>(a.Zj + (size_t)v * D + q);
            zj[q] = x.x, zj[q + 1] = x.y, zj[q + 2] = x.z, zj[q + 3] = x.w;
        }
    }
    const int n_mt = (a.n_i + kTileM - 1) / kTileM;
    uint32_t parity = 0;

    for (int k = worker; k < a.count; k += n_workers) {
        const float *loc = a.loc != nullptr ? a.loc + (size_t)k * a.loc_stride : nullptr;
        __syncthreads();  // the previous relation's MMAs and epilogue are done with Ms / B
        for (int i = tid; i < D * D; i += kThreads) Ms[i >> 5][i & 31] = relation_entry(a.decoder, a.glb, loc, i >> 5, i & 31);
        __syncthreads();
        // B_r[v][p] = sum_q M[p][q] Z_j[v][q]
        {
            float b[D];
#pragma unroll
            for (int p = 0; p < D; ++p) {
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < D; q += 4) {
                    const float4 m = *reinterpret_cast<const float4 *>(&Ms[p][q]);
                    s = fmaf(m.x, zj[q], s), s = fmaf(m.y, zj[q + 1], s), s = fmaf(m.z, zj[q + 2], s), s = fmaf(m.w, zj[q + 3], s);
                }
                b[p] = s;
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 hi, lo;
                hi.x = tf32_hi(b[4 * c]), hi.y = tf32_hi(b[4 * c + 1]), hi.z = tf32_hi(b[4 * c + 2]), hi.w = tf32_hi(b[4 * c + 3]);
                lo.x = b[4 * c] - hi.x, lo.y = b[4 * c + 1] - hi.y, lo.z = b[4 * c + 2] - hi.z, lo.w = b[4 * c + 3] - hi.w;
                const uint32_t off = sw128(tid, c);
                *reinterpret_cast<float4 *>(Bhi + off) = hi;
                *reinterpret_cast<float4 *>(Blo + off) = lo;
            }
        }
        float *out = a.out + (size_t)k * a.n_i * a.n_j;
        for (int mt = 0; mt < n_mt; ++mt) {
            const int u0 = mt * kTileM;
            // A tile: rows u0 .. u0 + 127 of Z_i, split; thread = row
            {
                const int row = tid;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (u0 + row < a.n_i) x = *reinterpret_cast<const float4 *>(a.Zi + (size_t)(u0 + row) * D + 4 * c);
                    float4 hi, lo;
                    hi.x = tf32_hi(x.x), hi.y = tf32_hi(x.y), hi.z = tf32_hi(x.z), hi.w = tf32_hi(x.w);
                    lo.x = x.x - hi.x, lo.y = x.y - hi.y, lo.z = x.z - hi.z, lo.w = x.w - hi.w;
                    const uint32_t off = sw128(row, c);
                    *reinterpret_cast<float4 *>(Ahi + off) = hi;
                    *reinterpret_cast<float4 *>(Alo + off) = lo;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t ahi = umma_desc(smem_u32(Ahi)), alo = umma_desc(smem_u32(Alo));
                const uint64_t bhi = umma_desc(smem_u32(Bhi)), blo = umma_desc(smem_u32(Blo));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {  // K = 8 tf32 = 32 bytes per instruction: + 2 in the encoded address
                    mma_tf32(tmem, ahi + 2 * ks, bhi + 2 * ks, ks > 0);
                    mma_tf32(tmem, alo + 2 * ks, bhi + 2 * ks, 1);
                    mma_tf32(tmem, ahi + 2 * ks, blo + 2 * ks, 1);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_done)) : "memory");
            }
            mbar_wait(&mma_done, parity);
            parity ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // epilogue: warp w reads TMEM lanes 32 w .. + 31 (rows), all 128 columns
            {
                float *st = stage + warp * 32 * 33;
                const int rbase = warp * 32, cbase = 0;
#pragma unroll 1
                for (int cc = 0; cc < 128; cc += 32) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem + ((uint32_t)rbase << 16) + (uint32_t)(cbase + cc);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) st[lane * 33 + j] = __uint_as_float(v[j]);
                    __syncwarp();
                    const int vcol = v0 + cbase + cc + lane;
                    if (vcol < a.n_j) {
#pragma unroll 8
                        for (int rr = 0; rr < 32; ++rr) {
                            const int u = u0 + rbase + rr;
                            if (u < a.n_i) out[(size_t)u * a.n_j + vcol] = st[rr * 33 + lane];
                        }
                    }
                    __syncwarp();
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();  // TMEM and the A tile are free again
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
}

}  // namespace

void launch_predict_tc(const PredictArgs &a, int n_sm, cudaStream_t s) {
    if (a.count == 0 || a.n_i == 0 || a.n_j == 0) return;
    const int n_nt = (a.n_j + kTileN - 1) / kTileN;
    const int n_workers = std::max(1, std::min(a.count, 2 * n_sm / n_nt));
    const size_t smem_bytes = 65536 + 4 * 32 * 33 * sizeof(float);
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(predict_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        configured = true;
    }
    predict_tc_kernel<<<n_nt * n_workers, kThreads, smem_bytes, s>>>(a, n_workers);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace dgn
