// All-pairs scoring on the 5th-generation tensor cores (reference: DecagonOptimizer.predict,
// decagon/deep/optimizer.py:87-106, evaluated for every relation by DecagonAccuracyEvaluator.py:57-91):
//     P_r = Z_i M_r Z_j^T,  M_r = loc_r glb_r loc_r (32 x 32),  P_r: [n_i, n_j] fp32
// for a run of relations of one group -- BASELINE config #4 (964 / 1928 relations x 645 x 645).
//
// tcgen05.mma has no fp32 input kind; the fp32 contract (rel-err 1e-5) is met with the 3 x TF32 split
//     a b ~= a_hi b_hi + a_lo b_hi + a_hi b_lo,   x_hi = x with the 13 low mantissa bits cleared, x_lo = x - x_hi
// (kind::tf32 ignores those 13 bits, so hi is exact and lo carries the next 11 bits: error ~2^-22).
//
// CTA = (column tile of NT <= 224 nodes of Z_j, worker), 256 threads, two CTAs per SM; a worker loops over relations.
// NT is the node count split evenly into tiles of at most 224 columns and rounded up to 16 (645 -> 3 x 224), so the
// padding of the last tile stays at a few per cent.  Per relation the threads form B_r = Z_j M_r^T for the CTA's
// columns on the CUDA cores (thread = column node, its Z_j row lives in registers for the whole kernel), split it
// and write hi / lo as K-major SWIZZLE_128B operand tiles.  Per 128-row tile of Z_i the rows are split and written
// the same way, one elected thread issues the 12 tcgen05.mma (3 passes x 4 k-steps of 8, N = NT) that accumulate
// the [128, NT] fp32 tile in tensor memory, tcgen05.commit signals an mbarrier, and the 8 warps read the accumulator
// back (tcgen05.ld, 32 lanes x 32 columns; the two warps of a lane quarter take alternate column chunks), transpose
// 32 x 32 blocks through shared memory (STS.128 into rows of 36 floats, LDS.32 along columns) and store 128-byte
// row pieces; full row tiles take a path without per-row bounds checks.  The kernel is bound by the output
// write (1.66 MB per relation) and by the instruction count of that read-back.
#include <algorithm>

#include "dgn_internal.cuh"
#include "tc_common.cuh"

namespace dgn {
namespace {

constexpr int D = 32;            // hidden2 = K of the GEMM
constexpr int kTileM = 128;      // rows of Z_i per MMA
constexpr int kMaxN = 224;       // columns (nodes of Z_j) per CTA at most
constexpr int kThreads = 256;
constexpr uint32_t kTmemCols = 256;
constexpr int kStageStride = 36;                       // floats per staged row: STS.128 / LDS.32 both conflict-free
constexpr int kStageBytes = 32 * kStageStride * 4;     // per warp
constexpr int kSmemB = kMaxN * 128;                    // one B tile (hi or lo)
constexpr int kSmemBytes = 2 * kSmemB + 8 * kStageBytes;  // B hi, B lo, then A hi / A lo overlaid by the 8 staging blocks
static_assert(8 * kStageBytes >= 2 * 16384, "the staging blocks cover the A tiles");

using namespace tc;

__device__ __forceinline__ float relation_entry(int decoder, const float *glb, const float *loc, int p, int q) {
    switch (decoder) {
        case DGN_DEC_INNERPRODUCT: return p == q ? 1.f : 0.f;
        case DGN_DEC_DISTMULT: return p == q ? loc[p] : 0.f;
        case DGN_DEC_BILINEAR: return loc[p * D + q];
        default: return loc[p] * glb[p * D + q] * loc[q];  // dedicom
    }
}

__global__ void __launch_bounds__(kThreads, 2) predict_tc_kernel(const PredictArgs a, int n_workers, int NT) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *Bhi = smem, *Blo = smem + kSmemB;
    unsigned char *Ahi = smem + 2 * kSmemB, *Alo = Ahi + 16384;
    __shared__ __align__(16) float Ms[D][D];
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ uint32_t tmem_base_slot;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nt = blockIdx.x / n_workers, worker = blockIdx.x % n_workers;
    const int v0 = nt * NT;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

    if (warp == 0) tmem_alloc<kTmemCols>(&tmem_base_slot);
    if (tid == 0) mbar_init(&mma_done, 1);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base_slot;
    if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B operand tiles need 1024-byte alignment

    const int n_mt = (a.n_i + kTileM - 1) / kTileM;
    const int n_chunks = (NT + 31) >> 5;
    const int quarter = warp & 3, half = warp >> 2;
    float *st = reinterpret_cast<float *>(Ahi + warp * kStageBytes);
    uint32_t parity = 0;
    const int arow = tid & 127, ac0 = (tid >> 7) * 4;
    float4 ax[4];
    auto load_a = [&](int mt) {
        const int u = mt * kTileM + arow;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            ax[c] = u < a.n_i ? __ldg(reinterpret_cast<const float4 *>(a.Zi + (size_t)u * D + 4 * (ac0 + c))) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    load_a(0);

    for (int k = worker; k < a.count; k += n_workers) {
        const float *loc = a.loc != nullptr ? a.loc + (size_t)k * a.loc_stride : nullptr;
        __syncthreads();  // the previous relation's MMAs and epilogue are done with Ms / B
        for (int i = tid; i < D * D; i += kThreads) Ms[i >> 5][i & 31] = relation_entry(a.decoder, a.glb, loc, i >> 5, i & 31);
        __syncthreads();
        // B_r[v][p] = sum_q M[p][q] Z_j[v][q]
        if (tid < NT) {
            // this thread's column node (re-read per relation from L2: keeping it would cost 32 registers that
            // the read-back loop needs for its loads in flight)
            float zj[D];
            {
                const int v = v0 + tid;
#pragma unroll
                for (int q = 0; q < D; q += 4) {
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (v < a.n_j) x = __ldg(reinterpret_cast<const float4 *>(a.Zj + (size_t)v * D + q));
                    zj[q] = x.x, zj[q + 1] = x.y, zj[q + 2] = x.z, zj[q + 3] = x.w;
                }
            }
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
                split4(make_float4(b[4 * c], b[4 * c + 1], b[4 * c + 2], b[4 * c + 3]), hi, lo);
                const uint32_t off = sw128(tid, c);
                st128(Bhi + off, hi);
                st128(Blo + off, lo);
            }
        }
        float *out = a.out + (size_t)k * a.n_i * a.n_j;
        for (int mt = 0; mt < n_mt; ++mt) {
            const int u0 = mt * kTileM;
            // A tile: rows u0 .. u0 + 127 of Z_i, split; thread = (row, half of the row's 8 chunks); the rows were
            // loaded one tile ahead
            {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4 hi, lo;
                    split4(ax[c], hi, lo);
                    const uint32_t off = sw128(arow, ac0 + c);
                    st128(Ahi + off, hi);
                    st128(Alo + off, lo);
                }
                load_a(mt + 1 < n_mt ? mt + 1 : 0);
            }
            fence_async_smem();
            fence_before();
            __syncthreads();
            if (warp == 0) {  // convergent warp, one elected lane issues
                fence_after();
                const uint64_t ahi = umma_desc(smem_u32(Ahi)), alo = umma_desc(smem_u32(Alo));
                const uint64_t bhi = umma_desc(smem_u32(Bhi)), blo = umma_desc(smem_u32(Blo));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {  // K = 8 tf32 = 32 bytes per instruction: + 2 in the encoded address
                    mma_tf32_elect(tmem, ahi + 2 * ks, bhi + 2 * ks, idesc, ks > 0);
                    mma_tf32_elect(tmem, alo + 2 * ks, bhi + 2 * ks, idesc, 1);
                    mma_tf32_elect(tmem, ahi + 2 * ks, blo + 2 * ks, idesc, 1);
                }
                mma_commit_elect(&mma_done);
            }
            mbar_wait(&mma_done, parity);
            parity ^= 1;
            fence_after();
            // read-back: warp (quarter, half) takes TMEM lanes 32 quarter .. + 31 (rows) and the column chunks
            // half, half + 2, ...; the staging block overlays the A tiles, which the finished MMAs no longer read
            {
                const int rbase = quarter * 32;
                const bool full = u0 + rbase + 32 <= a.n_i;
                for (int cc = half; cc < n_chunks; cc += 2) {
                    float v[32];
                    tmem_ld32(tmem + ((uint32_t)rbase << 16) + (uint32_t)(cc * 32), v);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4 *>(st + lane * kStageStride + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
                    const int col = cc * 32 + lane, vcol = v0 + col;
                    if (col < NT && vcol < a.n_j) {
                        float *p = out + (size_t)(u0 + rbase) * a.n_j + vcol;
                        if (full) {
#pragma unroll
                            for (int rr = 0; rr < 32; ++rr) {
                                *p = st[rr * kStageStride + lane];
                                p += a.n_j;
                            }
                        } else {
                            for (int rr = 0; rr < 32 && u0 + rbase + rr < a.n_i; ++rr) p[(size_t)rr * a.n_j] = st[rr * kStageStride + lane];
                        }
                    }
                    __syncwarp();
                }
            }
            fence_before();
            __syncthreads();  // TMEM and the A tiles / staging blocks are free again
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc<kTmemCols>(tmem);
}

}  // namespace

void launch_predict_tc(const PredictArgs &a, int n_sm, cudaStream_t s) {
    if (a.count == 0 || a.n_i == 0 || a.n_j == 0) return;
    const int n_nt = (a.n_j + kMaxN - 1) / kMaxN;
    const int NT = std::max(16, (((a.n_j + n_nt - 1) / n_nt) + 15) / 16 * 16);  // even split, MMA N granularity 16
    const int n_workers = std::max(1, std::min(a.count, 2 * n_sm / n_nt));
    static PerDeviceOnce configured;
    if (configured.first()) {
        CUDA_CHECK(cudaFuncSetAttribute(predict_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    }
    predict_tc_kernel<<<n_nt * n_workers, kThreads, kSmemBytes, s>>>(a, n_workers, NT);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace dgn
