// The dense per-relation contractions of layer 2 on the 5th-generation tensor cores (reference: tf.nn.dropout +
// tf.matmul(x, W2_k), decagon/deep/layers.py:112-113, and their autodiff):
//   project_tc_kernel : P2_k  = (H_j (.) m_k / q) W2_k                 [n_j, D1] x [D1, 32]
//   dh_tc_kernel      : dH_j += (G2_k W2_k^T) (.) m_k / q              [n_j, 32] x [32, D1], masked per relation
//   dw2_tc_kernel     : dW2_k = (H_j (.) m_k / q)^T G2_k               [D1, n_j] x [n_j, 32]
// tcgen05.mma has no fp32 input kind; the fp32 contract (rel-err 1e-5) is met with the TF32 split x = hi + lo
// (tc_common.cuh).  The hi and lo parts of ONE operand are stacked along a free dimension of the MMA, so two MMAs per
// k-step give all four products (hi hi, hi lo, lo hi, lo lo) and the epilogue adds the two halves:
//   project : B = [W2_hi | W2_lo] (N = 64), A passes hi, lo into the same accumulator; out[n] = D[n] + D[32 + n]
//   dh      : B = [W2_hi ; W2_lo] (N = 2 D1), A = G2 hi, lo;                         t[m]  = D[m] + D[D1 + m]
//   dw2     : A = [Hm_hi ; Hm_lo] (M = 128, lo rows at 64), B = [G2_hi | G2_lo] (N = 64): one MMA per k-step;
//             dW2[m][n] = D[m][n] + D[m][32 + n] + D[64 + m][n] + D[64 + m][32 + n]
// Operand tiles (K-major, SWIZZLE_128B) are written by the threads: the dropout mask differs per relation, so the
// masked operand cannot come from a TMA copy of H.  Thread = one row of the 128-row tile; its H row lives in
// registers for the CTA's life (project), its dH sums live in registers (dh), the dW2 accumulator lives in TMEM
// across the row tiles of a relation (dw2).  CTAs are sequential inside (build -> MMA -> read back); two CTAs per SM
// overlap their phases.  hidden1 is 32 or 64 here (128 stays on the CUDA-core kernels of dense.cu), hidden2 = 32.
#include <algorithm>

#include "dgn_internal.cuh"
#include "tc_common.cuh"

namespace dgn {
namespace {

using namespace tc;

constexpr int kD2 = 32;
constexpr int kThreads = 128;
constexpr int kTile = 128;  // rows per tile = threads

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float sel(uint32_t word, int bit, float x) { return (word >> bit) & 1u ? x : 0.f; }

__device__ __forceinline__ void slot_range(int slot, int n_slots, int K, int &k_begin, int &k_end) {
    k_begin = (int)((long long)slot * K / n_slots);
    k_end = (int)((long long)(slot + 1) * K / n_slots);
}

struct TcSetup {
    uint32_t tmem;
};
// TMEM allocation + mbarrier, common prologue
template <uint32_t COLS>
__device__ __forceinline__ uint32_t tc_prologue(uint32_t *tmem_slot, uint64_t *bar, const void *smem) {
    if ((threadIdx.x >> 5) == 0) tmem_alloc<COLS>(tmem_slot);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    fence_before();
    __syncthreads();
    fence_after();
    if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B operand tiles need 1024-byte alignment
    return *tmem_slot;
}

// ------------------------------------------------------------------------------ P2 = Hm W2
// CTA = (row tile, slot of relations).  smem: A [KB][hi, lo][128 x 128 B], B [KB][64 x 128 B]
template <int D1>
__global__ void __launch_bounds__(kThreads, 2) project_tc_kernel(const DenseArgs a) {
    constexpr int KB = D1 / 32;
    constexpr uint32_t kIdesc = idesc_tf32(128, 64);
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *As = smem;
    unsigned char *Bs = smem + KB * 2 * 16384;
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rt = blockIdx.x % a.n_rb, slot = blockIdx.x / a.n_rb;
    const int row = rt * kTile + tid;
    const bool valid = row < a.n_j;
    int k_begin, k_end;
    slot_range(slot, a.n_slots, a.K, k_begin, k_end);
    const uint32_t tmem = tc_prologue<64>(&tmem_slot, &mma_done, smem);

    float4 x[KB][8];
#pragma unroll
    for (int p = 0; p < KB; ++p)
#pragma unroll
        for (int c = 0; c < 8; ++c) x[p][c] = valid ? ld4(a.H + ((size_t)p * a.n_j + row) * 32 + 4 * c) : zero4();
    const float sc = a.mask != nullptr ? a.scale : 1.f;
    uint32_t parity = 0;

    for (int k = k_begin; k < k_end; ++k) {
        uint32_t mk[KB];
#pragma unroll
        for (int p = 0; p < KB; ++p)
            mk[p] = (a.mask != nullptr && valid) ? __ldg(a.mask + ((size_t)k * a.n_j + row) * KB + p) : 0xffffffffu;
        // B: thread (n = lane, 4 consecutive m): W2_k[4 mc + j][n] -> row n (hi) / 32 + n (lo), chunk mc
        const float *W = a.W2 + (size_t)k * D1 * kD2;
#pragma unroll
        for (int i = 0; i < D1 / 16; ++i) {
            const int mc = warp + 4 * i, kb = mc >> 3, c = mc & 7;
            float4 w, hi, lo;
            w.x = __ldg(W + (4 * mc + 0) * kD2 + lane), w.y = __ldg(W + (4 * mc + 1) * kD2 + lane);
            w.z = __ldg(W + (4 * mc + 2) * kD2 + lane), w.w = __ldg(W + (4 * mc + 3) * kD2 + lane);
            split4(w, hi, lo);
            st128(Bs + kb * 8192 + sw128(lane, c), hi);
            st128(Bs + kb * 8192 + sw128(32 + lane, c), lo);
        }
        // A: this thread's row, masked
#pragma unroll
        for (int p = 0; p < KB; ++p)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 v, hi, lo;
                v.x = sel(mk[p], 4 * c + 0, x[p][c].x), v.y = sel(mk[p], 4 * c + 1, x[p][c].y);
                v.z = sel(mk[p], 4 * c + 2, x[p][c].z), v.w = sel(mk[p], 4 * c + 3, x[p][c].w);
                split4(v, hi, lo);
                const uint32_t off = sw128(tid, c);
                st128(As + (p * 2 + 0) * 16384 + off, hi);
                st128(As + (p * 2 + 1) * 16384 + off, lo);
            }
        fence_async_smem();
        fence_before();
        __syncthreads();
        if (tid == 0) {
            fence_after();
#pragma unroll
            for (int p = 0; p < KB; ++p) {
                const uint64_t ahi = umma_desc(smem_u32(As + (p * 2 + 0) * 16384)), alo = umma_desc(smem_u32(As + (p * 2 + 1) * 16384));
                const uint64_t b = umma_desc(smem_u32(Bs + p * 8192));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    mma_tf32(tmem, ahi + 2 * ks, b + 2 * ks, kIdesc, (p | ks) != 0);
                    mma_tf32(tmem, alo + 2 * ks, b + 2 * ks, kIdesc, 1);
                }
            }
            mma_commit(&mma_done);
        }
        mbar_wait(&mma_done, parity);
        parity ^= 1;
        fence_after();
        {
            float v0[32], v1[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
            tmem_ld32(taddr, v0);
            tmem_ld32(taddr + 32, v1);
            if (valid) {
                float *dst = a.P2 + ((size_t)k * a.n_j + row) * kD2;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<float4 *>(dst + 4 * c) =
                        make_float4((v0[4 * c] + v1[4 * c]) * sc, (v0[4 * c + 1] + v1[4 * c + 1]) * sc,
                                    (v0[4 * c + 2] + v1[4 * c + 2]) * sc, (v0[4 * c + 3] + v1[4 * c + 3]) * sc);
            }
        }
        fence_before();
        __syncthreads();  // TMEM and the operand tiles are free again
    }
    if (warp == 0) tmem_dealloc<64>(tmem);
}

// ------------------------------------------------------------------------------ dH += (G2 W2^T) (.) m
// CTA = (row tile, slot of relations).  smem: A hi, lo [128 x 128 B] (G2 rows), B [2 D1 x 128 B] (W2 hi ; lo)
template <int D1>
__global__ void __launch_bounds__(kThreads, 2) dh_tc_kernel(const DenseArgs a) {
    constexpr int KB = D1 / 32, N = 2 * D1;
    constexpr uint32_t kIdesc = idesc_tf32(128, N);
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *Ahi = smem, *Alo = smem + 16384, *Bs = smem + 32768;
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int rt = blockIdx.x % a.n_rb, slot = blockIdx.x / a.n_rb;
    const int row = rt * kTile + tid;
    const bool valid = row < a.n_j;
    int k_begin, k_end;
    slot_range(slot, a.n_slots, a.K, k_begin, k_end);
    const uint32_t tmem = tc_prologue<(uint32_t)N>(&tmem_slot, &mma_done, smem);

    float acc[KB][32];
#pragma unroll
    for (int p = 0; p < KB; ++p)
#pragma unroll
        for (int f = 0; f < 32; ++f) acc[p][f] = 0.f;
    uint32_t parity = 0;

    for (int k = k_begin; k < k_end; ++k) {
        uint32_t mk[KB];
#pragma unroll
        for (int p = 0; p < KB; ++p)
            mk[p] = (a.mask != nullptr && valid) ? __ldg(a.mask + ((size_t)k * a.n_j + row) * KB + p) : 0xffffffffu;
        const float *grow = a.G2 + ((size_t)k * a.n_j + row) * kD2;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float4 hi, lo;
            split4(valid ? ld4(grow + 4 * c) : zero4(), hi, lo);
            const uint32_t off = sw128(tid, c);
            st128(Ahi + off, hi);
            st128(Alo + off, lo);
        }
        const float *W = a.W2 + (size_t)k * D1 * kD2;
#pragma unroll
        for (int j = 0; j < D1 / 16; ++j) {
            const int i = tid + kThreads * j, m = i >> 3, c = i & 7;
            float4 hi, lo;
            split4(ld4(W + 4 * i), hi, lo);
            st128(Bs + sw128(m, c), hi);
            st128(Bs + sw128(D1 + m, c), lo);
        }
        fence_async_smem();
        fence_before();
        __syncthreads();
        if (tid == 0) {
            fence_after();
            const uint64_t ahi = umma_desc(smem_u32(Ahi)), alo = umma_desc(smem_u32(Alo)), b = umma_desc(smem_u32(Bs));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                mma_tf32(tmem, ahi + 2 * ks, b + 2 * ks, kIdesc, ks != 0);
                mma_tf32(tmem, alo + 2 * ks, b + 2 * ks, kIdesc, 1);
            }
            mma_commit(&mma_done);
        }
        mbar_wait(&mma_done, parity);
        parity ^= 1;
        fence_after();
        {
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
            for (int p = 0; p < KB; ++p) {
                float v0[32], v1[32];
                tmem_ld32(taddr + 32 * p, v0);
                tmem_ld32(taddr + D1 + 32 * p, v1);
#pragma unroll
                for (int f = 0; f < 32; ++f) acc[p][f] += sel(mk[p], f, v0[f] + v1[f]);
            }
        }
        fence_before();
        __syncthreads();
    }
    if (valid) {
        const float sc = a.mask != nullptr ? a.scale : 1.f;
#pragma unroll
        for (int p = 0; p < KB; ++p) {
            float *dst = a.dHpart + (((size_t)slot * KB + p) * a.n_j + row) * 32;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4 *>(dst + 4 * c) =
                    make_float4(acc[p][4 * c] * sc, acc[p][4 * c + 1] * sc, acc[p][4 * c + 2] * sc, acc[p][4 * c + 3] * sc);
        }
    }
    if (warp == 0) tmem_dealloc<(uint32_t)N>(tmem);
}

// ------------------------------------------------------------------------------ dW2 = Hm^T G2
// CTA = (relation, chunk of row tiles): the [D1, 32] result accumulates in TMEM over the chunk's row tiles.
// smem: A [4 kb][128 x 128 B] (rows = features: hi at m, lo at 64 + m; K = the tile's 128 node rows, 32 per
// k-block = one warp), B [4 kb][64 x 128 B] (rows = n: hi at n, lo at 32 + n).  Both are transposed on the way in:
// lane = K position, so one scalar store per (feature, lane) and the 32 lanes of a store hit 32 banks.
template <int D1>
__global__ void __launch_bounds__(kThreads, 2) dw2_tc_kernel(const DenseArgs a, int n_tiles) {
    constexpr int KB = D1 / 32;
    constexpr uint32_t kIdesc = idesc_tf32(128, 64);
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *As = smem;            // 4 x 16384
    unsigned char *Bs = smem + 65536;    // 4 x 8192
    float *stage = reinterpret_cast<float *>(smem + 65536 + 32768);  // [64][33]
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_chunks = a.n_rb;
    const int k = blockIdx.x / n_chunks, chunk = blockIdx.x % n_chunks;
    const int t_begin = (int)((long long)chunk * n_tiles / n_chunks), t_end = (int)((long long)(chunk + 1) * n_tiles / n_chunks);
    if (D1 < 64) {  // feature rows D1 .. 63 and 64 + D1 .. 127 of A are never written: they must read as zero
        for (int i = tid; i < 65536 / 16; i += kThreads) st128(As + 16 * i, zero4());
    }
    const uint32_t tmem = tc_prologue<64>(&tmem_slot, &mma_done, smem);

    // byte offset inside a row group for row-in-group i at this lane's K position
    uint32_t xo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xo[i] = (uint32_t)(i * 128 + (((lane >> 2) ^ i) << 4) + (lane & 3) * 4);
    unsigned char *Aw = As + warp * 16384, *Bw = Bs + warp * 8192;
    uint32_t parity = 0;
    bool pending = false;

    for (int t = t_begin; t < t_end; ++t) {
        const int row = t * kTile + tid;
        const bool valid = row < a.n_j;
        uint32_t mk[KB];
        float4 h[KB][8], gv[8];
#pragma unroll
        for (int p = 0; p < KB; ++p) {
            mk[p] = (a.mask != nullptr && valid) ? __ldg(a.mask + ((size_t)k * a.n_j + row) * KB + p) : 0xffffffffu;
#pragma unroll
            for (int c = 0; c < 8; ++c) h[p][c] = valid ? ld4(a.H + ((size_t)p * a.n_j + row) * 32 + 4 * c) : zero4();
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) gv[c] = valid ? ld4(a.G2 + ((size_t)k * a.n_j + row) * kD2 + 4 * c) : zero4();
        if (pending) {  // the previous tile's MMAs still read the operand tiles
            mbar_wait(&mma_done, parity);
            parity ^= 1;
            fence_after();
        }
#pragma unroll
        for (int p = 0; p < KB; ++p)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 v, hi, lo;
                v.x = sel(mk[p], 4 * c + 0, h[p][c].x), v.y = sel(mk[p], 4 * c + 1, h[p][c].y);
                v.z = sel(mk[p], 4 * c + 2, h[p][c].z), v.w = sel(mk[p], 4 * c + 3, h[p][c].w);
                split4(v, hi, lo);
                // features m = 32 p + 4 c + j: row group (m >> 3) = 4 p + (c >> 1), row in group 4 (c & 1) + j
                unsigned char *g_hi = Aw + (4 * p + (c >> 1)) * 1024, *g_lo = g_hi + 8 * 1024;
                const int i0 = 4 * (c & 1);
                st32(g_hi + xo[i0 + 0], hi.x), st32(g_hi + xo[i0 + 1], hi.y), st32(g_hi + xo[i0 + 2], hi.z), st32(g_hi + xo[i0 + 3], hi.w);
                st32(g_lo + xo[i0 + 0], lo.x), st32(g_lo + xo[i0 + 1], lo.y), st32(g_lo + xo[i0 + 2], lo.z), st32(g_lo + xo[i0 + 3], lo.w);
            }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float4 hi, lo;
            split4(gv[c], hi, lo);
            unsigned char *g_hi = Bw + (c >> 1) * 1024, *g_lo = g_hi + 4 * 1024;
            const int i0 = 4 * (c & 1);
            st32(g_hi + xo[i0 + 0], hi.x), st32(g_hi + xo[i0 + 1], hi.y), st32(g_hi + xo[i0 + 2], hi.z), st32(g_hi + xo[i0 + 3], hi.w);
            st32(g_lo + xo[i0 + 0], lo.x), st32(g_lo + xo[i0 + 1], lo.y), st32(g_lo + xo[i0 + 2], lo.z), st32(g_lo + xo[i0 + 3], lo.w);
        }
        fence_async_smem();
        fence_before();
        __syncthreads();
        if (tid == 0) {
            fence_after();
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
                const uint64_t ad = umma_desc(smem_u32(As + kb * 16384)), bd = umma_desc(smem_u32(Bs + kb * 8192));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma_tf32(tmem, ad + 2 * ks, bd + 2 * ks, kIdesc, (t != t_begin) || (kb | ks) != 0);
            }
            mma_commit(&mma_done);
        }
        pending = true;
    }
    if (pending) {
        mbar_wait(&mma_done, parity);
        fence_after();
    }
    {
        float v0[32], v1[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        tmem_ld32(taddr, v0);
        tmem_ld32(taddr + 32, v1);
#pragma unroll
        for (int n = 0; n < 32; ++n) v0[n] += v1[n];
        if (warp >= 2) {
#pragma unroll
            for (int n = 0; n < 32; ++n) stage[(tid - 64) * 33 + n] = v0[n];
        }
        fence_before();
        __syncthreads();
        if (warp < 2 && tid < D1) {
            const float sc = a.mask != nullptr ? a.scale : 1.f;
            float *dst = a.dW2 + ((size_t)k * n_chunks + chunk) * D1 * kD2 + tid * kD2;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4 *>(dst + 4 * c) = make_float4(
                    (v0[4 * c] + stage[tid * 33 + 4 * c]) * sc, (v0[4 * c + 1] + stage[tid * 33 + 4 * c + 1]) * sc,
                    (v0[4 * c + 2] + stage[tid * 33 + 4 * c + 2]) * sc, (v0[4 * c + 3] + stage[tid * 33 + 4 * c + 3]) * sc);
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem);
}

template <typename Kernel>
void set_smem(Kernel kernel, size_t bytes) {
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

}  // namespace

bool dense_tc_supported(int D1, int D2) { return D2 == kD2 && (D1 == 32 || D1 == 64); }
int dense_tc_tiles(int n_j) { return (n_j + kTile - 1) / kTile; }

// a.n_rb = row tiles of 128, a.n_slots = CTAs per row tile
void launch_project_tc(const DenseArgs &a, int D1, cudaStream_t s) {
    if (a.K == 0 || a.n_j == 0) return;
    if (D1 == 64) {
        const size_t bytes = 2 * 2 * 16384 + 2 * 8192;
        set_smem(project_tc_kernel<64>, bytes);
        project_tc_kernel<64><<<a.n_rb * a.n_slots, kThreads, bytes, s>>>(a);
    } else {
        const size_t bytes = 2 * 16384 + 8192;
        set_smem(project_tc_kernel<32>, bytes);
        project_tc_kernel<32><<<a.n_rb * a.n_slots, kThreads, bytes, s>>>(a);
    }
    CUDA_CHECK(cudaGetLastError());
}

void launch_dh_tc(const DenseArgs &a, int D1, cudaStream_t s) {
    if (a.K == 0 || a.n_j == 0) return;
    if (D1 == 64) {
        const size_t bytes = 32768 + 128 * 128;
        set_smem(dh_tc_kernel<64>, bytes);
        dh_tc_kernel<64><<<a.n_rb * a.n_slots, kThreads, bytes, s>>>(a);
    } else {
        const size_t bytes = 32768 + 64 * 128;
        set_smem(dh_tc_kernel<32>, bytes);
        dh_tc_kernel<32><<<a.n_rb * a.n_slots, kThreads, bytes, s>>>(a);
    }
    CUDA_CHECK(cudaGetLastError());
}

// a.n_rb = chunks of row tiles per relation (partials [K][n_rb][D1 * 32], n_rb == 1: the gradient itself)
void launch_dw2_tc(const DenseArgs &a, int D1, cudaStream_t s) {
    if (a.K == 0 || a.n_j == 0) return;
    const size_t bytes = 65536 + 32768 + 64 * 33 * sizeof(float);
    const int n_tiles = dense_tc_tiles(a.n_j);
    if (D1 == 64) {
        set_smem(dw2_tc_kernel<64>, bytes);
        dw2_tc_kernel<64><<<a.K * a.n_rb, kThreads, bytes, s>>>(a, n_tiles);
    } else {
        set_smem(dw2_tc_kernel<32>, bytes);
        dw2_tc_kernel<32><<<a.K * a.n_rb, kThreads, bytes, s>>>(a, n_tiles);
    }
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace dgn
