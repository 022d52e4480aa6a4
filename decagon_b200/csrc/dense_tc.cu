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
// Operand tiles are written by the threads: the dropout mask differs per relation, so the masked operand cannot come
// from a TMA copy of H.  Layouts: K-major SWIZZLE_128B where the matrix lies K-contiguous in memory (project A, dh A
// and B), MN-major SWIZZLE_128B_BASE32B where it lies M/N-contiguous (project B = W2, dw2 A = Hm^T and B = G2^T; see
// tc_common.cuh and tools/umma_probe.cu), tensor memory for the A operand of project at hidden1 = 64
// (project_ts_kernel).  project: the thread's H row lives in registers for the CTA's life; dh: the dH sums live in
// registers; dw2: the accumulator lives in TMEM across the row tiles of a relation.  Global loads are coalesced and one
// tile ahead.  CTAs are sequential inside (write operands -> MMA -> read back); two CTAs per SM overlap their
// phases.  hidden1 is 32 or 64 here (128 stays on the CUDA-core kernels of dense.cu), hidden2 = 32.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "dgn_internal.cuh"
#include "tc_common.cuh"

namespace dgn {
namespace {

using namespace tc;

constexpr int kD2 = 32;
constexpr int kThreads = 128;

// Optional phase clock of project_tc_kernel (build with NVCC_EXTRA=-DDGN_TC_PROFILE): cycles summed over CTAs and
// iterations for [operand writes, fence + barrier, MMA issue, wait for the MMAs, read-back + stores, closing barrier]
#ifdef DGN_TC_PROFILE
__device__ unsigned long long g_tc_prof[8];
#define TC_CLK(i)                                   \
    do {                                            \
        const long long now__ = clock64();          \
        prof_acc[i] += now__ - prof_t;              \
        prof_t = now__;                             \
    } while (0)
#else
#define TC_CLK(i) do { } while (0)
#endif
constexpr int kTile = 128;  // rows per tile = threads

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float sel(uint32_t word, int bit, float x) { return (word >> bit) & 1u ? x : 0.f; }

__device__ __forceinline__ void slot_range(int slot, int n_slots, int K, int &k_begin, int &k_end) {
    k_begin = (int)((long long)slot * K / n_slots);
    k_end = (int)((long long)(slot + 1) * K / n_slots);
}

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
// CTA = (row tile, slot of relations), 256 threads: thread = (row of the tile, half h).  smem: A [KB][hi, lo]
// [128 x 128 B] K-major (masked rows), B [D1 / 8 k-blocks][hi, lo atoms] MN-major: W2_k is [D1][32] row-major, i.e.
// N-contiguous, and is copied as it lies.  Half h of a row's threads writes half of the row's chunks and reads back
// half of its output columns (the two warps of a TMEM lane quarter).  The next relation's keep words and W2 are
// loaded right after the tiles are written, so they fly during the MMAs and the read-back.
constexpr int kProjThreads = 256;
template <int D1>
__global__ void __launch_bounds__(kProjThreads, 2) project_tc_kernel(const DenseArgs a) {
    constexpr int KB = D1 / 32, NW = D1 / 32, NC = KB * 4;  // float4 of W2 per thread, chunks of the row per thread
    constexpr uint32_t kIdesc = idesc_tf32(128, 64, 0, 1);
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *As = smem;
    unsigned char *Bs = smem + KB * 2 * 16384;
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int rl = tid & 127, h = tid >> 7;
    const int rt = blockIdx.x % a.n_rb, slot = blockIdx.x / a.n_rb;
    const int row = rt * kTile + rl;
    const bool valid = row < a.n_j;
    int k_begin, k_end;
    slot_range(slot, a.n_slots, a.K, k_begin, k_end);
    const uint32_t tmem = tc_prologue<64>(&tmem_slot, &mma_done, smem);

    // this thread's chunks of the row: global chunk index gc = h NC + c (panel gc >> 3, chunk gc & 7)
    float4 x[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int gc = h * NC + c;
        x[c] = valid ? ld4(a.H + ((size_t)(gc >> 3) * a.n_j + row) * 32 + 4 * (gc & 7)) : zero4();
    }
    const float sc = a.mask != nullptr ? a.scale : 1.f;
    uint32_t parity = 0;
    uint32_t mk = 0xffffffffu;  // keep word of the panel this thread's chunks lie in (KB == 1: both halves share it)
    float4 w[NW];
    const int mp = (h * NC) >> 3;  // panel of this thread's chunks
    auto fetch = [&](int k) {
        mk = (a.mask != nullptr && valid) ? __ldg(a.mask + ((size_t)k * a.n_j + row) * KB + mp) : 0xffffffffu;
        const float *W = a.W2 + (size_t)k * D1 * kD2;
#pragma unroll
        for (int j = 0; j < NW; ++j) w[j] = ld4(W + 4 * (tid + kProjThreads * j));
    };
    if (k_begin < k_end) fetch(k_begin);
#ifdef DGN_TC_PROFILE
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = clock64();
#endif

    for (int k = k_begin; k < k_end; ++k) {
        // B: float4 i of W2_k = (K index m = i >> 3, N chunk c = i & 7)
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            const int i = tid + kProjThreads * j, m = i >> 3, c = i & 7;
            float4 hi, lo;
            split4(w[j], hi, lo);
            unsigned char *atom = Bs + (m >> 3) * 2048 + mn_off(m, c);
            st128(atom, hi);
            st128(atom + 1024, lo);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int gc = h * NC + c, cp = gc & 7;
            float4 v, hi, lo;
            v.x = sel(mk, 4 * cp + 0, x[c].x), v.y = sel(mk, 4 * cp + 1, x[c].y);
            v.z = sel(mk, 4 * cp + 2, x[c].z), v.w = sel(mk, 4 * cp + 3, x[c].w);
            split4(v, hi, lo);
            const uint32_t off = sw128(rl, cp);
            st128(As + ((gc >> 3) * 2 + 0) * 16384 + off, hi);
            st128(As + ((gc >> 3) * 2 + 1) * 16384 + off, lo);
        }
        if (k + 1 < k_end) fetch(k + 1);
        TC_CLK(0);
        fence_async_smem();
        fence_before();
        __syncthreads();
        TC_CLK(1);
        if (warp == 0) {  // the whole warp, convergent: one elected lane issues (tc_common.cuh)
            fence_after();
#pragma unroll
            for (int p = 0; p < KB; ++p) {
                const uint64_t ahi = umma_desc(smem_u32(As + (p * 2 + 0) * 16384)), alo = umma_desc(smem_u32(As + (p * 2 + 1) * 16384));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t b = umma_desc_mn(smem_u32(Bs + (p * 4 + ks) * 2048), 1024);
                    mma_tf32_elect(tmem, ahi + 2 * ks, b, kIdesc, (p | ks) != 0);
                    mma_tf32_elect(tmem, alo + 2 * ks, b, kIdesc, 1);
                }
            }
            mma_commit_elect(&mma_done);
        }
        TC_CLK(2);
        mbar_wait(&mma_done, parity);
        parity ^= 1;
        fence_after();
        TC_CLK(3);
        {
            // warp = (lane quarter warp & 3 = rl >> 5, half h): output columns [16 h, 16 h + 16)
            float v0[16], v1[16];
            const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 16 * h;
            tmem_ld16(taddr, v0);
            tmem_ld16(taddr + 32, v1);
            if (valid) {
                float *dst = a.P2 + ((size_t)k * a.n_j + row) * kD2 + 16 * h;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<float4 *>(dst + 4 * c) =
                        make_float4((v0[4 * c] + v1[4 * c]) * sc, (v0[4 * c + 1] + v1[4 * c + 1]) * sc,
                                    (v0[4 * c + 2] + v1[4 * c + 2]) * sc, (v0[4 * c + 3] + v1[4 * c + 3]) * sc);
            }
        }
        TC_CLK(4);
        fence_before();
        __syncthreads();  // TMEM and the operand tiles are free again
        TC_CLK(5);
    }
#ifdef DGN_TC_PROFILE
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) atomicAdd(&g_tc_prof[i], (unsigned long long)prof_acc[i]);
        atomicAdd(&g_tc_prof[6], (unsigned long long)(k_end - k_begin));
    }
#endif
    if (warp == 0) tmem_dealloc<64>(tmem);
}

// ------------------------------------------------------------------------------ P2 = Hm W2, A in tensor memory
// Same geometry as project_tc_kernel; the masked, split rows go to TENSOR memory (tcgen05.st: lane = the thread's
// row, one column per K element, no swizzle arithmetic, 256 B / clk) and the MMAs take A from there (TS form).
// Shared memory only carries B (16 KB written, 32 KB read per relation instead of 80 + 96 KB), so the MMA is no
// longer paced by operand reads.  TMEM columns: D [0, 64), A hi [64, 64 + D1), A lo [64 + D1, 64 + 2 D1).
// D1 = 64 only (each of the two threads of a row holds one 32-wide panel = one tcgen05.st.x32 per part).
__global__ void __launch_bounds__(kProjThreads, 2) project_ts_kernel(const DenseArgs a) {
    constexpr int D1 = 64, KB = 2, NW = 2;
    constexpr uint32_t kIdesc = idesc_tf32(128, 64, 0, 1);
    constexpr uint32_t kAhi = 64, kAlo = 64 + D1;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *Bs = smem;
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int rl = tid & 127, h = tid >> 7;
    const int rt = blockIdx.x % a.n_rb, slot = blockIdx.x / a.n_rb;
    const int row = rt * kTile + rl;
    const bool valid = row < a.n_j;
    int k_begin, k_end;
    slot_range(slot, a.n_slots, a.K, k_begin, k_end);
    const uint32_t tmem = tc_prologue<256>(&tmem_slot, &mma_done, smem);
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);

    float x[32];  // panel h of this thread's row
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 v = valid ? ld4(a.H + ((size_t)h * a.n_j + row) * 32 + 4 * c) : zero4();
        x[4 * c] = v.x, x[4 * c + 1] = v.y, x[4 * c + 2] = v.z, x[4 * c + 3] = v.w;
    }
    const float sc = a.mask != nullptr ? a.scale : 1.f;
    uint32_t parity = 0;
    uint32_t mk = 0xffffffffu;
    float4 w[NW];
    auto fetch = [&](int k) {
        mk = (a.mask != nullptr && valid) ? __ldg(a.mask + ((size_t)k * a.n_j + row) * KB + h) : 0xffffffffu;
        const float *W = a.W2 + (size_t)k * D1 * kD2;
#pragma unroll
        for (int j = 0; j < NW; ++j) w[j] = ld4(W + 4 * (tid + kProjThreads * j));
    };
    if (k_begin < k_end) fetch(k_begin);
#ifdef DGN_TC_PROFILE
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = clock64();
#endif

    for (int k = k_begin; k < k_end; ++k) {
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            const int i = tid + kProjThreads * j, m = i >> 3, c = i & 7;
            float4 hi, lo;
            split4(w[j], hi, lo);
            unsigned char *atom = Bs + (m >> 3) * 2048 + mn_off(m, c);
            st128(atom, hi);
            st128(atom + 1024, lo);
        }
        {
            float hi[32], lo[32];
#pragma unroll
            for (int f = 0; f < 32; ++f) {
                const float v = sel(mk, f, x[f]);
                hi[f] = tf32_hi(v);
                lo[f] = v - hi[f];
            }
            tmem_st32(lane_base + kAhi + 32 * h, hi);
            tmem_st32(lane_base + kAlo + 32 * h, lo);
            tmem_st_wait();
        }
        if (k + 1 < k_end) fetch(k + 1);
        TC_CLK(0);
        fence_async_smem();
        fence_before();
        __syncthreads();
        TC_CLK(1);
        if (warp == 0) {
            fence_after();
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const uint64_t b = umma_desc_mn(smem_u32(Bs + ks * 2048), 1024);
                mma_tf32_ts_elect(tmem, tmem + kAhi + 8 * ks, b, kIdesc, ks != 0);
                mma_tf32_ts_elect(tmem, tmem + kAlo + 8 * ks, b, kIdesc, 1);
            }
            mma_commit_elect(&mma_done);
        }
        TC_CLK(2);
        mbar_wait(&mma_done, parity);
        parity ^= 1;
        fence_after();
        TC_CLK(3);
        {
            float v0[16], v1[16];
            tmem_ld16(lane_base + 16 * h, v0);
            tmem_ld16(lane_base + 32 + 16 * h, v1);
            if (valid) {
                float *dst = a.P2 + ((size_t)k * a.n_j + row) * kD2 + 16 * h;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<float4 *>(dst + 4 * c) =
                        make_float4((v0[4 * c] + v1[4 * c]) * sc, (v0[4 * c + 1] + v1[4 * c + 1]) * sc,
                                    (v0[4 * c + 2] + v1[4 * c + 2]) * sc, (v0[4 * c + 3] + v1[4 * c + 3]) * sc);
            }
        }
        TC_CLK(4);
        fence_before();
        __syncthreads();  // the accumulator, the A columns and the B tile are free again
        TC_CLK(5);
    }
#ifdef DGN_TC_PROFILE
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) atomicAdd(&g_tc_prof[i], (unsigned long long)prof_acc[i]);
        atomicAdd(&g_tc_prof[6], (unsigned long long)(k_end - k_begin));
    }
#endif
    if (warp == 0) tmem_dealloc<256>(tmem);
}

// ------------------------------------------------------------------------------ dH += (G2 W2^T) (.) m
// CTA = (row tile, slot of relations).  smem: A hi, lo [128 x 128 B] (G2 rows, K-major), B [2 D1 x 128 B] (W2 hi ; lo,
// K-major as it lies).  Loads are coalesced (8 lanes = one 128-byte row) and one relation ahead.
template <int D1>
__global__ void __launch_bounds__(kThreads, 2) dh_tc_kernel(const DenseArgs a) {
    constexpr int KB = D1 / 32, N = 2 * D1, NW = D1 / 16;
    constexpr uint32_t kIdesc = idesc_tf32(128, N);
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *Ahi = smem, *Alo = smem + 16384, *Bs = smem + 32768;
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int rt = blockIdx.x % a.n_rb, slot = blockIdx.x / a.n_rb;
    const int row = rt * kTile + tid;  // the row this thread reads back from TMEM
    const bool valid = row < a.n_j;
    const int lrow = tid >> 3, lc = tid & 7;  // loads: rows lrow + 16 i, 16-byte chunk lc
    int k_begin, k_end;
    slot_range(slot, a.n_slots, a.K, k_begin, k_end);
    const uint32_t tmem = tc_prologue<(uint32_t)N>(&tmem_slot, &mma_done, smem);

    float acc[KB][32];
#pragma unroll
    for (int p = 0; p < KB; ++p)
#pragma unroll
        for (int f = 0; f < 32; ++f) acc[p][f] = 0.f;
    uint32_t parity = 0;
    uint32_t mk[KB], mkn[KB];
    float4 gq[8], w[NW];
    auto fetch = [&](int k) {
#pragma unroll
        for (int p = 0; p < KB; ++p)
            mkn[p] = (a.mask != nullptr && valid) ? __ldg(a.mask + ((size_t)k * a.n_j + row) * KB + p) : 0xffffffffu;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = rt * kTile + lrow + 16 * i;
            gq[i] = r < a.n_j ? ld4(a.G2 + ((size_t)k * a.n_j + r) * kD2 + 4 * lc) : zero4();
        }
        const float *W = a.W2 + (size_t)k * D1 * kD2;
#pragma unroll
        for (int j = 0; j < NW; ++j) w[j] = ld4(W + 4 * (tid + kThreads * j));
    };
    if (k_begin < k_end) fetch(k_begin);

    for (int k = k_begin; k < k_end; ++k) {
#pragma unroll
        for (int p = 0; p < KB; ++p) mk[p] = mkn[p];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 hi, lo;
            split4(gq[i], hi, lo);
            const uint32_t off = sw128(lrow + 16 * i, lc);
            st128(Ahi + off, hi);
            st128(Alo + off, lo);
        }
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            const int i = tid + kThreads * j, m = i >> 3, c = i & 7;
            float4 hi, lo;
            split4(w[j], hi, lo);
            st128(Bs + sw128(m, c), hi);
            st128(Bs + sw128(D1 + m, c), lo);
        }
        if (k + 1 < k_end) fetch(k + 1);
        fence_async_smem();
        fence_before();
        __syncthreads();
        if (warp == 0) {  // convergent warp, one elected lane issues
            fence_after();
            const uint64_t ahi = umma_desc(smem_u32(Ahi)), alo = umma_desc(smem_u32(Alo)), b = umma_desc(smem_u32(Bs));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                mma_tf32_elect(tmem, ahi + 2 * ks, b + 2 * ks, kIdesc, ks != 0);
                mma_tf32_elect(tmem, alo + 2 * ks, b + 2 * ks, kIdesc, 1);
            }
            mma_commit_elect(&mma_done);
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
// Persistent CTAs over units (relation, chunk of row tiles); the [D1, 32] result of a unit accumulates in TMEM over
// the unit's row tiles.  Both operands are MN-major, i.e. they are written the way H and G2 lie in memory (a node
// row = one K index, 32 features / columns = 128 contiguous bytes):
//   A [16 k-blocks of 8 nodes][4 atoms: hi panel 0, hi panel 1, lo panel 0, lo panel 1]  (M = 128, lo rows at 64 + m)
//   B [16 k-blocks of 8 nodes][2 atoms: hi, lo]                                           (N = 64)
// Loads are coalesced and one row tile ahead (issued right after the tiles of the current one are written); the
// read-back of a finished unit happens after the wait that precedes the next tile's writes.
template <int D1>
__global__ void __launch_bounds__(kThreads, 2) dw2_tc_kernel(const DenseArgs a, int n_tiles) {
    constexpr int KB = D1 / 32;
    constexpr uint32_t kIdesc = idesc_tf32(128, 64, 1, 1);
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *As = smem;            // 16 x 4096
    unsigned char *Bs = smem + 65536;    // 16 x 2048
    float *stage = reinterpret_cast<float *>(smem + 65536 + 32768);  // [64][33]
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int n_chunks = a.n_rb;
    const int n_units = a.K * n_chunks;
    const int u_begin = (int)((long long)blockIdx.x * n_units / gridDim.x), u_end = (int)((long long)(blockIdx.x + 1) * n_units / gridDim.x);
    if (D1 < 64) {  // the atoms of panel 1 are never written: they must read as zero
        for (int i = tid; i < 65536 / 16; i += kThreads) st128(As + 16 * i, zero4());
    }
    const uint32_t tmem = tc_prologue<64>(&tmem_slot, &mma_done, smem);
    const int lrow = tid >> 3, lc = tid & 7;      // loads: rows lrow + 16 i of the tile, 16-byte chunk lc
    const uint32_t toff = mn_off(lrow, lc);       // (lrow + 16 i) & 7 == lrow & 7
    const float sc = a.mask != nullptr ? a.scale : 1.f;

    uint32_t mk[KB][8];
    float4 h[KB][8], gv[8];
    auto fetch = [&](int k, int t) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = t * kTile + lrow + 16 * i;
            const bool ok = r < a.n_j;
#pragma unroll
            for (int p = 0; p < KB; ++p) {
                mk[p][i] = (a.mask != nullptr && ok) ? __ldg(a.mask + ((size_t)k * a.n_j + r) * KB + p) : 0xffffffffu;
                h[p][i] = ok ? ld4(a.H + ((size_t)p * a.n_j + r) * 32 + 4 * lc) : zero4();
            }
            gv[i] = ok ? ld4(a.G2 + ((size_t)k * a.n_j + r) * kD2 + 4 * lc) : zero4();
        }
    };
    // read the finished unit back: rows m (hi) + rows 64 + m (lo), columns n (hi) + 32 + n (lo)
    auto unit_epilogue = [&](int k, int chunk) {
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
            float *dst = a.dW2 + ((size_t)k * n_chunks + chunk) * D1 * kD2 + tid * kD2;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4 *>(dst + 4 * c) = make_float4(
                    (v0[4 * c] + stage[tid * 33 + 4 * c]) * sc, (v0[4 * c + 1] + stage[tid * 33 + 4 * c + 1]) * sc,
                    (v0[4 * c + 2] + stage[tid * 33 + 4 * c + 2]) * sc, (v0[4 * c + 3] + stage[tid * 33 + 4 * c + 3]) * sc);
        }
        __syncthreads();  // the staging buffer and TMEM are free again
        fence_after();
    };

    int u = u_begin;
    int k = 0, chunk = 0, t = 0, t_end = 0;
    auto open_unit = [&]() {
        k = u / n_chunks, chunk = u % n_chunks;
        t = (int)((long long)chunk * n_tiles / n_chunks), t_end = (int)((long long)(chunk + 1) * n_tiles / n_chunks);
    };
    if (u < u_end) {
        open_unit();
        fetch(k, t);
    }
    uint32_t parity = 0;
    bool pending = false, prev_closed = false, first = true;
    int prev_k = 0, prev_chunk = 0;
#ifdef DGN_TC_PROFILE
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = clock64();
    long long prof_tiles = 0;
#endif
    while (u < u_end) {
#ifdef DGN_TC_PROFILE
        ++prof_tiles;
#endif
        if (pending) {  // the previous tile's MMAs still read the operand tiles
            mbar_wait(&mma_done, parity);
            parity ^= 1;
            fence_after();
            if (prev_closed) unit_epilogue(prev_k, prev_chunk);
        }
        TC_CLK(0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int kblk = 2 * i + (lrow >> 3);  // (lrow + 16 i) >> 3
#pragma unroll
            for (int p = 0; p < KB; ++p) {
                float4 v, hi, lo;
                v.x = sel(mk[p][i], 4 * lc + 0, h[p][i].x), v.y = sel(mk[p][i], 4 * lc + 1, h[p][i].y);
                v.z = sel(mk[p][i], 4 * lc + 2, h[p][i].z), v.w = sel(mk[p][i], 4 * lc + 3, h[p][i].w);
                split4(v, hi, lo);
                unsigned char *atom = As + kblk * 4096 + p * 1024 + toff;
                st128(atom, hi);
                st128(atom + 2048, lo);
            }
            float4 hi, lo;
            split4(gv[i], hi, lo);
            unsigned char *atom = Bs + kblk * 2048 + toff;
            st128(atom, hi);
            st128(atom + 1024, lo);
        }
        const bool first_tile = first;
        // advance to the next tile and start its loads before the MMAs of this one are issued
        prev_k = k, prev_chunk = chunk, prev_closed = (t + 1 == t_end);
        first = false;
        if (++t == t_end) {
            ++u;
            first = true;
            if (u < u_end) open_unit();
        }
        TC_CLK(1);
        if (u < u_end) fetch(k, t);
        TC_CLK(2);
        fence_async_smem();
        fence_before();
        __syncthreads();
        TC_CLK(3);
        if (warp == 0) {  // convergent warp, one elected lane issues
            fence_after();
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
                const uint64_t ad = umma_desc_mn(smem_u32(As + ks * 4096), 1024), bd = umma_desc_mn(smem_u32(Bs + ks * 2048), 1024);
                mma_tf32_elect(tmem, ad, bd, kIdesc, !first_tile || ks != 0);
            }
            mma_commit_elect(&mma_done);
        }
        TC_CLK(4);
        pending = true;
    }
#ifdef DGN_TC_PROFILE
    if (tid == 0) {
        for (int i = 0; i < 5; ++i) atomicAdd(&g_tc_prof[i], (unsigned long long)prof_acc[i]);
        atomicAdd(&g_tc_prof[6], (unsigned long long)prof_tiles);
    }
#endif
    if (pending) {
        mbar_wait(&mma_done, parity);
        fence_after();
        unit_epilogue(prev_k, prev_chunk);
    }
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
    const char *ss = getenv("DGN_PROJECT_SS");  // 1: both operands from shared memory (project_tc_kernel)
    if (D1 == 64 && !(ss && ss[0] == '1')) {
        project_ts_kernel<<<a.n_rb * a.n_slots, kProjThreads, 16384, s>>>(a);
#ifdef DGN_TC_PROFILE
        if (a.K > 100) {
            unsigned long long h[8];
            cudaStreamSynchronize(s);
            cudaMemcpyFromSymbol(h, g_tc_prof, sizeof(h));
            const double it = (double)h[6];
            fprintf(stderr, "project_ts phases (thread 0, cycles per CTA-iteration over %.0f): writes %.0f  fence+bar %.0f  issue %.0f  "
                            "wait %.0f  readback %.0f  bar %.0f\n", it, h[0] / it, h[1] / it, h[2] / it, h[3] / it, h[4] / it, h[5] / it);
            unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z));
        }
#endif
    } else if (D1 == 64) {
        const size_t bytes = 2 * 2 * 16384 + 2 * 8192;
        set_smem(project_tc_kernel<64>, bytes);
        project_tc_kernel<64><<<a.n_rb * a.n_slots, kProjThreads, bytes, s>>>(a);
#ifdef DGN_TC_PROFILE
        if (a.K > 100) {
            unsigned long long h[8];
            cudaStreamSynchronize(s);
            cudaMemcpyFromSymbol(h, g_tc_prof, sizeof(h));
            const double it = (double)h[6];
            fprintf(stderr, "project_tc phases (thread 0, cycles per CTA-iteration over %.0f): writes %.0f  fence+bar %.0f  issue %.0f  "
                            "wait %.0f  readback %.0f  bar %.0f\n", it, h[0] / it, h[1] / it, h[2] / it, h[3] / it, h[4] / it, h[5] / it);
            unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z));
        }
#endif
    } else {
        const size_t bytes = 2 * 16384 + 8192;
        set_smem(project_tc_kernel<32>, bytes);
        project_tc_kernel<32><<<a.n_rb * a.n_slots, kProjThreads, bytes, s>>>(a);
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

// a.n_rb = chunks of row tiles per relation (partials [K][n_rb][D1 * 32], n_rb == 1: the gradient itself);
// a.n_slots = persistent CTAs
void launch_dw2_tc(const DenseArgs &a, int D1, cudaStream_t s) {
    if (a.K == 0 || a.n_j == 0) return;
    const size_t bytes = 65536 + 32768 + 64 * 33 * sizeof(float);
    const int n_tiles = dense_tc_tiles(a.n_j);
    const int grid = std::max(1, std::min(a.K * a.n_rb, a.n_slots));
    if (D1 == 64) {
        set_smem(dw2_tc_kernel<64>, bytes);
        dw2_tc_kernel<64><<<grid, kThreads, bytes, s>>>(a, n_tiles);
#ifdef DGN_TC_PROFILE
        if (a.K > 100) {
            unsigned long long h[8];
            cudaStreamSynchronize(s);
            cudaMemcpyFromSymbol(h, g_tc_prof, sizeof(h));
            const double it = (double)h[6];
            fprintf(stderr, "dw2_tc phases (thread 0, cycles per tile over %.0f tiles): wait prev MMA (+ unit epilogue) %.0f  operand writes %.0f  "
                            "next loads issued %.0f  fence+bar %.0f  issue %.0f\n", it, h[0] / it, h[1] / it, h[2] / it, h[3] / it, h[4] / it);
            unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z));
        }
#endif
    } else {
        set_smem(dw2_tc_kernel<32>, bytes);
        dw2_tc_kernel<32><<<grid, kThreads, bytes, s>>>(a, n_tiles);
    }
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace dgn
