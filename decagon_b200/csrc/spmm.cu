// Multi-relation sparse x dense products of the encoder (reference: the two
// tf.sparse_tensor_dense_matmul call sites per relation and the tf.add_n over relations,
// decagon/deep/layers.py:89-92 and :113-116, and their autodiff transposes).
//
// Two kernels:
//   spmm_seg_kernel     -- "gather" path.  The K relation matrices of a group are viewed as ONE
//                          CSR matrix [A_0 | A_1 | ... | A_{K-1}] (forward) or its transpose
//                          (backward); rows are cut into segments of <= seg_len non-zeros, one
//                          warp per segment, lanes = the 32 features of a panel.  The dense
//                          operand is gathered from L2 / L1.
//   spmm_staged_kernel  -- "staged" path for groups of many small relations (the 1928
//                          drug-drug matrices): a persistent CTA per SM owns one 32-feature
//                          panel and a list of relations; each relation's operand tile
//                          ([n_j, 32] floats, contiguous in HBM) is brought into shared memory
//                          with one bulk-async copy (TMA, cp.async.bulk) double-buffered against
//                          the gather of the previous relation, and the row sums are accumulated
//                          in registers across all relations of the CTA.
// Dense matrices use the panel layout [P][rows][32] so that every tile is contiguous.
#include "dgn_internal.cuh"

namespace dgn {

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ bool mask_bit(const uint32_t *__restrict__ mask, int idx) {
    return (__ldg(mask + (idx >> 5)) >> (idx & 31)) & 1u;
}

// ------------------------------------------------------------------------------ gather path
// One quarter-warp (8 lanes x float4 = a 32-feature panel row) per segment, 4 segments per warp
// in lockstep; the 8 lanes fetch 8 (column, value) pairs at a time and broadcast them inside the
// quarter-warp; operand rows are 128-bit loads through L1.
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

template <int P>
__global__ void __launch_bounds__(256) spmm_seg_kernel(const SpmmArgs a) {
    const int l8 = threadIdx.x & 7;
    const int seg = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 3);
    const bool live = seg < a.n_seg;

    int row = 0, begin = 0, end = 0;
    bool single = true;
    if (live) {
        if (a.seg_row != nullptr) {
            row = __ldg(a.seg_row + seg);
            begin = __ldg(a.seg_begin + seg);
            // the segment ends where the row's next segment begins (hub rows have longer segments than seg_len)
            end = (seg + 1 < a.n_seg && __ldg(a.seg_row + seg + 1) == row) ? __ldg(a.seg_begin + seg + 1) : __ldg(a.rowptr + row + 1);
            single = (__ldg(a.row_seg_ptr + row + 1) - __ldg(a.row_seg_ptr + row)) == 1;
        } else {
            row = seg;
            begin = __ldg(a.rowptr + row);
            end = __ldg(a.rowptr + row + 1);
        }
    }
    const int n = end - begin;
    int nmax = max(n, __shfl_xor_sync(kFull, n, 8));
    nmax = max(nmax, __shfl_xor_sync(kFull, nmax, 16));

    float4 acc[P];
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);

    const float *__restrict__ op = a.op + (l8 << 2);
    // the (column, value) pairs of the NEXT batch of 8 are loaded before the operand rows of the current one are
    // gathered: one memory latency per batch instead of two dependent ones (ncu: long-scoreboard 26 warps per issue)
    auto fetch = [&](int base, int &c, float &v) {
        const int idx = begin + base + l8;
        c = 0;
        v = 0.f;
        if (idx < end) {
            c = __ldg(a.col + idx);
            v = __ldg(a.val + idx);
            if (a.col_mask) v = mask_bit(a.mask, c) ? v * a.scale : 0.f;
        }
    };
    int c_next;
    float v_next;
    fetch(0, c_next, v_next);
    for (int base = 0; base < nmax; base += 8) {
        const int c = c_next;
        const float v = v_next;
        if (base + 8 < nmax) fetch(base + 8, c_next, v_next);
        const int cnt = min(8, nmax - base);
#pragma unroll 8
        for (int j = 0; j < cnt; ++j) {
            const int cc = __shfl_sync(kFull, c, j, 8);
            const float vv = __shfl_sync(kFull, v, j, 8);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float4 x = ldg4(op + ((size_t)p * a.op_rows + cc) * 32);
                acc[p].x = fmaf(vv, x.x, acc[p].x);
                acc[p].y = fmaf(vv, x.y, acc[p].y);
                acc[p].z = fmaf(vv, x.z, acc[p].z);
                acc[p].w = fmaf(vv, x.w, acc[p].w);
            }
        }
    }
    if (!live) return;

    if (a.force_partial || !single) {
#pragma unroll
        for (int p = 0; p < P; ++p)
            *reinterpret_cast<float4 *>(a.partial + ((size_t)seg * P + p) * 32 + (l8 << 2)) = acc[p];
    } else {
        float s = 1.f;
        if (a.row_mask) s = mask_bit(a.mask, row) ? a.scale : 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p)
            *reinterpret_cast<float4 *>(a.out + ((size_t)p * a.out_rows + row) * 32 + (l8 << 2)) =
                make_float4(acc[p].x * s, acc[p].y * s, acc[p].z * s, acc[p].w * s);
    }
}

// rows of the backward products that were split over several segments: ordered sum
template <int P>
__global__ void __launch_bounds__(256) seg_reduce_kernel(const SpmmArgs a, const int *__restrict__ multi_rows,
                                                         int n_multi) {
    const int lane = threadIdx.x & 31;
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (w >= n_multi) return;
    const int row = multi_rows[w];
    const int s0 = a.row_seg_ptr[row], s1 = a.row_seg_ptr[row + 1];
    // four interleaved partial sums (a fixed order all the same): hub rows span hundreds of segments and a
    // single dependent chain pays one L2 round trip per segment
    float acc[P], b1[P], b2[P], b3[P];
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = b1[p] = b2[p] = b3[p] = 0.f;
    int s = s0;
    for (; s + 4 <= s1; s += 4)
#pragma unroll
        for (int p = 0; p < P; ++p) {
            acc[p] += a.partial[((size_t)s * P + p) * 32 + lane];
            b1[p] += a.partial[((size_t)(s + 1) * P + p) * 32 + lane];
            b2[p] += a.partial[((size_t)(s + 2) * P + p) * 32 + lane];
            b3[p] += a.partial[((size_t)(s + 3) * P + p) * 32 + lane];
        }
    for (; s < s1; ++s)
#pragma unroll
        for (int p = 0; p < P; ++p) acc[p] += a.partial[((size_t)s * P + p) * 32 + lane];
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = (acc[p] + b1[p]) + (b2[p] + b3[p]);
    float sc = 1.f;
    if (a.row_mask) sc = mask_bit(a.mask, row) ? a.scale : 0.f;
#pragma unroll
    for (int p = 0; p < P; ++p) a.out[((size_t)p * a.out_rows + row) * 32 + lane] = acc[p] * sc;
}

// ------------------------------------------------------------------------------ staged path
constexpr int kStagedThreads = 1024;
constexpr int kStagedWarps = kStagedThreads / 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    // bounded: a lost transaction must trap, not hang the device
    for (long long spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (spin > (1ll << 26)) __trap();
    }
}
// TMA bulk copy global -> shared, completion signalled on the mbarrier (UBLKCP in SASS)
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Quarter-warp rows: 8 lanes x float4 = the 32 features of a panel row, so one warp works on
// 4 adjacent output rows at once and one LDS.128 serves 4 non-zeros (4 x 128 B wavefronts).
// Row u is owned by quarter-warp (u % 128) for the whole kernel (slot u / 128), its running sum
// lives in registers across every relation of the CTA.
constexpr int kQuarters = kStagedThreads / 8;  // 128 quarter-warps per CTA

template <int RPQ>
__global__ void __launch_bounds__(kStagedThreads, 1) spmm_staged_kernel(const StagedArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[2];

    const int lane = threadIdx.x & 31;
    const int l8 = threadIdx.x & 7, qw = threadIdx.x >> 3;
    const int p = blockIdx.x % a.P, slot = blockIdx.x / a.P;
    const int r_begin = a.slot_ptr[slot], n_rel = a.slot_ptr[slot + 1] - r_begin;
    const int tile_floats = a.n_j * 32;
    const uint32_t tile_bytes = (uint32_t)tile_floats * 4u;
    const unsigned qmask = 0xffu << (lane & 24);  // the 8 lanes of this quarter-warp

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int t) {  // thread 0 only
        const int k = a.slot_rel[r_begin + t];
        uint64_t *bar = &full[t & 1];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, tile_bytes);
        bulk_load(smem_raw + (size_t)(t & 1) * tile_bytes, a.op + ((size_t)p * a.K + k) * tile_floats, tile_bytes, bar);
    };
    if (threadIdx.x == 0 && n_rel > 0) issue(0);

    float4 acc[RPQ];
#pragma unroll
    for (int s = 0; s < RPQ; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int t = 0; t < n_rel; ++t) {
        // buffer (t+1)&1 was last read for relation t-1; every warp passed the barrier that
        // closed iteration t-1, so it may be overwritten now
        if (threadIdx.x == 0 && t + 1 < n_rel) issue(t + 1);
        const int k = a.slot_rel[r_begin + t];
        const int *__restrict__ rp = a.rowptr + (size_t)k * (a.n_i + 1);
        // lane s of the quarter-warp fetches the row pointers of its s-th row
        int my_b = 0, my_e = 0;
        {
            const int u = qw + l8 * kQuarters;
            if (l8 < RPQ && u < a.n_i) {
                my_b = __ldg(rp + u);
                my_e = __ldg(rp + u + 1);
            }
        }
        // first 8 entries of a row slot are fetched one slot ahead of their use
        auto first_chunk = [&](int s, int &off, float &v) {
            const int b = __shfl_sync(qmask, my_b, s, 8), e = __shfl_sync(qmask, my_e, s, 8);
            const int idx = b + l8;
            off = 0;
            v = 0.f;
            if (idx < e) {
                off = __ldg(a.col + idx) << 7;  // byte offset of the operand row in the tile
                v = __ldg(a.val + idx);
            }
        };
        int next_off;
        float next_v;
        first_chunk(0, next_off, next_v);
        mbar_wait(&full[t & 1], (uint32_t)((t >> 1) & 1));
        unsigned char *tile = smem_raw + (size_t)(t & 1) * tile_bytes;
        if (a.mask != nullptr) {
            // layer-1 dropout on identity features drops whole operand rows: zero them in the
            // staged tile (the 1/keep scale is applied once, to the final sums)
            const int base_bit = k * a.n_j;
            for (int c = threadIdx.x >> 3; c < a.n_j; c += kQuarters)
                if (!mask_bit(a.mask, base_bit + c))
                    *reinterpret_cast<float4 *>(tile + ((size_t)c << 7) + (l8 << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
        }
        const unsigned char *xrow = tile + (l8 << 4);
#pragma unroll
        for (int s = 0; s < RPQ; ++s) {
            const int b = __shfl_sync(qmask, my_b, s, 8), e = __shfl_sync(qmask, my_e, s, 8);
            int n = e - b;  // non-zeros of this quarter-warp's row; the warp runs to the longest of its 4 rows
            int nmax = max(n, __shfl_xor_sync(kFull, n, 8));
            nmax = max(nmax, __shfl_xor_sync(kFull, nmax, 16));
            int off = next_off;
            float v = next_v;
            if (s + 1 < RPQ) first_chunk(s + 1, next_off, next_v);
            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int base = 0; base < nmax; base += 8) {
                if (base > 0) {
                    const int idx = b + base + l8;
                    off = 0;
                    v = 0.f;
                    if (idx < e) {
                        off = __ldg(a.col + idx) << 7;
                        v = __ldg(a.val + idx);
                    }
                }
                const int cnt = min(8, nmax - base);
#pragma unroll 4
                for (int j = 0; j < cnt; ++j) {
                    const int o = __shfl_sync(kFull, off, j, 8);
                    const float vv = __shfl_sync(kFull, v, j, 8);
                    const float4 x = *reinterpret_cast<const float4 *>(xrow + o);
                    sum.x = fmaf(vv, x.x, sum.x);
                    sum.y = fmaf(vv, x.y, sum.y);
                    sum.z = fmaf(vv, x.z, sum.z);
                    sum.w = fmaf(vv, x.w, sum.w);
                }
            }
            acc[s].x += sum.x;
            acc[s].y += sum.y;
            acc[s].z += sum.z;
            acc[s].w += sum.w;
        }
        __syncthreads();
    }

    const float sc = a.mask != nullptr ? a.scale : 1.f;
#pragma unroll
    for (int s = 0; s < RPQ; ++s) {
        const int u = qw + s * kQuarters;
        if (u < a.n_i)
            *reinterpret_cast<float4 *>(a.partial + (((size_t)slot * a.P + p) * a.n_i + u) * 32 + (l8 << 2)) =
                make_float4(acc[s].x * sc, acc[s].y * sc, acc[s].z * sc, acc[s].w * sc);
    }
}


// ------------------------------------------------------------------------------ staged path v3
// Warp-task streams (TaskArgs).  Every warp reads ONE contiguous stream of (offset, value) steps that
// spans all relations of its CTA.  The stream comes straight from HBM, so it is moved by cp.async
// through a warp-private 4-stage ring in shared memory (two blocks always in flight per warp, no
// registers held); the inner loop is LDS.128 (two steps of the stream) + 2 x (LDS.128 of the operand row
// + 4 FFMA): no row pointers, no shuffles, no per-row control beyond one counter.
// Forward: operand tiles move through a full / ready / empty mbarrier pipeline run by a dedicated
// producer warp (TMA bulk copy, then layer-1 dropout = zeroing the dropped operand rows in the tile);
// there is NO CTA-wide barrier per relation and the row sums of all relations of the CTA stay in
// registers.
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fma4(float4 &acc, float v, const float4 &x) {
    acc.x = fmaf(v, x.x, acc.x);
    acc.y = fmaf(v, x.y, acc.y);
    acc.z = fmaf(v, x.z, acc.z);
    acc.w = fmaf(v, x.w, acc.w);
}
// one stream step: acc += value * tile[row].  Padding steps carry (offset 0, value 0): they read row 0 and
// add nothing, so the loop has no predicates or branches.
__device__ __forceinline__ void gather_fma(float4 &acc, const unsigned char *xrow, int off, int vbits) {
    fma4(acc, __int_as_float(vbits), *reinterpret_cast<const float4 *>(xrow + off));
}
// uint16 pair-step count of slot s: word s / 2 of the (relation, warp) header, held by lane s / 2
__device__ __forceinline__ int task_count(int h, int s) {
    return (__shfl_sync(kFull, h, s >> 1) >> ((s & 1) * 16)) & 0xffff;
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// one instruction moves a whole contiguous range towards L2 (bytes: a multiple of 16)
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

constexpr int kRingStages = 4;
constexpr int kRingBlockPs = 8;                                  // pair-steps per block
constexpr int kRingBlockBytes = kRingBlockPs * 64;               // 512 B = 32 lanes x 16 B
constexpr int kRingBytes = kRingStages * kRingBlockBytes;        // per warp

struct StreamReader {
    unsigned char *ring;  // this warp's ring + quarter * 16
    const int4 *src;      // the warp's stream + lane (what this lane copies)
    int gi;               // pair-steps consumed
    int fetched;          // blocks issued

    __device__ __forceinline__ void init(unsigned char *warp_ring, const int4 *stream, int lane) {
        ring = warp_ring + (lane >> 3) * 16;
        src = stream + lane;
        gi = 0;
        fetched = 0;
    }
    // issue the next block (it overwrites the stage of block fetched - 4, which is fully consumed), then
    // wait until at most two blocks are pending: blocks <= fetched - 3 are complete
    __device__ __forceinline__ void refill() {
        const int lane = threadIdx.x & 31;
        unsigned char *dst = ring - (lane >> 3) * 16 + (fetched & (kRingStages - 1)) * kRingBlockBytes + lane * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src + (size_t)fetched * (kRingBlockPs * 4))
                     : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        ++fetched;
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncwarp();
    }
    // pair-steps gi .. gi + d are readable afterwards (d < kRingBlockPs)
    __device__ __forceinline__ void ensure(int d) {
        while (((gi + d) >> 3) > fetched - 3) refill();
    }
    __device__ __forceinline__ int4 load(int j) const {
        return *reinterpret_cast<const int4 *>(ring + (((gi + j) << 6) & (kRingBytes - 1)));
    }
};

// Forward variant: the lock-step padding of the forward streams is ~25 % of the steps.  A padding step is
// predicated off, so a quarter-warp whose step is padding issues no shared-memory wavefront (the LDS.128 of a
// warp is served one quarter-warp at a time); costs one ISETP per step, saves the wavefront.
__device__ __forceinline__ void gather_fma_pred(float4 &acc, const unsigned char *xrow, int off, int vbits) {
    if (vbits != 0) fma4(acc, __int_as_float(vbits), *reinterpret_cast<const float4 *>(xrow + off));
}
// n2 pair-steps of the warp's stream
template <typename F>
__device__ __forceinline__ void stream_steps(StreamReader &rd, int n2, F &&step) {
    int ps = 0;
#pragma unroll 1
    for (; ps + 4 <= n2; ps += 4) {
        rd.ensure(3);
        const int4 e0 = rd.load(0), e1 = rd.load(1), e2 = rd.load(2), e3 = rd.load(3);
        rd.gi += 4;
        step(e0.x, e0.y), step(e0.z, e0.w);
        step(e1.x, e1.y), step(e1.z, e1.w);
        step(e2.x, e2.y), step(e2.z, e2.w);
        step(e3.x, e3.y), step(e3.z, e3.w);
    }
#pragma unroll 1
    for (; ps < n2; ++ps) {
        rd.ensure(0);
        const int4 e0 = rd.load(0);
        rd.gi += 1;
        step(e0.x, e0.y), step(e0.z, e0.w);
    }
}

constexpr bool kForwardPredicated = false;  // measured on B200: 276.7 us vs 277.3 us for layer 1 at the polypharmacy shape, no gain
// slots S .. RPQ - 1 of one relation (compile-time recursion keeps acc[] in registers)
template <int S, int RPQ>
__device__ __forceinline__ void stream_slots(float4 (&acc)[RPQ], int h, StreamReader &rd, const unsigned char *xrow) {
    if constexpr (S < RPQ) {
        if (kForwardPredicated) {
            stream_steps(rd, task_count(h, S), [&](int off, int vbits) { gather_fma_pred(acc[S], xrow, off, vbits); });
        } else {
            stream_steps(rd, task_count(h, S), [&](int off, int vbits) { gather_fma(acc[S], xrow, off, vbits); });
        }
        stream_slots<S + 1, RPQ>(acc, h, rd, xrow);
    }
}

template <int RPQ>
__global__ void __launch_bounds__(kStagedThreads, 1) spmm_staged3_kernel(const TaskArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[2], ready[2], empty[2];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l8 = lane & 7, quarter = lane >> 3;
    const int p = blockIdx.x % a.P, slot = blockIdx.x / a.P;
    const int r_begin = a.slot_ptr[slot], n_rel = a.slot_ptr[slot + 1] - r_begin;
    const uint32_t tile_bytes = (uint32_t)a.n_op_rows * 128u;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            mbar_init(&full[b], 1);
            mbar_init(&ready[b], 1);
            mbar_init(&empty[b], kS3Warps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kS3Warps) {  // producer warp
        for (int t = 0; t < n_rel; ++t) {
            const int b = t & 1;
            const int k = a.slot_rel[r_begin + t];
            unsigned char *tile = smem_raw + (size_t)b * tile_bytes;
            if (lane == 0) {
                if (t >= 2) mbar_wait(&empty[b], (uint32_t)(((t >> 1) - 1) & 1));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&full[b], tile_bytes);
                bulk_load(tile, a.op + ((size_t)p * a.K + k) * a.n_op_rows * 32, tile_bytes, &full[b]);
            }
            __syncwarp();
            mbar_wait(&full[b], (uint32_t)((t >> 1) & 1));
            if (a.mask != nullptr) {
                // layer-1 dropout on identity features drops whole operand rows (1/keep is applied to the sums)
                const int base_bit = k * a.n_op_rows;
                for (int c = lane; c < a.n_op_rows; c += 32)
                    if (!mask_bit(a.mask, base_bit + c)) {
                        float4 *row = reinterpret_cast<float4 *>(tile + ((size_t)c << 7));
#pragma unroll
                        for (int q = 0; q < 8; ++q) row[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[b]);
        }
        return;
    }

    float4 acc[RPQ];
#pragma unroll
    for (int s = 0; s < RPQ; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);

    StreamReader rd;
    rd.init(smem_raw + 2 * (size_t)tile_bytes + (size_t)warp * kRingBytes,
            a.ent + (size_t)__ldg(a.wstart + slot * kS3Warps + warp) * 4, lane);
    int h_next = 0;
    if (n_rel > 0) {
        h_next = __ldg(a.hdr + ((size_t)a.slot_rel[r_begin] * kS3Warps + warp) * 4 + (lane & 3));
        rd.ensure(3);  // three blocks in flight before the first tile is waited for
    }
    for (int t = 0; t < n_rel; ++t) {
        const int h = h_next;
        if (t + 1 < n_rel) h_next = __ldg(a.hdr + ((size_t)a.slot_rel[r_begin + t + 1] * kS3Warps + warp) * 4 + (lane & 3));
        mbar_wait(&ready[t & 1], (uint32_t)((t >> 1) & 1));
        mbar_wait(&full[t & 1], (uint32_t)((t >> 1) & 1));
        const unsigned char *xrow = smem_raw + (size_t)(t & 1) * tile_bytes + (l8 << 4);
        stream_slots<0, RPQ>(acc, h, rd, xrow);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[t & 1]);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    const float sc = a.mask != nullptr ? a.scale : 1.f;
#pragma unroll
    for (int s = 0; s < RPQ; ++s) {
        const int u = __ldg(a.orow + (warp * a.rpq + s) * 4 + quarter);
        if (u >= 0)
            *reinterpret_cast<float4 *>(a.out + (((size_t)slot * a.P + p) * a.n_out_rows + u) * 32 + (l8 << 2)) =
                make_float4(acc[s].x * sc, acc[s].y * sc, acc[s].z * sc, acc[s].w * sc);
    }
}

// Backward products G_k = A_k^T dS for every relation of a group of many small relations: dS (all P
// panels) stays in shared memory for the whole kernel, each relation's rows come in their own sorted
// order and are written straight to HBM.  No barrier after the initial load.
constexpr int kAhead = 1;  // slots between the L2 prefetch of a row's optimizer state and its use
template <int P>
__global__ void __launch_bounds__(kStagedThreads, 1) spmm_tstaged_kernel(const TaskArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l8 = lane & 7, quarter = lane >> 3;
    const int slot = blockIdx.x;
    const int r_begin = a.slot_ptr[slot], n_rel = a.slot_ptr[slot + 1] - r_begin;
    const uint32_t panel_bytes = (uint32_t)a.n_op_rows * 128u;

    StreamReader rd;
    rd.init(smem_raw + (size_t)P * panel_bytes + (size_t)warp * kRingBytes,
            a.ent + (size_t)__ldg(a.wstart + slot * kTsWarps + warp) * 4, lane);
    int h_next = 0;
    if (n_rel > 0) {
        h_next = __ldg(a.hdr + ((size_t)a.slot_rel[r_begin] * kTsWarps + warp) * 4 + (lane & 3));
        rd.ensure(3);
    }
    {
        const float4 *src = reinterpret_cast<const float4 *>(a.op);
        float4 *dst = reinterpret_cast<float4 *>(smem_raw);
        const int n4 = P * a.n_op_rows * 8;
        for (int i = threadIdx.x; i < n4; i += kStagedThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const unsigned char *xrow = smem_raw + (l8 << 4);
    const size_t out_rows = (size_t)a.K * a.n_out_rows;
    float alpha = 0.f, omb1 = 0.f, omb2 = 0.f, eps = 0.f;
    if (a.adam_p != nullptr) alpha = a.dyn->alpha, omb1 = a.dyn->omb1, omb2 = a.dyn->omb2, eps = a.dyn->eps;

    for (int t = 0; t < n_rel; ++t) {
        const int k = a.slot_rel[r_begin + t];
        const int h = h_next;
        const int *__restrict__ orow = a.orow + (size_t)k * a.orow_stride + warp * a.rpq * 4 + quarter;
        if (t + 1 < n_rel) h_next = __ldg(a.hdr + ((size_t)a.slot_rel[r_begin + t + 1] * kTsWarps + warp) * 4 + (lane & 3));
        for (int s = 0; s < a.rpq; ++s) {
            const int n2 = task_count(h, s);
            const int c = __ldg(orow + s * 4);
            if (a.adam_p != nullptr) {
                // the optimizer state of a row is needed right after its gather: start moving it to L2 kAhead slots
                // earlier (a whole-relation bulk prefetch was measured: 549 us instead of 425, it thrashes L2)
                auto prefetch_row = [&](int cc) {
                    if (cc < 0) return;
#pragma unroll
                    for (int pp = 0; pp < P; ++pp) {
                        const size_t o = ((size_t)pp * out_rows + (size_t)k * a.n_out_rows + cc) * 32 + (l8 << 2);
                        prefetch_l2(a.adam_p + o);
                        prefetch_l2(a.adam_m + o);
                        prefetch_l2(a.adam_v + o);
                    }
                };
                if (s < kAhead) prefetch_row(c);
                if (kAhead > 0 && s + kAhead < a.rpq) prefetch_row(__ldg(orow + (s + kAhead) * 4));
            }
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
            stream_steps(rd, n2, [&](int off, int vbits) {
                gather_fma(s0, xrow, off, vbits);
                if constexpr (P > 1) gather_fma(s1, xrow + panel_bytes, off, vbits);
                if constexpr (P > 2) gather_fma(s2, xrow + 2 * panel_bytes, off, vbits);
                if constexpr (P > 2) gather_fma(s3, xrow + 3 * panel_bytes, off, vbits);
            });
            const float4 sum[4] = {s0, s1, s2, s3};
            if (c >= 0) {
                float sc = 1.f;
                if (a.mask != nullptr) sc = mask_bit(a.mask, k * a.n_out_rows + c) ? a.scale : 0.f;
#pragma unroll
                for (int pp = 0; pp < P; ++pp) {
                    const size_t o = ((size_t)pp * out_rows + (size_t)k * a.n_out_rows + c) * 32 + (l8 << 2);
                    const float4 gr = make_float4(sum[pp].x * sc, sum[pp].y * sc, sum[pp].z * sc, sum[pp].w * sc);
                    if (a.adam_p == nullptr) {
                        *reinterpret_cast<float4 *>(a.out + o) = gr;
                    } else {
                        // TF 1.8 ApplyAdam, the same arithmetic as adam_kernel (node.cu)
                        float4 P4 = *reinterpret_cast<const float4 *>(a.adam_p + o), M4 = *reinterpret_cast<const float4 *>(a.adam_m + o);
                        float4 V4 = *reinterpret_cast<const float4 *>(a.adam_v + o);
                        const float gg[4] = {gr.x, gr.y, gr.z, gr.w};
                        float *pq = &P4.x, *mq = &M4.x, *vq = &V4.x;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            adam_update(pq[q], mq[q], vq[q], gg[q], alpha, omb1, omb2, eps);
                        }
                        *reinterpret_cast<float4 *>(a.adam_p + o) = P4;
                        *reinterpret_cast<float4 *>(a.adam_m + o) = M4;
                        *reinterpret_cast<float4 *>(a.adam_v + o) = V4;
                    }
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace

// ------------------------------------------------------------------------------ launchers
void launch_spmm(const SpmmArgs &a, int P, cudaStream_t s) {
    if (a.n_seg == 0) return;
    const int segs_per_block = 32;  // 256 threads = 32 quarter-warps
    dim3 grid((unsigned)((a.n_seg + segs_per_block - 1) / segs_per_block)), block(256);
    switch (P) {
        case 1: spmm_seg_kernel<1><<<grid, block, 0, s>>>(a); break;
        case 2: spmm_seg_kernel<2><<<grid, block, 0, s>>>(a); break;
        case 4: spmm_seg_kernel<4><<<grid, block, 0, s>>>(a); break;
        default: DGN_FAIL(DGN_ERR_UNSUPPORTED, "spmm: %d panels (hidden sizes must be 32, 64 or 128)", P);
    }
    CUDA_CHECK(cudaGetLastError());
}

void launch_seg_reduce(const SpmmArgs &a, const int *multi_rows, int n_multi, int P, cudaStream_t s) {
    if (n_multi == 0) return;
    const int warps_per_block = 8;
    dim3 grid((unsigned)((n_multi + warps_per_block - 1) / warps_per_block)), block(warps_per_block * 32);
    switch (P) {
        case 1: seg_reduce_kernel<1><<<grid, block, 0, s>>>(a, multi_rows, n_multi); break;
        case 2: seg_reduce_kernel<2><<<grid, block, 0, s>>>(a, multi_rows, n_multi); break;
        case 4: seg_reduce_kernel<4><<<grid, block, 0, s>>>(a, multi_rows, n_multi); break;
        default: DGN_FAIL(DGN_ERR_UNSUPPORTED, "seg_reduce: %d panels", P);
    }
    CUDA_CHECK(cudaGetLastError());
}

size_t staged_smem_bytes(int n_j) { return (size_t)2 * n_j * 32 * sizeof(float); }

bool staged_supported(int n_i, int n_j, int K) {
    return K >= 8 && staged_smem_bytes(n_j) <= 200 * 1024 && n_i <= 8 * kQuarters;
}

template <int RPQ>
static void launch_staged_t(const StagedArgs &a, cudaStream_t s) {
    const size_t smem = staged_smem_bytes(a.n_j);
    static PerDeviceOnce configured;
    if (configured.first()) {
        CUDA_CHECK(cudaFuncSetAttribute(spmm_staged_kernel<RPQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        200 * 1024));
    }
    spmm_staged_kernel<RPQ><<<a.n_slots * a.P, kStagedThreads, smem, s>>>(a);
    CUDA_CHECK(cudaGetLastError());
}

void launch_spmm_staged(const StagedArgs &a, cudaStream_t s) {
    const int rpq = (a.n_i + kQuarters - 1) / kQuarters;  // rows per quarter-warp
    if (rpq <= 2) launch_staged_t<2>(a, s);
    else if (rpq <= 4) launch_staged_t<4>(a, s);
    else if (rpq <= 6) launch_staged_t<6>(a, s);
    else if (rpq <= 8) launch_staged_t<8>(a, s);
    else DGN_FAIL(DGN_ERR_UNSUPPORTED, "staged spmm: %d rows per quarter-warp", rpq);
}

constexpr size_t kMaxDynSmem = 227 * 1024 - 256;  // per-CTA limit minus the static barriers
static size_t staged3_smem(int n_j) { return staged_smem_bytes(n_j) + (size_t)kS3Warps * kRingBytes; }
static size_t tstaged_smem(int n_i, int P) { return (size_t)P * n_i * 128 + (size_t)kTsWarps * kRingBytes; }

bool staged3_supported(int n_i, int n_j, int K) { return K >= 8 && staged3_smem(n_j) <= kMaxDynSmem && n_i <= 8 * 4 * kS3Warps; }

bool tstaged_supported(int n_i, int n_j, int K, int P) {
    return K >= 8 && tstaged_smem(n_i, P) <= kMaxDynSmem && n_j <= 8 * 4 * kTsWarps && P <= 4;
}

template <int RPQ>
static void launch_staged3_t(const TaskArgs &a, cudaStream_t s) {
    static PerDeviceOnce configured;
    if (configured.first()) {
        CUDA_CHECK(cudaFuncSetAttribute(spmm_staged3_kernel<RPQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
    }
    spmm_staged3_kernel<RPQ><<<a.n_slots * a.P, kStagedThreads, staged3_smem(a.n_op_rows), s>>>(a);
    CUDA_CHECK(cudaGetLastError());
}

void launch_spmm_staged3(const TaskArgs &a, cudaStream_t s) {
    if (a.rpq <= 2) launch_staged3_t<2>(a, s);
    else if (a.rpq <= 4) launch_staged3_t<4>(a, s);
    else if (a.rpq <= 6) launch_staged3_t<6>(a, s);
    else if (a.rpq <= 8) launch_staged3_t<8>(a, s);
    else DGN_FAIL(DGN_ERR_UNSUPPORTED, "staged spmm: %d rows per quarter-warp", a.rpq);
}

template <int P>
static void launch_tstaged_t(const TaskArgs &a, cudaStream_t s) {
    static PerDeviceOnce configured;
    if (configured.first()) {
        CUDA_CHECK(cudaFuncSetAttribute(spmm_tstaged_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
    }
    spmm_tstaged_kernel<P><<<a.n_slots, kStagedThreads, tstaged_smem(a.n_op_rows, P), s>>>(a);
    CUDA_CHECK(cudaGetLastError());
}

void launch_spmm_tstaged(const TaskArgs &a, cudaStream_t s) {
    switch (a.P) {
        case 1: launch_tstaged_t<1>(a, s); break;
        case 2: launch_tstaged_t<2>(a, s); break;
        case 4: launch_tstaged_t<4>(a, s); break;
        default: DGN_FAIL(DGN_ERR_UNSUPPORTED, "transposed staged spmm: %d panels", a.P);
    }
}

}  // namespace dgn
