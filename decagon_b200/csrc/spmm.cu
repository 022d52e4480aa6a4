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
template <int P>
__global__ void __launch_bounds__(256) spmm_seg_kernel(const SpmmArgs a) {
    const int lane = threadIdx.x & 31;
    const int seg = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (seg >= a.n_seg) return;

    int row, begin, end;
    bool single = true;
    if (a.seg_row != nullptr) {
        row = __ldg(a.seg_row + seg);
        begin = __ldg(a.seg_begin + seg);
        end = min(begin + a.seg_len, __ldg(a.rowptr + row + 1));
        single = (__ldg(a.row_seg_ptr + row + 1) - __ldg(a.row_seg_ptr + row)) == 1;
    } else {
        row = seg;
        begin = __ldg(a.rowptr + row);
        end = __ldg(a.rowptr + row + 1);
    }

    float acc[P];
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = 0.f;

    const float *__restrict__ op = a.op + lane;
    for (int base = begin; base < end; base += 32) {
        const int idx = base + lane;
        int c = 0;
        float v = 0.f;
        if (idx < end) {
            c = __ldg(a.col + idx);
            v = __ldg(a.val + idx);
            if (a.col_mask) v = mask_bit(a.mask, c) ? v * a.scale : 0.f;
        }
        const int n = min(32, end - base);
        // padded to a multiple of 4: the padding lanes hold (col 0, value 0)
        for (int t = 0; t < n; t += 4) {
            int cc[4];
            float vv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                cc[q] = __shfl_sync(kFull, c, (t + q) & 31);
                vv[q] = __shfl_sync(kFull, v, (t + q) & 31);
            }
            float x[4][P];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int p = 0; p < P; ++p) x[q][p] = __ldg(op + ((size_t)p * a.op_rows + cc[q]) * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int p = 0; p < P; ++p) acc[p] = fmaf(vv[q], x[q][p], acc[p]);
        }
    }

    if (a.force_partial || !single) {
#pragma unroll
        for (int p = 0; p < P; ++p) a.partial[((size_t)seg * P + p) * 32 + lane] = acc[p];
    } else {
        float s = 1.f;
        if (a.row_mask) s = mask_bit(a.mask, row) ? a.scale : 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) a.out[((size_t)p * a.out_rows + row) * 32 + lane] = acc[p] * s;
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
    float acc[P];
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = 0.f;
    for (int s = s0; s < s1; ++s)
#pragma unroll
        for (int p = 0; p < P; ++p) acc[p] += a.partial[((size_t)s * P + p) * 32 + lane];
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

template <int RPW>
__global__ void __launch_bounds__(kStagedThreads, 1) spmm_staged_kernel(const StagedArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tile_floats = a.n_j * 32;
    float *tile[2] = {reinterpret_cast<float *>(smem_raw), reinterpret_cast<float *>(smem_raw) + tile_floats};
    __shared__ __align__(8) uint64_t full[2];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.x % a.P, slot = blockIdx.x / a.P;
    const int r_begin = a.slot_ptr[slot], r_end = a.slot_ptr[slot + 1];
    const uint32_t tile_bytes = (uint32_t)tile_floats * 4u;

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int t) {  // thread 0 only
        const int k = a.slot_rel[r_begin + t];
        uint64_t *bar = &full[t & 1];
        mbar_expect_tx(bar, tile_bytes);
        bulk_load(tile[t & 1], a.op + ((size_t)p * a.K + k) * tile_floats, tile_bytes, bar);
    };
    const int n_rel = r_end - r_begin;
    if (threadIdx.x == 0 && n_rel > 0) issue(0);

    float acc[RPW];
#pragma unroll
    for (int s = 0; s < RPW; ++s) acc[s] = 0.f;

    for (int t = 0; t < n_rel; ++t) {
        // buffer (t+1)&1 was last read for relation t-1; every warp passed the barrier that
        // closed iteration t-1, so it may be overwritten now
        if (threadIdx.x == 0 && t + 1 < n_rel) issue(t + 1);
        const int k = a.slot_rel[r_begin + t];
        const int *__restrict__ rp = a.rowptr + (size_t)k * (a.n_i + 1);
        // lane s of this warp fetches the row pointers of the warp's s-th row
        int my_b = 0, my_e = 0;
        {
            const int u = warp + lane * kStagedWarps;
            if (lane < RPW && u < a.n_i) {
                my_b = __ldg(rp + u);
                my_e = __ldg(rp + u + 1);
            }
        }
        mbar_wait(&full[t & 1], (uint32_t)((t >> 1) & 1));
        const float *__restrict__ x = tile[t & 1] + lane;
        const int mask_base = k * a.n_j;
#pragma unroll
        for (int s = 0; s < RPW; ++s) {
            const int begin = __shfl_sync(kFull, my_b, s), end = __shfl_sync(kFull, my_e, s);
            float sum = 0.f;
            for (int base = begin; base < end; base += 32) {
                const int idx = base + lane;
                int c = 0;
                float v = 0.f;
                if (idx < end) {
                    c = __ldg(a.col + idx);
                    v = __ldg(a.val + idx);
                    if (a.mask != nullptr) v = mask_bit(a.mask, mask_base + c) ? v * a.scale : 0.f;
                }
                const int n = min(32, end - base);
                for (int q = 0; q < n; q += 4) {
                    const int c0 = __shfl_sync(kFull, c, q), c1 = __shfl_sync(kFull, c, (q + 1) & 31);
                    const int c2 = __shfl_sync(kFull, c, (q + 2) & 31), c3 = __shfl_sync(kFull, c, (q + 3) & 31);
                    const float v0 = __shfl_sync(kFull, v, q), v1 = __shfl_sync(kFull, v, (q + 1) & 31);
                    const float v2 = __shfl_sync(kFull, v, (q + 2) & 31), v3 = __shfl_sync(kFull, v, (q + 3) & 31);
                    const float x0 = x[c0 * 32], x1 = x[c1 * 32], x2 = x[c2 * 32], x3 = x[c3 * 32];
                    sum = fmaf(v0, x0, sum);
                    sum = fmaf(v1, x1, sum);
                    sum = fmaf(v2, x2, sum);
                    sum = fmaf(v3, x3, sum);
                }
            }
            acc[s] += sum;
        }
        __syncthreads();
    }

#pragma unroll
    for (int s = 0; s < RPW; ++s) {
        const int u = warp + s * kStagedWarps;
        if (u < a.n_i) a.partial[(((size_t)slot * a.P + p) * a.n_i + u) * 32 + lane] = acc[s];
    }
}

}  // namespace

// ------------------------------------------------------------------------------ launchers
void launch_spmm(const SpmmArgs &a, int P, cudaStream_t s) {
    if (a.n_seg == 0) return;
    const int warps_per_block = 8;
    dim3 grid((unsigned)((a.n_seg + warps_per_block - 1) / warps_per_block)), block(warps_per_block * 32);
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
    return K >= 8 && staged_smem_bytes(n_j) <= 200 * 1024 && n_i <= 32 * kStagedWarps;
}

template <int RPW>
static void launch_staged_t(const StagedArgs &a, cudaStream_t s) {
    const size_t smem = staged_smem_bytes(a.n_j);
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(spmm_staged_kernel<RPW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        200 * 1024));
        configured = true;
    }
    spmm_staged_kernel<RPW><<<a.n_slots * a.P, kStagedThreads, smem, s>>>(a);
    CUDA_CHECK(cudaGetLastError());
}

void launch_spmm_staged(const StagedArgs &a, int rows_per_warp, cudaStream_t s) {
    if (rows_per_warp <= 8) launch_staged_t<8>(a, s);
    else if (rows_per_warp <= 16) launch_staged_t<16>(a, s);
    else if (rows_per_warp <= 24) launch_staged_t<24>(a, s);
    else if (rows_per_warp <= 32) launch_staged_t<32>(a, s);
    else DGN_FAIL(DGN_ERR_UNSUPPORTED, "staged spmm: %d rows per warp", rows_per_warp);
}

}  // namespace dgn
