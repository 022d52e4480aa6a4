// Row-local pieces of the encoder and the optimizer update:
//   node_epilogue_kernel -- ordered sum of the SpMM partials of every (i, j) group, l2-normalise
//                           per group (tf.nn.l2_normalize, layers.py:93,117), sum over the groups of
//                           a node type and optional ReLU (model.py:74-75 / :86-88)
//   l2norm_bwd_kernel    -- its backward per group
//   relu_bwd_kernel      -- ordered sum of the dH partials + ReLU mask
//   gen_mask_kernel      -- Philox4x32-10 dropout keep-bits (layers.py:23-31, :112)
//   adam_kernel          -- TF-1.8 ApplyAdam (optimizer.py:111-113)
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "dgn_internal.cuh"
#include "philox.cuh"

namespace dgn {
namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
    return x;
}

template <int P>
__global__ void __launch_bounds__(256) node_epilogue_kernel(const EpiArgs a) {
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (row >= a.n_rows) return;
    float total[P];
#pragma unroll
    for (int p = 0; p < P; ++p) total[p] = 0.f;
    for (int gi = 0; gi < a.n_groups; ++gi) {
        const EpiGroup &g = a.g[gi];
        float s[P];
#pragma unroll
        for (int p = 0; p < P; ++p) s[p] = 0.f;
        if (g.n_peers > 0) {
            for (int r = 0; r < g.n_peers; ++r)
#pragma unroll
                for (int p = 0; p < P; ++p) s[p] += g.peer[r][((size_t)p * a.n_rows + row) * 32 + lane];
        } else if (g.row_seg_ptr != nullptr) {
            const int s0 = g.row_seg_ptr[row], s1 = g.row_seg_ptr[row + 1];
            float b1[P], b2[P], b3[P];  // four interleaved partial sums: hub rows span many segments
#pragma unroll
            for (int p = 0; p < P; ++p) b1[p] = b2[p] = b3[p] = 0.f;
            int sg = s0;
            for (; sg + 4 <= s1; sg += 4)
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    s[p] += g.partial[((size_t)sg * P + p) * 32 + lane];
                    b1[p] += g.partial[((size_t)(sg + 1) * P + p) * 32 + lane];
                    b2[p] += g.partial[((size_t)(sg + 2) * P + p) * 32 + lane];
                    b3[p] += g.partial[((size_t)(sg + 3) * P + p) * 32 + lane];
                }
            for (; sg < s1; ++sg)
#pragma unroll
                for (int p = 0; p < P; ++p) s[p] += g.partial[((size_t)sg * P + p) * 32 + lane];
#pragma unroll
            for (int p = 0; p < P; ++p) s[p] = (s[p] + b1[p]) + (b2[p] + b3[p]);
        } else {
            for (int sl = 0; sl < g.n_slots; ++sl)
#pragma unroll
                for (int p = 0; p < P; ++p) s[p] += g.partial[(((size_t)sl * P + p) * a.n_rows + row) * 32 + lane];
        }
        float sq = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) sq = fmaf(s[p], s[p], sq);
        sq = warp_sum(sq);
        const float nrm = sqrtf(fmaxf(sq, kL2Eps));
        const float inv = 1.f / nrm;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float y = s[p] * inv;
            g.Y[((size_t)p * a.n_rows + row) * 32 + lane] = y;
            total[p] += y;
        }
        if (lane == 0) g.nrm[row] = nrm;
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
        a.out[((size_t)p * a.n_rows + row) * 32 + lane] = a.relu ? fmaxf(total[p], 0.f) : total[p];
}

// Same result layout for node types with few rows whose groups carry many slot partials (645 drugs x 74..148
// slots): one CTA per row, the 8 warps sum contiguous chunks of the slots, the chunk sums are added in order.
template <int P>
__global__ void __launch_bounds__(256) node_epilogue_wide_kernel(const EpiArgs a) {
    __shared__ float red[8][P][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x;
    float total[P];
#pragma unroll
    for (int p = 0; p < P; ++p) total[p] = 0.f;
    for (int gi = 0; gi < a.n_groups; ++gi) {
        const EpiGroup &g = a.g[gi];
        float s[P];
#pragma unroll
        for (int p = 0; p < P; ++p) s[p] = 0.f;
        if (g.n_peers > 0) {
            for (int r = warp; r < g.n_peers; r += 8)
#pragma unroll
                for (int p = 0; p < P; ++p) s[p] += g.peer[r][((size_t)p * a.n_rows + row) * 32 + lane];
        } else if (g.row_seg_ptr != nullptr) {
            const int s0 = g.row_seg_ptr[row], s1 = g.row_seg_ptr[row + 1];
            const int per = (s1 - s0 + 7) / 8;
            for (int sg = s0 + warp * per; sg < min(s0 + (warp + 1) * per, s1); ++sg)
#pragma unroll
                for (int p = 0; p < P; ++p) s[p] += g.partial[((size_t)sg * P + p) * 32 + lane];
        } else {
            const int per = (g.n_slots + 7) / 8;
            for (int sl = warp * per; sl < min((warp + 1) * per, g.n_slots); ++sl)
#pragma unroll
                for (int p = 0; p < P; ++p) s[p] += g.partial[(((size_t)sl * P + p) * a.n_rows + row) * 32 + lane];
        }
#pragma unroll
        for (int p = 0; p < P; ++p) red[warp][p][lane] = s[p];
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float t = 0.f;
                if (g.n_peers > 0 && g.n_peers <= 8) {
                    for (int r = 0; r < g.n_peers; ++r) t += red[r][p][lane];  // rank order: every rank gets the same bits
                } else {
                    for (int w = 0; w < 8; ++w) t += red[w][p][lane];
                }
                s[p] = t;
            }
            float sq = 0.f;
#pragma unroll
            for (int p = 0; p < P; ++p) sq = fmaf(s[p], s[p], sq);
            sq = warp_sum(sq);
            const float nrm = sqrtf(fmaxf(sq, kL2Eps));
            const float inv = 1.f / nrm;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float y = s[p] * inv;
                g.Y[((size_t)p * a.n_rows + row) * 32 + lane] = y;
                total[p] += y;
            }
            if (lane == 0) g.nrm[row] = nrm;
        }
        __syncthreads();
    }
    if (warp == 0) {
#pragma unroll
        for (int p = 0; p < P; ++p)
            a.out[((size_t)p * a.n_rows + row) * 32 + lane] = a.relu ? fmaxf(total[p], 0.f) : total[p];
    }
}

// y = s / n, n = sqrt(max(|s|^2, eps)):  ds = (dy - y (y . dy)) / n; clamped rows: ds = dy / n
template <int P>
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const L2BwdArgs a) {
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (row >= a.n_rows) return;
    float y[P], dy[P], dot = 0.f;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const size_t o = ((size_t)p * a.n_rows + row) * 32 + lane;
        y[p] = a.Y[o];
        dy[p] = a.dY[o];
        dot = fmaf(y[p], dy[p], dot);
    }
    dot = warp_sum(dot);
    const float nrm = a.nrm[row];
    if (nrm * nrm <= kL2Eps * 1.0000001f) dot = 0.f;  // clamp active: y = s * rsqrt(eps) is linear in s
    const float inv = 1.f / nrm;
#pragma unroll
    for (int p = 0; p < P; ++p) a.dS[((size_t)p * a.n_rows + row) * 32 + lane] = (dy[p] - y[p] * dot) * inv;
}

template <int P>
__global__ void __launch_bounds__(256) relu_bwd_kernel(const ReluBwdArgs a) {
    const size_t n = (size_t)P * a.n_rows * 32;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int gi = 0; gi < a.n_groups; ++gi) {
            if (a.g[gi].n_peers > 0)
                for (int r = 0; r < a.g[gi].n_peers; ++r) s += a.g[gi].peer[r][i];
            else
                for (int c = 0; c < a.g[gi].n_chunks; ++c) s += a.g[gi].part[(size_t)c * n + i];
        }
        a.dA[i] = a.H[i] > 0.f ? s : 0.f;
    }
}

// One thread per output word.  Bit b of relation `rel` is element e = b of that relation's
// stream: word (e & 3) of Philox(counter = (e >> 2, relation, stream, step), key = seed).
// words_per_rel == 0: relations are packed back to back at bit granularity (layer 1, bit index
// = rel * bits_per_rel + e); otherwise each relation owns words_per_rel whole words (layer 2).
__device__ __forceinline__ void mask_word(long long w, uint32_t *__restrict__ words, long long bits_per_rel, int words_per_rel,
                                          const int *__restrict__ rel_ids, uint32_t stream_id, uint32_t step, uint32_t seed_lo,
                                          uint32_t seed_hi, uint32_t threshold, long long total_bits) {
    uint32_t out = 0;
    const uint2 key = make_uint2(seed_lo, seed_hi);
    if (words_per_rel > 0) {
        // word-aligned relations: 8 Philox calls give the 32 bits of this word
        const long long rel = w / words_per_rel;
        const long long e0 = (w - rel * words_per_rel) * 32;
        const uint32_t rid = (uint32_t)rel_ids[rel];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const long long e = e0 + 4 * q;
            if (e >= bits_per_rel) break;
            const uint4 rnd = philox4x32_10(make_uint4((uint32_t)(e >> 2), rid, stream_id, step), key);
            out |= (rnd.x >= threshold ? 1u : 0u) << (4 * q);
            if (e + 1 < bits_per_rel) out |= (rnd.y >= threshold ? 1u : 0u) << (4 * q + 1);
            if (e + 2 < bits_per_rel) out |= (rnd.z >= threshold ? 1u : 0u) << (4 * q + 2);
            if (e + 3 < bits_per_rel) out |= (rnd.w >= threshold ? 1u : 0u) << (4 * q + 3);
        }
        words[w] = out;
        return;
    }
    // packed relations: walk the 32 bits of the word, one division per word instead of two per bit
    long long rel = (w * 32) / bits_per_rel, e = (w * 32) % bits_per_rel;
    long long cached_ctr = -1;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    uint32_t rid = (uint32_t)rel_ids[rel];
    for (int b = 0; b < 32; ++b) {
        if (w * 32 + b >= total_bits) break;
        const long long ctr = e >> 2;
        if (ctr != cached_ctr) {
            rnd = philox4x32_10(make_uint4((uint32_t)ctr, rid, stream_id, step), key);
            cached_ctr = ctr;
        }
        const uint32_t u = (e & 3) == 0 ? rnd.x : (e & 3) == 1 ? rnd.y : (e & 3) == 2 ? rnd.z : rnd.w;
        if (u >= threshold) out |= 1u << b;
        if (++e == bits_per_rel) {
            e = 0, ++rel, cached_ctr = -1;
            rid = (uint32_t)rel_ids[rel];  // rel_ids carries 33 spare entries: a word spans at most 32 relation boundaries
        }
    }
    words[w] = out;
}

// Grid-stride: the grid is capped at a few CTAs per SM so that the kernel (integer ALU bound) leaves thread and
// register slots to the L2-bound gather kernels of the other stream lane that run beside it.
__global__ void __launch_bounds__(256) gen_mask_kernel(uint32_t *__restrict__ words, long long n_words, long long bits_per_rel,
                                                       int words_per_rel, const int *__restrict__ rel_ids, uint32_t stream_id,
                                                       const StepDyn *__restrict__ dyn, uint32_t step_offset, long long total_bits) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const uint32_t step = dyn->step + step_offset, seed_lo = dyn->seed_lo, seed_hi = dyn->seed_hi, threshold = dyn->threshold;
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < n_words; w += stride)
        mask_word(w, words, bits_per_rel, words_per_rel, rel_ids, stream_id, step, seed_lo, seed_hi, threshold, total_bits);
}

__global__ void __launch_bounds__(256) adam_kernel(float *__restrict__ p, const float *__restrict__ g,
                                                   float *__restrict__ m, float *__restrict__ v, long long n,
                                                   const StepDyn *__restrict__ dyn) {
    const float alpha = dyn->alpha, omb1 = dyn->omb1, omb2 = dyn->omb2, eps = dyn->eps;
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P4 = reinterpret_cast<float4 *>(p)[i], G4 = reinterpret_cast<const float4 *>(g)[i];
        float4 M4 = reinterpret_cast<float4 *>(m)[i], V4 = reinterpret_cast<float4 *>(v)[i];
        float *pp = &P4.x, *gg = &G4.x, *mm = &M4.x, *vv = &V4.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            adam_update(pp[q], mm[q], vv[q], gg[q], alpha, omb1, omb2, eps);
        }
        reinterpret_cast<float4 *>(p)[i] = P4;
        reinterpret_cast<float4 *>(m)[i] = M4;
        reinterpret_cast<float4 *>(v)[i] = V4;
    }
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        float mi = m[i], vi = v[i];
        const float gi = g[i];
        float pi = p[i];
        adam_update(pi, mi, vi, gi, alpha, omb1, omb2, eps);
        p[i] = pi;
        m[i] = mi;
        v[i] = vi;
    }
}

// ---- multi-GPU exchange -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) publish_kernel(const float4 *__restrict__ partial, int n_chunks, size_t n4,
                                                      float4 *__restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < n_chunks; ++c) {
            const float4 x = partial[(size_t)c * n4 + i];
            s.x += x.x, s.y += x.y, s.z += x.z, s.w += x.w;
        }
        out[i] = s;
    }
}

// One CTA, thread t talks to rank t.  The exchange buffer was written by the previous kernel of this
// stream; peers read it through NVLink after they see the stamp.
__global__ void signal_wait_kernel(uint32_t *const *__restrict__ peer_flags, uint32_t *my_flags, int rank, int world, int x,
                                   const StepDyn *__restrict__ dyn, unsigned long long timeout_ns, int *error_flag) {
    const int t = threadIdx.x;
    if (t >= world) return;
    const uint32_t stamp = dyn->stamp[x];
    __threadfence_system();
    volatile uint32_t *dst = peer_flags[t] + rank * kMaxExchanges + x;
    *dst = stamp;
    volatile uint32_t *src = my_flags + t * kMaxExchanges + x;
    // Ranks drift apart on the host (logging, evaluation on rank 0, a re-finalize): poll with a back-off and give
    // up by WALL CLOCK, not by iteration count.  A peer that never arrives must not hang the GPU and must not kill
    // the context either: the error flag is raised, the kernel returns and the host reports DGN_ERR_CUDA at its next
    // synchronisation point (dgn_train_step with loss_out, dgn_sync).
    unsigned long long t0 = 0;
    unsigned ns = 32;
    for (long long spin = 0; (int)(*src - stamp) < 0; ++spin) {
        if (spin < 64) continue;
        if (t0 == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        __nanosleep(ns);
        if (ns < 2048) ns <<= 1;
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > timeout_ns) {
            atomicExch(error_flag, 1 + t);
            break;
        }
    }
    __threadfence_system();
}

template <typename F>
void dispatch_panels(int P, F &&f) {
    switch (P) {
        case 1: f(std::integral_constant<int, 1>{}); break;
        case 2: f(std::integral_constant<int, 2>{}); break;
        case 4: f(std::integral_constant<int, 4>{}); break;
        default: DGN_FAIL(DGN_ERR_UNSUPPORTED, "%d panels (hidden sizes must be 32, 64 or 128)", P);
    }
}

}  // namespace

void launch_node_epilogue(const EpiArgs &a, int P, cudaStream_t s) {
    if (a.n_rows == 0) return;
    int max_terms = 0;  // longest ordered sum of a row
    for (int gi = 0; gi < a.n_groups; ++gi)
        max_terms = std::max(max_terms, a.g[gi].n_peers > 0 ? 1 : a.g[gi].row_seg_ptr != nullptr ? 1 : a.g[gi].n_slots);
    if (max_terms >= 16 && a.n_rows <= 4096) {
        dispatch_panels(P, [&](auto tag) { node_epilogue_wide_kernel<decltype(tag)::value><<<a.n_rows, 256, 0, s>>>(a); });
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    dim3 grid((unsigned)((a.n_rows + 7) / 8)), block(256);
    dispatch_panels(P, [&](auto tag) { node_epilogue_kernel<decltype(tag)::value><<<grid, block, 0, s>>>(a); });
    CUDA_CHECK(cudaGetLastError());
}

void launch_l2norm_bwd(const L2BwdArgs &a, int P, cudaStream_t s) {
    if (a.n_rows == 0) return;
    dim3 grid((unsigned)((a.n_rows + 7) / 8)), block(256);
    dispatch_panels(P, [&](auto tag) { l2norm_bwd_kernel<decltype(tag)::value><<<grid, block, 0, s>>>(a); });
    CUDA_CHECK(cudaGetLastError());
}

void launch_relu_bwd(const ReluBwdArgs &a, int P, cudaStream_t s) {
    if (a.n_rows == 0) return;
    const size_t n = (size_t)P * a.n_rows * 32;
    dim3 grid((unsigned)std::min<size_t>((n + 255) / 256, 148 * 8)), block(256);
    dispatch_panels(P, [&](auto tag) { relu_bwd_kernel<decltype(tag)::value><<<grid, block, 0, s>>>(a); });
    CUDA_CHECK(cudaGetLastError());
}

void launch_publish(const float *partial, int n_chunks, size_t floats, float *out, cudaStream_t s) {
    const size_t n4 = floats / 4;
    if (n4 == 0) return;
    dim3 grid((unsigned)std::min<size_t>((n4 + 255) / 256, 148 * 4)), block(256);
    publish_kernel<<<grid, block, 0, s>>>(reinterpret_cast<const float4 *>(partial), n_chunks, n4, reinterpret_cast<float4 *>(out));
    CUDA_CHECK(cudaGetLastError());
}

void launch_signal_wait(uint32_t *const *peer_flags_dev, uint32_t *my_flags, int rank, int world, int x, const StepDyn *dyn,
                        unsigned long long timeout_ns, int *error_flag, cudaStream_t s) {
    signal_wait_kernel<<<1, 32, 0, s>>>(peer_flags_dev, my_flags, rank, world, x, dyn, timeout_ns, error_flag);
    CUDA_CHECK(cudaGetLastError());
}

// layer-1 keep bits (packed mode) of several groups in one launch: blockIdx.y = group
__global__ void __launch_bounds__(256) gen_mask_multi_kernel(const MaskBatch mb, uint32_t stream_id, const StepDyn *__restrict__ dyn) {
    const int q = blockIdx.y;
    const uint32_t step = dyn->step, seed_lo = dyn->seed_lo, seed_hi = dyn->seed_hi, threshold = dyn->threshold;
    const long long n_words = mb.n_words[q], stride = (long long)gridDim.x * blockDim.x;
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < n_words; w += stride)
        mask_word(w, mb.words[q], mb.bits_per_rel[q], 0, mb.rel_ids[q], stream_id, step, seed_lo, seed_hi, threshold, n_words * 32);
}

void launch_gen_mask_multi(const MaskBatch &mb, uint32_t stream_id, const StepDyn *dyn, cudaStream_t s) {
    if (mb.n == 0) return;
    long long most = 0;
    for (int q = 0; q < mb.n; ++q) most = std::max(most, mb.n_words[q]);
    if (most == 0) return;
    dim3 grid((unsigned)std::min<long long>((most + 255) / 256, 148 * 3), (unsigned)mb.n), block(256);
    gen_mask_multi_kernel<<<grid, block, 0, s>>>(mb, stream_id, dyn);
    CUDA_CHECK(cudaGetLastError());
}

void launch_gen_mask(uint32_t *words, long long n_words, long long bits_per_rel, int words_per_rel, const int *rel_ids,
                     uint32_t stream_id, const StepDyn *dyn, uint32_t step_offset, cudaStream_t s) {
    if (n_words == 0) return;
    const long long total_bits = n_words * 32;  // packed mode: caller rounds the word count up
    static int ctas_per_sm = 0;  // DGN_MASK_CTAS: CTAs per SM the layer-2 mask kernel may occupy (default 3)
    if (ctas_per_sm == 0) {
        const char *e = getenv("DGN_MASK_CTAS");
        ctas_per_sm = e && atoi(e) > 0 ? atoi(e) : 3;
    }
    dim3 grid((unsigned)std::min<long long>((n_words + 255) / 256, 148LL * ctas_per_sm)), block(256);
    gen_mask_kernel<<<grid, block, 0, s>>>(words, n_words, bits_per_rel, words_per_rel, rel_ids, stream_id, dyn, step_offset, total_bits);
    CUDA_CHECK(cudaGetLastError());
}

void launch_adam(float *p, const float *g, float *m, float *v, long long n, const StepDyn *dyn, cudaStream_t s) {
    if (n == 0) return;
    dim3 grid(148 * 8), block(256);
    adam_kernel<<<grid, block, 0, s>>>(p, g, m, v, n, dyn);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace dgn
