// Minibatch edge decoding fused with the loss and its backward, and all-pairs scoring.
//
// Reference: DecagonOptimizer.batch_predict (decagon/deep/optimizer.py:63-85) builds a B x B
// score matrix per call and keeps its diagonal (:52,:56); negatives come from
// fixed_unigram_candidate_sampler (:36-49) and replace the ROW node; loss is the hinge
// (:116-120, active) or sigmoid cross-entropy (:122-127); DecagonOptimizer.predict (:87-106)
// is the all-pairs form.  Decoder parameter matrices: model.py:116-137.
//
// decode_kernel: ONE CTA (the whole batch is 512 edges x 32 features = 64 KB of gathers, far
// below one SM's bandwidth; a single CTA keeps the loss and the decoder-parameter gradient
// sums ordered, hence reproducible).  One warp per edge, lane = embedding feature:
//   M = loc glb loc;  a = M z_v;  s+ = z_u . a;  s- = z_n . a
//   dZ_i[u] += ds+ a;  dZ_i[n] += ds- a;  dZ_j[v] += M^T (ds+ z_u + ds- z_n)   (float atomics)
//   dM += (ds+ z_u + ds- z_n) z_v^T   (register tile per warp, ordered reduction over warps)
#include <algorithm>

#include "dgn_internal.cuh"
#include "philox.cuh"

namespace dgn {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int D = 32;  // hidden2
constexpr int kDecodeThreads = 512;

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
    return x;
}

// M[p][q] of relation with parameters (glb, loc), model.py:116-137
__device__ __forceinline__ float relation_entry(int decoder, const float *glb, const float *loc, int p, int q) {
    switch (decoder) {
        case DGN_DEC_INNERPRODUCT: return p == q ? 1.f : 0.f;
        case DGN_DEC_DISTMULT: return p == q ? loc[p] : 0.f;
        case DGN_DEC_BILINEAR: return loc[p * D + q];
        default: return loc[p] * glb[p * D + q] * loc[q];  // dedicom
    }
}

constexpr float kFixedScale = 1099511627776.f;  // 2^40
constexpr float kFixedLimit = 4194304.f;        // 2^22: |x| below it keeps 2^40 x (and sums of 2 * B of them) inside int64
// Returns false when x cannot be represented (too large, inf or NaN): the caller turns the step's loss into NaN,
// which is what the reference's float arithmetic would show for a diverged model, instead of a finite, wrong dZ.
__device__ __forceinline__ bool fixed_add(long long *dst, float x) {
    atomicAdd(reinterpret_cast<unsigned long long *>(dst), (unsigned long long)__float2ll_rn(x * kFixedScale));
    return fabsf(x) < kFixedLimit;
}
__global__ void fixed_to_float_kernel(const long long *__restrict__ q, float *__restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (float)((double)q[i] * (1.0 / 1099511627776.0));
}

// Grid of kDecodeCtas CTAs, each takes a contiguous slice of the batch (one warp per edge); the last CTA to
// finish (ticket counter) adds the CTAs' loss and dM partials in CTA order, so the sums have a fixed order.
__global__ void __launch_bounds__(kDecodeThreads, 1) decode_kernel(const StepDyn *__restrict__ dyn) {
    const DecodeArgs a = dyn->dec;  // per-step arguments live in device memory (CUDA-graph replay)
    __shared__ float Ms[D][D + 1];
    __shared__ float dMs[D][D + 1];
    __shared__ float loss_w[kDecodeThreads / 32];
    __shared__ int is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int per = (a.B + gridDim.x - 1) / gridDim.x;
    const int b0 = blockIdx.x * per, b1 = min(b0 + per, a.B);

    for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
        const int p = i >> 5, q = i & 31;
        Ms[p][q] = relation_entry(a.decoder, a.glb, a.loc, p, q);
        dMs[p][q] = 0.f;
    }
    // negatives: index = #{v : thr[v] <= u32}, clamped (optimizer.py:40-47 restated)
    for (int b = b0 + threadIdx.x; b < b1; b += blockDim.x) {
        long long neg;
        if (a.neg_in != nullptr) {
            neg = a.neg_in[b];
        } else {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)(b >> 2), a.relation, kStreamNegatives, a.step),
                                          make_uint2(a.seed_lo, a.seed_hi));
            const uint32_t u = (b & 3) == 0 ? r.x : (b & 3) == 1 ? r.y : (b & 3) == 2 ? r.z : r.w;
            int lo = 0, hi = a.n_thr;  // first index with thr[idx] > u
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (a.thr[mid] <= u) lo = mid + 1; else hi = mid;
            }
            neg = min(lo, a.n_thr - 1);
        }
        a.neg_out[b] = neg;
    }
    __syncthreads();

    float dMcol[D];  // lane q holds column q of this warp's dM
#pragma unroll
    for (int p = 0; p < D; ++p) dMcol[p] = 0.f;
    float loss = 0.f;
    bool ok = true;  // every dZ contribution fits the fixed-point range

    for (int b = b0 + warp; b < b1; b += n_warps) {
        const int u = a.batch[2 * b], v = a.batch[2 * b + 1];
        const int ng = (int)a.neg_out[b];
        const float zu = a.Zi[(size_t)u * D + lane], zn = a.Zi[(size_t)ng * D + lane], zv = a.Zj[(size_t)v * D + lane];
        float av = 0.f;  // a[lane] = sum_q M[lane][q] z_v[q]
#pragma unroll
        for (int q = 0; q < D; ++q) av = fmaf(Ms[lane][q], __shfl_sync(kFull, zv, q), av);
        const float pos = warp_sum(zu * av), neg = warp_sum(zn * av);
        float dpos, dneg;
        if (a.loss_kind == DGN_LOSS_HINGE) {
            const float diff = neg - (pos - a.margin);
            const float act = diff > 0.f ? 1.f : 0.f;
            loss += fmaxf(diff, 0.f);
            dpos = -act;
            dneg = act;
        } else {
            // softplus(-pos) + w softplus(neg), stable form
            loss += fmaxf(-pos, 0.f) + log1pf(expf(-fabsf(pos))) + a.neg_weight * (fmaxf(neg, 0.f) + log1pf(expf(-fabsf(neg))));
            dpos = -1.f / (1.f + expf(pos));
            dneg = a.neg_weight / (1.f + expf(-neg));
        }
        if (lane == 0) {
            a.pos_out[b] = pos;
            a.neg_score_out[b] = neg;
        }
        const float w = dpos * zu + dneg * zn;
        float cv = 0.f;  // (M^T w)[lane]
#pragma unroll
        for (int p = 0; p < D; ++p) {
            const float wp = __shfl_sync(kFull, w, p);
            cv = fmaf(Ms[p][lane], wp, cv);
            dMcol[p] = fmaf(wp, zv, dMcol[p]);
        }
        // nodes repeat inside a batch: the scatter-add runs on 2^-40 fixed-point integers, whose sums do not
        // depend on the order of the atomics (float atomics would make the replicas of a multi-GPU run drift)
        if (dpos != 0.f) ok &= fixed_add(a.dZi + (size_t)u * D + lane, dpos * av);
        if (dneg != 0.f) ok &= fixed_add(a.dZi + (size_t)ng * D + lane, dneg * av);
        if (dpos != 0.f || dneg != 0.f) ok &= fixed_add(a.dZj + (size_t)v * D + lane, cv);
    }
    const int overflow = __syncthreads_or(!ok);  // every thread of the CTA gets here (no early exit above)

    if (lane == 0) loss_w[warp] = loss;
    for (int w = 0; w < n_warps; ++w) {  // ordered reduction over warps
        if (warp == w) {
#pragma unroll
            for (int p = 0; p < D; ++p) dMs[p][lane] += dMcol[p];
        }
        __syncthreads();
    }
    // this CTA's partials
    float *part = a.scratch + (size_t)blockIdx.x * (D * D + 1);
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) part[i] = dMs[i >> 5][i & 31];
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < n_warps; ++w) s += loss_w[w];
        part[D * D] = overflow ? __int_as_float(0x7fc00000) : s;  // NaN poisons the sum over CTAs
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(a.ticket, 1u);
        is_last = ticket == gridDim.x - 1;
        if (is_last) *a.ticket = 0u;  // ready for the next launch
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
        float s = 0.f;
        for (unsigned c = 0; c < gridDim.x; ++c) s += a.scratch[(size_t)c * (D * D + 1) + i];
        dMs[i >> 5][i & 31] = s;
    }
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (unsigned c = 0; c < gridDim.x; ++c) s += a.scratch[(size_t)c * (D * D + 1) + D * D];
        a.loss_out[0] = s;
        a.loss_out[1] = __uint_as_float(dyn->seq);  // travels to the host with the loss (one 8-byte copy)
    }
    __syncthreads();
    // decoder-parameter gradients
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) {  // blockDim is a multiple of 32: q == lane
        const int p = i >> 5, q = i & 31;
        switch (a.decoder) {
            case DGN_DEC_BILINEAR: a.g_loc[p * D + q] = dMs[p][q]; break;
            case DGN_DEC_DISTMULT:
                if (p == 0) a.g_loc[q] = dMs[q][q];
                break;
            case DGN_DEC_DEDICOM: {
                a.g_glb[p * D + q] = a.loc[p] * a.loc[q] * dMs[p][q];
                // dd[p] = sum_q dM[p][q] R[p][q] d[q] + dM[q][p] R[q][p] d[q]
                const float t = (dMs[p][q] * a.glb[p * D + q] + dMs[q][p] * a.glb[q * D + p]) * a.loc[q];
                const float s = warp_sum(t);
                if (q == 0) a.g_loc[p] = s;
            } break;
            default: break;
        }
    }
}

// ------------------------------------------------------------------------------ all pairs
// out[r][u][v] = sum_{p,q} Z_i[u][p] M_r[p][q] Z_j[v][q]; CTA = 64 x 64 tile of one relation,
// 256 threads, 4 x 4 outputs per thread.  T = Z_i M_r is formed on the fly in shared memory.
__global__ void __launch_bounds__(256) predict_kernel(const PredictArgs a) {
    __shared__ float Ms[D][D + 1];
    __shared__ float Zs[64][D + 1];
    __shared__ float Ts[64][D + 1];
    __shared__ float Vs[64][D + 1];
    const int r = blockIdx.z;
    const float *loc = a.loc + (size_t)r * a.loc_stride;
    const int u0 = blockIdx.y * 64, v0 = blockIdx.x * 64;
    for (int i = threadIdx.x; i < D * D; i += 256) Ms[i >> 5][i & 31] = relation_entry(a.decoder, a.glb, loc, i >> 5, i & 31);
    for (int i = threadIdx.x; i < 64 * D; i += 256) {
        const int rr = i >> 5, c = i & 31;
        Zs[rr][c] = u0 + rr < a.n_i ? a.Zi[(size_t)(u0 + rr) * D + c] : 0.f;
        Vs[rr][c] = v0 + rr < a.n_j ? a.Zj[(size_t)(v0 + rr) * D + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * D; i += 256) {
        const int rr = i >> 5, c = i & 31;
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < D; ++p) s = fmaf(Zs[rr][p], Ms[p][c], s);
        Ts[rr][c] = s;
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // columns tx + 16 j, rows ty + 16 i
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int p = 0; p < D; ++p) {
        float t[4], z[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = Ts[ty + 16 * i][p], z[i] = Vs[tx + 16 * i][p];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(t[i], z[j], acc[i][j]);
    }
    float *out = a.out + (size_t)r * a.n_i * a.n_j;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int u = u0 + ty + 16 * i;
        if (u >= a.n_i) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = v0 + tx + 16 * j;
            if (v < a.n_j) out[(size_t)u * a.n_j + v] = acc[i][j];
        }
    }
}

// one warp per edge: sigma(z_u^T M z_v)
__global__ void __launch_bounds__(256) predict_edges_kernel(const PredictArgs a, const int *__restrict__ edges,
                                                            int n_edges, int apply_sigmoid, float *__restrict__ out) {
    __shared__ float Ms[D][D + 1];
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) Ms[i >> 5][i & 31] = relation_entry(a.decoder, a.glb, a.loc, i >> 5, i & 31);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int e = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (e >= n_edges) return;
    const int u = edges[2 * e], v = edges[2 * e + 1];
    const float zu = a.Zi[(size_t)u * D + lane], zv = a.Zj[(size_t)v * D + lane];
    float av = 0.f;
#pragma unroll
    for (int q = 0; q < D; ++q) av = fmaf(Ms[lane][q], __shfl_sync(kFull, zv, q), av);
    const float s = warp_sum(zu * av);
    if (lane == 0) out[e] = apply_sigmoid ? 1.f / (1.f + expf(-s)) : s;
}

// dense (glb, loc) of model.py:116-137 for the latent_inters / latent_varies fetches
__global__ void relation_matrices_kernel(int decoder, const float *glb, const float *loc, float *glb_out, float *loc_out) {
    const int p = threadIdx.x >> 5, q = threadIdx.x & 31;
    const float eye = p == q ? 1.f : 0.f;
    float gv = eye, lv = eye;
    if (decoder == DGN_DEC_DISTMULT) gv = p == q ? loc[p] : 0.f;
    else if (decoder == DGN_DEC_BILINEAR) gv = loc[p * D + q];
    else if (decoder == DGN_DEC_DEDICOM) { gv = glb[p * D + q]; lv = p == q ? loc[p] : 0.f; }
    glb_out[p * D + q] = gv;
    loc_out[p * D + q] = lv;
}

}  // namespace

void launch_decode(const StepDyn *dyn, cudaStream_t s) {
    decode_kernel<<<kDecodeCtas, kDecodeThreads, 0, s>>>(dyn);
    CUDA_CHECK(cudaGetLastError());
}

// every node type in one launch (blockIdx.y = type); the fixed-point accumulator is cleared on the way out, so the
// next step needs no memset between its forward and backward passes
__global__ void fixed_to_float_clear_kernel(const FixedBatch fb) {
    long long *q = fb.q[blockIdx.y];
    float *out = fb.out[blockIdx.y];
    const size_t n = fb.n[blockIdx.y];
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        out[i] = (float)((double)q[i] * (1.0 / 1099511627776.0));
        q[i] = 0;
    }
}
void launch_fixed_to_float_clear(const FixedBatch &fb, cudaStream_t s) {
    size_t most = 0;
    for (int t = 0; t < fb.count; ++t) most = std::max(most, fb.n[t]);
    if (fb.count == 0 || most == 0) return;
    dim3 grid((unsigned)std::min<size_t>((most + 255) / 256, 148 * 4), (unsigned)fb.count);
    fixed_to_float_clear_kernel<<<grid, 256, 0, s>>>(fb);
    CUDA_CHECK(cudaGetLastError());
}

void launch_fixed_to_float(const long long *q, float *out, size_t n, cudaStream_t s) {
    if (n == 0) return;
    fixed_to_float_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, s>>>(q, out, n);
    CUDA_CHECK(cudaGetLastError());
}

void launch_predict(const PredictArgs &a, cudaStream_t s) {
    if (a.count == 0) return;
    dim3 grid((unsigned)((a.n_j + 63) / 64), (unsigned)((a.n_i + 63) / 64), (unsigned)a.count);
    predict_kernel<<<grid, 256, 0, s>>>(a);
    CUDA_CHECK(cudaGetLastError());
}

void launch_predict_edges(const PredictArgs &a, const int *edges, int n_edges, int apply_sigmoid, float *out,
                          cudaStream_t s) {
    if (n_edges == 0) return;
    predict_edges_kernel<<<(unsigned)((n_edges + 7) / 8), 256, 0, s>>>(a, edges, n_edges, apply_sigmoid, out);
    CUDA_CHECK(cudaGetLastError());
}

void launch_relation_matrices(int decoder, const float *glb, const float *loc, float *glb_out, float *loc_out,
                              cudaStream_t s) {
    relation_matrices_kernel<<<1, 1024, 0, s>>>(decoder, glb, loc, glb_out, loc_out);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace dgn
