// Batched evaluation: the scores of the sampled validation coordinates of MANY
// relations of one group in one launch, and AUROC / AUPRC of the pooled scores on the device
// (DecagonAccuracyEvaluator.evaluateAll, DecagonAccuracyEvaluator.py:57-91, 115-186; sklearn's
// roc_auc_score / average_precision_score definitions: distinct-threshold ROC trapezoid, step-wise AP).
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>

#include "dgn_internal.cuh"

namespace dgn {
namespace {

constexpr int D = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
    return x;
}

// row `p` of M_r = loc glb loc (model.py:116-137), entry q
__device__ __forceinline__ float relation_entry(int decoder, const float *glb, const float *loc, int p, int q) {
    switch (decoder) {
        case DGN_DEC_INNERPRODUCT: return p == q ? 1.f : 0.f;
        case DGN_DEC_DISTMULT: return p == q ? loc[p] : 0.f;
        case DGN_DEC_BILINEAR: return loc[p * D + q];
        default: return loc[p] * glb[p * D + q] * loc[q];  // dedicom
    }
}

// one warp per edge (k, u, v): sigma(z_u^T M_k z_v); same operation order as predict_edges_kernel
__global__ void __launch_bounds__(256) predict_edges_multi_kernel(const PredictArgs a, const int *__restrict__ rel_k,
                                                                  const int *__restrict__ edges, long long n_edges,
                                                                  int apply_sigmoid, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long e = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (e >= n_edges) return;
    const float *loc = a.loc ? a.loc + (long long)rel_k[e] * a.loc_stride : nullptr;
    const int u = edges[2 * e], v = edges[2 * e + 1];
    const float zu = a.Zi[(size_t)u * D + lane], zv = a.Zj[(size_t)v * D + lane];
    float av = 0.f;
#pragma unroll
    for (int q = 0; q < D; ++q) av = fmaf(relation_entry(a.decoder, a.glb, loc, lane, q), __shfl_sync(kFull, zv, q), av);
    const float s = warp_sum(zu * av);
    if (lane == 0) out[e] = apply_sigmoid ? 1.f / (1.f + expf(-s)) : s;
}

// Scores sorted in descending order with their labels.  The ROC / PR curves have one point per DISTINCT score
// (the end of each run of equal scores): (FP, TP) after the run.  AUROC = sum of trapezoids between consecutive
// points / (P N), kept in integers (2 x area); AP = sum over points of (TP - TP_prev) / P * TP / (TP + FP).
// One CTA: thread t owns the contiguous chunk t, a serial pass over the chunks' summaries carries the counts
// and the previous point, the partial sums are reduced in a fixed order (deterministic).
constexpr int kAucThreads = 512;
__global__ void __launch_bounds__(kAucThreads) auc_kernel(const float *__restrict__ key, const unsigned char *__restrict__ lab,
                                                          long long n, double *__restrict__ out /* auroc, auprc, P, N */) {
    __shared__ long long tp_chunk[kAucThreads];       // positives inside the chunk
    __shared__ long long last_end[kAucThreads];       // index + 1 of the chunk's last run end, 0 = none
    __shared__ long long last_tp[kAucThreads];        // positives inside the chunk up to that run end
    __shared__ long long tp_before[kAucThreads];      // positives before the chunk
    __shared__ long long prev_idx[kAucThreads], prev_tp[kAucThreads];  // the curve point preceding the chunk
    __shared__ unsigned long long area2[kAucThreads];
    __shared__ double ap[kAucThreads];
    __shared__ long long total_pos;
    const int t = threadIdx.x;
    const long long per = (n + kAucThreads - 1) / kAucThreads;
    const long long lo = min(n, per * t), hi = min(n, lo + per);
    {
        long long tp = 0, le = 0, lt = 0;
        for (long long i = lo; i < hi; ++i) {
            tp += lab[i] ? 1 : 0;
            if (i + 1 == n || key[i + 1] != key[i]) le = i + 1, lt = tp;
        }
        tp_chunk[t] = tp, last_end[t] = le, last_tp[t] = lt;
    }
    __syncthreads();
    if (t == 0) {
        long long tp = 0, pi = 0, pt = 0;
        for (int c = 0; c < kAucThreads; ++c) {
            tp_before[c] = tp, prev_idx[c] = pi, prev_tp[c] = pt;
            if (last_end[c]) pi = last_end[c], pt = tp + last_tp[c];
            tp += tp_chunk[c];
        }
        total_pos = tp;
    }
    __syncthreads();
    {
        const double P = (double)total_pos;
        long long tp = tp_before[t], pi = prev_idx[t], pt = prev_tp[t];
        unsigned long long a2 = 0;
        double s = 0.0;
        for (long long i = lo; i < hi; ++i) {
            tp += lab[i] ? 1 : 0;
            if (i + 1 == n || key[i + 1] != key[i]) {
                const long long fp = (i + 1) - tp, pfp = pi - pt;
                a2 += (unsigned long long)(fp - pfp) * (unsigned long long)(tp + pt);
                if (tp > pt) s += ((double)(tp - pt) / P) * ((double)tp / (double)(i + 1));
                pi = i + 1, pt = tp;
            }
        }
        area2[t] = a2, ap[t] = s;
    }
    __syncthreads();
    for (int w = kAucThreads / 2; w > 0; w >>= 1) {
        if (t < w) area2[t] += area2[t + w], ap[t] += ap[t + w];
        __syncthreads();
    }
    if (t == 0) {
        const double P = (double)total_pos, N = (double)(n - total_pos);
        out[0] = (total_pos > 0 && total_pos < n) ? (double)area2[0] / (2.0 * P * N) : nan("");
        out[1] = total_pos > 0 ? ap[0] : nan("");
        out[2] = P, out[3] = N;
    }
}

__global__ void iota_kernel(int *__restrict__ idx, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) idx[i] = (int)i;
}

}  // namespace

// candidate ranking (GreedyActiveLearner._getRankedPossibilities, main/ActiveLearner/GreedyActiveLearner.py:84-92):
// argsort of the scores in descending order, stable (ties keep their input order)
size_t rank_sort_bytes(long long n) {
    size_t bytes = 0;
    CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const float *)nullptr, (float *)nullptr, (const int *)nullptr,
                                                         (int *)nullptr, n));
    return bytes;
}
void launch_rank(const float *scores, long long n, int *idx_in, float *sorted_scores, int *order, void *tmp, size_t tmp_bytes,
                 cudaStream_t s) {
    iota_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, s>>>(idx_in, n);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, scores, sorted_scores, idx_in, order, n, 0, 32, s));
}

void launch_predict_edges_multi(const PredictArgs &a, const int *rel_k, const int *edges, long long n_edges,
                                int apply_sigmoid, float *out, cudaStream_t s) {
    if (n_edges == 0) return;
    predict_edges_multi_kernel<<<(unsigned)((n_edges + 7) / 8), 256, 0, s>>>(a, rel_k, edges, n_edges, apply_sigmoid, out);
    CUDA_CHECK(cudaGetLastError());
}

size_t auc_sort_bytes(long long n) {
    size_t bytes = 0;
    CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const float *)nullptr, (float *)nullptr,
                                                         (const unsigned char *)nullptr, (unsigned char *)nullptr, n));
    return bytes;
}

// scores / labels -> sorted copies -> out[4] = {auroc, auprc, positives, negatives} (device doubles)
void launch_auc(const float *scores, const unsigned char *labels, long long n, float *sorted_scores,
                unsigned char *sorted_labels, void *tmp, size_t tmp_bytes, double *out, cudaStream_t s) {
    CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, scores, sorted_scores, labels, sorted_labels, n,
                                                         0, 32, s));
    auc_kernel<<<1, kAucThreads, 0, s>>>(sorted_scores, sorted_labels, n, out);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace dgn
