// First-layer feature products for node types whose features are NOT the identity (reference:
// dropout_sparse + tf.sparse_tensor_dense_matmul(x, W1_k), decagon/deep/layers.py:23-31,89-90; public data: the
// multi-hot drug features of DecagonPublicDataNodeFeaturesBuilder.py:34-51):
//   forward  : P1_k  = (X_j (.) m_k / q) W1_k        [n_j, F_j] sparse x [F_j, d1]  -> the operand of the SpMM A_k P1_k
//   backward : dW1_k = (X_j (.) m_k / q)^T G1_k      G1_k = A_k^T dS1 [n_j, d1]     -> [F_j, d1]
// One kernel serves both: a CSR view (X by rows, or X^T by feature) whose entries carry the position of the
// non-zero in X's canonical (row, col) order, which is the index of its dropout bit (a fresh mask per relation).
// One warp per (relation, view row, panel): lane = feature of the 32-wide panel, rows of the dense operand are
// gathered through L1 / L2.  HBM-bound: the dense operand is read about once per relation.
#include "dgn_internal.cuh"

namespace dgn {
namespace {

__global__ void __launch_bounds__(256) feature_product_kernel(const FeatArgs a) {
    const int lane = threadIdx.x & 31;
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long per_panel = (long long)a.K * a.n_rows;
    if (w >= per_panel * a.P) return;
    const int p = (int)(w / per_panel);
    const long long kr = w - (long long)p * per_panel;
    const int k = (int)(kr / a.n_rows), row = (int)(kr - (long long)k * a.n_rows);
    const float *in = a.in + ((size_t)p * a.K * a.in_rows + (size_t)k * a.in_rows) * 32 + lane;
    const long long bit0 = (long long)k * a.nnz;
    float acc = 0.f;
    for (int e = a.rowptr[row]; e < a.rowptr[row + 1]; ++e) {
        float x = a.val[e];
        if (a.mask != nullptr) {
            const long long bit = bit0 + (a.eid != nullptr ? a.eid[e] : e);
            x = (a.mask[bit >> 5] >> (bit & 31)) & 1u ? x * a.scale : 0.f;
        }
        acc = fmaf(x, in[(size_t)a.col[e] * 32], acc);
    }
    a.out[((size_t)p * a.K * a.n_rows + (size_t)k * a.n_rows + row) * 32 + lane] = acc;
}

}  // namespace

void launch_feature_product(const FeatArgs &a, cudaStream_t s) {
    const long long warps = (long long)a.K * a.n_rows * a.P;
    if (warps == 0) return;
    feature_product_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(a);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace dgn
