// Dense per-relation contractions of layer 2 (reference: tf.nn.dropout + tf.matmul(x, W2_k),
// decagon/deep/layers.py:112-113, and their autodiff):
//   project_kernel : P2_k  = (H_j (.) m_k / q) W2_k                 [n_j, D1] x [D1, 32]
//   dw2_kernel     : dW2_k = (H_j (.) m_k / q)^T G2_k               [D1, n_j] x [n_j, 32]
//   dh_kernel      : dH_j += (G2_k W2_k^T) (.) m_k / q   summed over the relations of a chunk
// CUDA-core fp32 (FFMA) versions: exact fp32 semantics, register-tiled with the small operand
// (W2_k, 8 KB) in shared memory.  hidden2 is fixed at 32 (one warp-wide panel).
#include "dgn_internal.cuh"

namespace dgn {
namespace {

constexpr int kD2 = 32;

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

template <int D1>
__global__ void __launch_bounds__(128) project_kernel(const DenseArgs a) {
    constexpr int P1 = D1 / 32;
    __shared__ __align__(16) float Ws[D1 * kD2];
    const int k = blockIdx.y;
    const int c = blockIdx.x * 128 + threadIdx.x;
    const float *W = a.W2 + (size_t)k * D1 * kD2;
    for (int i = threadIdx.x * 4; i < D1 * kD2; i += 128 * 4) *reinterpret_cast<float4 *>(Ws + i) = ld4(W + i);
    __syncthreads();
    if (c >= a.n_j) return;

    float h[D1];
#pragma unroll
    for (int p = 0; p < P1; ++p) {
        const float *src = a.H + ((size_t)p * a.n_j + c) * 32;
        uint32_t bits = 0xffffffffu;
        float sc = 1.f;
        if (a.mask != nullptr) {
            bits = a.mask[((size_t)k * a.n_j + c) * P1 + p];
            sc = a.scale;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 x = ld4(src + 4 * q);
            h[p * 32 + 4 * q + 0] = (bits >> (4 * q + 0)) & 1u ? x.x * sc : 0.f;
            h[p * 32 + 4 * q + 1] = (bits >> (4 * q + 1)) & 1u ? x.y * sc : 0.f;
            h[p * 32 + 4 * q + 2] = (bits >> (4 * q + 2)) & 1u ? x.z * sc : 0.f;
            h[p * 32 + 4 * q + 3] = (bits >> (4 * q + 3)) & 1u ? x.w * sc : 0.f;
        }
    }
    float out[kD2];
#pragma unroll
    for (int n = 0; n < kD2; ++n) out[n] = 0.f;
#pragma unroll
    for (int m = 0; m < D1; ++m) {
#pragma unroll
        for (int n4 = 0; n4 < kD2 / 4; ++n4) {
            const float4 w = ld4(Ws + m * kD2 + 4 * n4);
            out[4 * n4 + 0] = fmaf(h[m], w.x, out[4 * n4 + 0]);
            out[4 * n4 + 1] = fmaf(h[m], w.y, out[4 * n4 + 1]);
            out[4 * n4 + 2] = fmaf(h[m], w.z, out[4 * n4 + 2]);
            out[4 * n4 + 3] = fmaf(h[m], w.w, out[4 * n4 + 3]);
        }
    }
    float *dst = a.P2 + ((size_t)k * a.n_j + c) * kD2;
#pragma unroll
    for (int n4 = 0; n4 < kD2 / 4; ++n4)
        *reinterpret_cast<float4 *>(dst + 4 * n4) =
            make_float4(out[4 * n4], out[4 * n4 + 1], out[4 * n4 + 2], out[4 * n4 + 3]);
}

// block = 2 * D1 threads; thread owns a 4 (rows of dW2) x 4 (columns) patch
template <int D1>
__global__ void __launch_bounds__(2 * D1) dw2_kernel(const DenseArgs a) {
    constexpr int P1 = D1 / 32, TR = 32, NT = 2 * D1;
    __shared__ __align__(16) float Hs[TR * D1];
    __shared__ __align__(16) float Gs[TR * kD2];
    const int k = blockIdx.y, chunk = blockIdx.x;
    const int row0 = chunk * a.rows_per_chunk;
    const int row1 = min(row0 + a.rows_per_chunk, a.n_j);
    const int n0 = (threadIdx.x & 7) * 4, m0 = (threadIdx.x >> 3) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int base = row0; base < row1; base += TR) {
        // stage TR rows of the masked, scaled H (panel layout -> row-major) and of G2
        for (int i = threadIdx.x; i < TR * D1 / 4; i += NT) {
            const int rl = i / (D1 / 4), q4 = i % (D1 / 4);  // q4-th float4 of row rl
            const int p = q4 / 8, q = q4 % 8, c = base + rl;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < row1) {
                x = ld4(a.H + ((size_t)p * a.n_j + c) * 32 + 4 * q);
                if (a.mask != nullptr) {
                    const uint32_t bits = a.mask[((size_t)k * a.n_j + c) * P1 + p] >> (4 * q);
                    x.x = bits & 1u ? x.x * a.scale : 0.f;
                    x.y = bits & 2u ? x.y * a.scale : 0.f;
                    x.z = bits & 4u ? x.z * a.scale : 0.f;
                    x.w = bits & 8u ? x.w * a.scale : 0.f;
                }
            }
            *reinterpret_cast<float4 *>(Hs + rl * D1 + 4 * q4) = x;
        }
        for (int i = threadIdx.x; i < TR * kD2 / 4; i += NT) {
            const int rl = i / 8, q = i % 8, c = base + rl;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < row1) x = ld4(a.G2 + ((size_t)k * a.n_j + c) * kD2 + 4 * q);
            *reinterpret_cast<float4 *>(Gs + rl * kD2 + 4 * q) = x;
        }
        __syncthreads();
#pragma unroll 8
        for (int rl = 0; rl < TR; ++rl) {
            const float4 hv = ld4(Hs + rl * D1 + m0);
            const float4 gv = ld4(Gs + rl * kD2 + n0);
            const float hh[4] = {hv.x, hv.y, hv.z, hv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(hh[i], gg[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *dst = a.dW2 + ((size_t)k * a.n_row_chunks + chunk) * D1 * kD2;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4 *>(dst + (m0 + i) * kD2 + n0) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
}

__global__ void dw2_reduce_kernel(const float *__restrict__ part, float *__restrict__ out, int n_chunks, int elems,
                                  long long total) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long k = i / elems, e = i % elems;
    float s = 0.f;
    for (int c = 0; c < n_chunks; ++c) s += part[(k * n_chunks + c) * elems + e];
    out[i] = s;
}

template <int D1>
__global__ void __launch_bounds__(128) dh_kernel(const DenseArgs a) {
    constexpr int P1 = D1 / 32;
    __shared__ __align__(16) float Ws[D1 * kD2];
    const int c = blockIdx.x * 128 + threadIdx.x;
    const bool live = c < a.n_j;
    const int k0 = blockIdx.y * a.rel_per_chunk, k1 = min(k0 + a.rel_per_chunk, a.K);
    float acc[D1];
#pragma unroll
    for (int m = 0; m < D1; ++m) acc[m] = 0.f;

    for (int k = k0; k < k1; ++k) {
        const float *W = a.W2 + (size_t)k * D1 * kD2;
        __syncthreads();  // previous relation's Ws is no longer read
        for (int i = threadIdx.x * 4; i < D1 * kD2; i += 128 * 4) *reinterpret_cast<float4 *>(Ws + i) = ld4(W + i);
        __syncthreads();
        if (!live) continue;
        float g[kD2];
        const float *src = a.G2 + ((size_t)k * a.n_j + c) * kD2;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 x = ld4(src + 4 * q);
            g[4 * q] = x.x, g[4 * q + 1] = x.y, g[4 * q + 2] = x.z, g[4 * q + 3] = x.w;
        }
        uint32_t bits[P1];
        float sc = 1.f;
#pragma unroll
        for (int p = 0; p < P1; ++p) bits[p] = 0xffffffffu;
        if (a.mask != nullptr) {
            sc = a.scale;
#pragma unroll
            for (int p = 0; p < P1; ++p) bits[p] = a.mask[((size_t)k * a.n_j + c) * P1 + p];
        }
#pragma unroll
        for (int m = 0; m < D1; ++m) {
            float dot = 0.f;
#pragma unroll
            for (int n4 = 0; n4 < kD2 / 4; ++n4) {
                const float4 w = ld4(Ws + m * kD2 + 4 * n4);
                dot = fmaf(g[4 * n4 + 0], w.x, dot);
                dot = fmaf(g[4 * n4 + 1], w.y, dot);
                dot = fmaf(g[4 * n4 + 2], w.z, dot);
                dot = fmaf(g[4 * n4 + 3], w.w, dot);
            }
            if ((bits[m >> 5] >> (m & 31)) & 1u) acc[m] = fmaf(dot, sc, acc[m]);
        }
    }
    if (!live) return;
#pragma unroll
    for (int p = 0; p < P1; ++p) {
        float *dst = a.dHpart + (((size_t)blockIdx.y * P1 + p) * a.n_j + c) * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4 *>(dst + 4 * q) =
                make_float4(acc[p * 32 + 4 * q], acc[p * 32 + 4 * q + 1], acc[p * 32 + 4 * q + 2], acc[p * 32 + 4 * q + 3]);
    }
}

#define DGN_DISPATCH_D1(D1, D2, CALL)                                                                              \
    do {                                                                                                           \
        if ((D2) != kD2) DGN_FAIL(DGN_ERR_UNSUPPORTED, "hidden2 = %d is not supported (must be 32)", (D2));        \
        switch (D1) {                                                                                              \
            case 32: { constexpr int kD1 = 32; CALL; } break;                                                      \
            case 64: { constexpr int kD1 = 64; CALL; } break;                                                      \
            case 128: { constexpr int kD1 = 128; CALL; } break;                                                    \
            default: DGN_FAIL(DGN_ERR_UNSUPPORTED, "hidden1 = %d is not supported (32, 64 or 128)", (D1));         \
        }                                                                                                          \
    } while (0)

}  // namespace

void launch_project(const DenseArgs &a, int D1, int D2, cudaStream_t s) {
    if (a.K == 0 || a.n_j == 0) return;
    dim3 grid((unsigned)((a.n_j + 127) / 128), (unsigned)a.K), block(128);
    DGN_DISPATCH_D1(D1, D2, (project_kernel<kD1><<<grid, block, 0, s>>>(a)));
    CUDA_CHECK(cudaGetLastError());
}

void launch_dw2(const DenseArgs &a, int D1, int D2, cudaStream_t s) {
    if (a.K == 0 || a.n_j == 0) return;
    dim3 grid((unsigned)a.n_row_chunks, (unsigned)a.K), block(2 * D1);
    DGN_DISPATCH_D1(D1, D2, (dw2_kernel<kD1><<<grid, block, 0, s>>>(a)));
    CUDA_CHECK(cudaGetLastError());
}

void launch_dw2_reduce(const float *part, float *out, int K, int n_chunks, int elems, cudaStream_t s) {
    const long long total = (long long)K * elems;
    if (total == 0) return;
    dw2_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(part, out, n_chunks, elems, total);
    CUDA_CHECK(cudaGetLastError());
}

void launch_dh(const DenseArgs &a, int D1, int D2, cudaStream_t s) {
    if (a.K == 0 || a.n_j == 0) return;
    dim3 grid((unsigned)((a.n_j + 127) / 128), (unsigned)a.n_kchunks), block(128);
    DGN_DISPATCH_D1(D1, D2, (dh_kernel<kD1><<<grid, block, 0, s>>>(a)));
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace dgn
