// Dense per-relation contractions of layer 2 (reference: tf.nn.dropout + tf.matmul(x, W2_k),
// decagon/deep/layers.py:112-113, and their autodiff):
//   project_kernel : P2_k  = (H_j (.) m_k / q) W2_k                 [n_j, D1] x [D1, 32]
//   dw2_kernel     : dW2_k = (H_j (.) m_k / q)^T G2_k               [D1, n_j] x [n_j, 32]
//   dh_kernel      : dH_j += (G2_k W2_k^T) (.) m_k / q   summed over the relations of a slot
//
// Exact-fp32 CUDA-core (FFMA) kernels, register-tiled 4 x 8 per thread.  The work is
// K x n_j rows (1928 x 645 at the polypharmacy shape) against operands that are tiny per relation,
// so every kernel is persistent over relations: a CTA owns one block of kRowBlock rows of H_j,
// keeps it in shared memory for its whole life and streams W2_k / G2_k / the dropout words of
// its relations through a cp.async double buffer (one CTA barrier per relation).  The dropout
// keep-bit of (row, feature) is folded into the FFMA as a predicate (one LOP3 per 8 FFMA); the
// 1/keep scale is applied once to the result.  hidden2 is fixed at 32.
// (Measured and rejected: 8 x 8 tiles on packed fma.rn.f32x2 with 6 warps per SM -- 209 us against 172 us
// for project at the polypharmacy shape; predicated f32x2 FMAs compile to FFMA2 + 2 SEL.)
#include <stdlib.h>

#include "dgn_internal.cuh"

namespace dgn {
namespace {

constexpr int kD2 = 32;

template <int D1>
struct DenseCfg {
    // rows of H_j per CTA = threads of project/dh (warp w owns rows [32 w, 32 w + 32))
    static constexpr int RB = D1 == 128 ? 128 : 352;
    // dw2: D1 threads tile the [D1, 32] result, NT2 / D1 row ranges are reduced through smem
    static constexpr int NT2 = D1 == 128 ? 256 : 384;
};

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// acc[0..7] += h * (w0, w1) if (word & BIT): one LOP3 setting a predicate + 8 predicated FFMA
template <uint32_t BIT>
__device__ __forceinline__ void fma8_if(float (&acc)[8], float h, const float4 &w0, const float4 &w1, uint32_t word) {
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %9, %10;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "@p fma.rn.f32 %0, %8, %11, %0;\n\t"
        "@p fma.rn.f32 %1, %8, %12, %1;\n\t"
        "@p fma.rn.f32 %2, %8, %13, %2;\n\t"
        "@p fma.rn.f32 %3, %8, %14, %3;\n\t"
        "@p fma.rn.f32 %4, %8, %15, %4;\n\t"
        "@p fma.rn.f32 %5, %8, %16, %5;\n\t"
        "@p fma.rn.f32 %6, %8, %17, %6;\n\t"
        "@p fma.rn.f32 %7, %8, %18, %7;\n\t}"
        : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]), "+f"(acc[6]), "+f"(acc[7])
        : "f"(h), "r"(word), "n"(BIT), "f"(w0.x), "f"(w0.y), "f"(w0.z), "f"(w0.w), "f"(w1.x), "f"(w1.y), "f"(w1.z),
          "f"(w1.w));
}
// acc += t if (word & BIT)
template <uint32_t BIT>
__device__ __forceinline__ void add_if(float &acc, float t, uint32_t word) {
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %2, %3;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "@p add.f32 %0, %0, %1;\n\t}"
        : "+f"(acc)
        : "f"(t), "r"(word), "n"(BIT));
}

__device__ __forceinline__ float comp(const float4 &v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// relations [k_begin, k_end) of slot `slot` out of n_slots (equal work per relation)
__device__ __forceinline__ void slot_range(int slot, int n_slots, int K, int &k_begin, int &k_end) {
    k_begin = (int)((long long)slot * K / n_slots);
    k_end = (int)((long long)(slot + 1) * K / n_slots);
}

// ------------------------------------------------------------------------------ P2 = Hm W2
// thread (warp w, rg = lane / 4, tn = lane % 4): rows 32 w + rg + 8 i (i < 4), columns 4 tn + {0..3} and
// 16 + 4 tn + {0..3}.  Hs rows are padded to D1 + 4 floats so that the 8 row groups of a warp hit 8
// different 16-byte bank groups.
template <int D1>
__global__ void __launch_bounds__(DenseCfg<D1>::RB, 1) project_kernel(const DenseArgs a) {
    constexpr int P1 = D1 / 32, RB = DenseCfg<D1>::RB, NT = RB, HS = D1 + 4;
    extern __shared__ __align__(16) float smem[];
    float *Hs = smem;            // [RB][HS]
    float *Ws = smem + RB * HS;  // [2][D1][32]
    const int rb = blockIdx.x % a.n_rb, slot = blockIdx.x / a.n_rb;
    const int row0 = rb * RB;
    int k_begin, k_end;
    slot_range(slot, a.n_slots, a.K, k_begin, k_end);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, rg = lane >> 2, tn = lane & 3;

    for (int i = threadIdx.x; i < RB * (D1 / 4); i += NT) {
        const int rl = i / (D1 / 4), q4 = i % (D1 / 4), p = q4 >> 3, q = q4 & 7, c = row0 + rl;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < a.n_j) x = ld4(a.H + ((size_t)p * a.n_j + c) * 32 + 4 * q);
        *reinterpret_cast<float4 *>(Hs + rl * HS + 4 * q4) = x;
    }
    auto issue_w = [&](int k, int buf) {
        const float *src = a.W2 + (size_t)k * D1 * kD2;
        for (int i = threadIdx.x; i < D1 * kD2 / 4; i += NT) cp_async16(Ws + buf * D1 * kD2 + 4 * i, src + 4 * i);
        cp_async_commit();
    };
    int crow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) crow[i] = row0 + warp * 32 + rg + 8 * i;
    auto load_mask = [&](int k, uint32_t (&mk)[P1][4]) {
#pragma unroll
        for (int p = 0; p < P1; ++p)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                mk[p][i] = 0xffffffffu;
                if (a.mask != nullptr && crow[i] < a.n_j) mk[p][i] = __ldg(a.mask + ((size_t)k * a.n_j + crow[i]) * P1 + p);
            }
    };
    uint32_t mk[P1][4], nmk[P1][4];
    if (k_begin < k_end) {
        issue_w(k_begin, 0);
        load_mask(k_begin, mk);
    }
    const float *hrow = Hs + (warp * 32 + rg) * HS;
    for (int k = k_begin; k < k_end; ++k) {
        const int buf = (k - k_begin) & 1;
        cp_async_wait_all();
        __syncthreads();  // W2_k landed; nobody still reads the other buffer (and, first time, Hs is complete)
        if (k + 1 < k_end) {
            issue_w(k + 1, buf ^ 1);
            load_mask(k + 1, nmk);
        }
        const float *Wb = Ws + buf * D1 * kD2 + tn * 4;
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll
        for (int p = 0; p < P1; ++p) {
#pragma unroll
            for (int ms = 0; ms < 8; ++ms) {
                const int m0 = p * 32 + ms * 4;
                float4 h[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) h[i] = ld4(hrow + 8 * i * HS + m0);
#pragma unroll
                for (int mm = 0; mm < 4; ++mm) {
                    const float4 w0 = ld4(Wb + (m0 + mm) * kD2), w1 = ld4(Wb + (m0 + mm) * kD2 + 16);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // bit (ms * 4 + mm) of the row's keep word of panel p
                        switch (ms * 4 + mm) {
#define DGN_CASE(B) case B: fma8_if<(1u << B)>(acc[i], comp(h[i], mm), w0, w1, mk[p][i]); break;
                            DGN_CASE(0) DGN_CASE(1) DGN_CASE(2) DGN_CASE(3) DGN_CASE(4) DGN_CASE(5) DGN_CASE(6) DGN_CASE(7)
                            DGN_CASE(8) DGN_CASE(9) DGN_CASE(10) DGN_CASE(11) DGN_CASE(12) DGN_CASE(13) DGN_CASE(14) DGN_CASE(15)
                            DGN_CASE(16) DGN_CASE(17) DGN_CASE(18) DGN_CASE(19) DGN_CASE(20) DGN_CASE(21) DGN_CASE(22) DGN_CASE(23)
                            DGN_CASE(24) DGN_CASE(25) DGN_CASE(26) DGN_CASE(27) DGN_CASE(28) DGN_CASE(29) DGN_CASE(30) DGN_CASE(31)
#undef DGN_CASE
                        }
                    }
                }
            }
        }
        const float sc = a.mask != nullptr ? a.scale : 1.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (crow[i] < a.n_j) {
                float *dst = a.P2 + ((size_t)k * a.n_j + crow[i]) * kD2 + tn * 4;
                *reinterpret_cast<float4 *>(dst) = make_float4(acc[i][0] * sc, acc[i][1] * sc, acc[i][2] * sc, acc[i][3] * sc);
                *reinterpret_cast<float4 *>(dst + 16) = make_float4(acc[i][4] * sc, acc[i][5] * sc, acc[i][6] * sc, acc[i][7] * sc);
            }
        }
        if (k + 1 < k_end) {
#pragma unroll
            for (int p = 0; p < P1; ++p)
#pragma unroll
                for (int i = 0; i < 4; ++i) mk[p][i] = nmk[p][i];
        }
    }
}

// ------------------------------------------------------------------------------ dH += (G2 W2^T) (.) m
// CTA = (panel p of the D1 features, row block, slot).  thread: rows 32 w + rg + 8 i, features
// m = 32 p + tn + 4 j (j < 8).  G2_k rows and the 32 rows of W2_k^T that belong to panel p are staged
// with a row stride of 36 floats (conflict-free LDS.128 along the contraction index).
template <int D1>
__global__ void __launch_bounds__(DenseCfg<D1>::RB, 1) dh_kernel(const DenseArgs a) {
    constexpr int P1 = D1 / 32, RB = DenseCfg<D1>::RB, NT = RB, GS = 36;
    extern __shared__ __align__(16) float smem[];
    float *Gs = smem;                // [2][RB][GS]
    float *Wt = smem + 2 * RB * GS;  // [2][32][GS]
    const int p = blockIdx.x % P1, rb = (blockIdx.x / P1) % a.n_rb, slot = blockIdx.x / (P1 * a.n_rb);
    const int row0 = rb * RB;
    int k_begin, k_end;
    slot_range(slot, a.n_slots, a.K, k_begin, k_end);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, rg = lane >> 2, tn = lane & 3;

    for (int i = threadIdx.x; i < 2 * RB * GS + 2 * 32 * GS; i += NT) smem[i] = 0.f;
    __syncthreads();
    auto issue = [&](int k, int buf) {
        for (int i = threadIdx.x; i < RB * 8; i += NT) {
            const int rl = i >> 3, ch = i & 7, c = row0 + rl;
            if (c < a.n_j) cp_async16(Gs + (buf * RB + rl) * GS + ch * 4, a.G2 + ((size_t)k * a.n_j + c) * kD2 + ch * 4);
        }
        for (int i = threadIdx.x; i < 32 * 8; i += NT) {
            const int m = i >> 3, ch = i & 7;
            cp_async16(Wt + (buf * 32 + m) * GS + ch * 4, a.W2 + ((size_t)k * D1 + p * 32 + m) * kD2 + ch * 4);
        }
        cp_async_commit();
    };
    int crow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) crow[i] = row0 + warp * 32 + rg + 8 * i;
    auto load_mask = [&](int k, uint32_t (&mk)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            mk[i] = 0xffffffffu;
            if (a.mask != nullptr && crow[i] < a.n_j) mk[i] = __ldg(a.mask + ((size_t)k * a.n_j + crow[i]) * P1 + p);
            mk[i] >>= tn;  // bit 4 j now is feature tn + 4 j
        }
    };
    uint32_t mk[4], nmk[4];
    if (k_begin < k_end) {
        issue(k_begin, 0);
        load_mask(k_begin, mk);
    }
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k = k_begin; k < k_end; ++k) {
        const int buf = (k - k_begin) & 1;
        cp_async_wait_all();
        __syncthreads();
        if (k + 1 < k_end) {
            issue(k + 1, buf ^ 1);
            load_mask(k + 1, nmk);
        }
        const float *gb = Gs + (buf * RB + warp * 32 + rg) * GS;
        const float *wb = Wt + (buf * 32 + tn) * GS;
        float t[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) t[i][j] = 0.f;
#pragma unroll
        for (int n0 = 0; n0 < kD2; n0 += 4) {
            float4 gv[4], wv[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) gv[i] = ld4(gb + 8 * i * GS + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) wv[j] = ld4(wb + 4 * j * GS + n0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    t[i][j] = fmaf(gv[i].x, wv[j].x, t[i][j]);
                    t[i][j] = fmaf(gv[i].y, wv[j].y, t[i][j]);
                    t[i][j] = fmaf(gv[i].z, wv[j].z, t[i][j]);
                    t[i][j] = fmaf(gv[i].w, wv[j].w, t[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            add_if<(1u << 0)>(acc[i][0], t[i][0], mk[i]);
            add_if<(1u << 4)>(acc[i][1], t[i][1], mk[i]);
            add_if<(1u << 8)>(acc[i][2], t[i][2], mk[i]);
            add_if<(1u << 12)>(acc[i][3], t[i][3], mk[i]);
            add_if<(1u << 16)>(acc[i][4], t[i][4], mk[i]);
            add_if<(1u << 20)>(acc[i][5], t[i][5], mk[i]);
            add_if<(1u << 24)>(acc[i][6], t[i][6], mk[i]);
            add_if<(1u << 28)>(acc[i][7], t[i][7], mk[i]);
        }
        if (k + 1 < k_end) {
#pragma unroll
            for (int i = 0; i < 4; ++i) mk[i] = nmk[i];
        }
    }
    const float sc = a.mask != nullptr ? a.scale : 1.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (crow[i] >= a.n_j) continue;
        float *dst = a.dHpart + (((size_t)slot * P1 + p) * a.n_j + crow[i]) * 32 + tn;
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[4 * j] = acc[i][j] * sc;
    }
}

// ------------------------------------------------------------------------------ dW2 = Hm^T G2
// CTA = (row block, slot).  D1 threads tile the [D1, 32] result (thread: rows m = 4 mg + {0..3}, columns
// 4 tn + {0..3} and 16 + 4 tn + {0..3}); NT2 / D1 such groups each contract a range of the block's rows,
// the ranges are summed in order through shared memory.
template <int D1>
__global__ void __launch_bounds__(DenseCfg<D1>::NT2, 1) dw2_kernel(const DenseArgs a) {
    constexpr int P1 = D1 / 32, RB = DenseCfg<D1>::RB, NT = DenseCfg<D1>::NT2, NR = NT / D1, RR = (RB + NR - 1) / NR;
    static_assert((NR - 1) * D1 <= RB, "range partials must fit the G2 buffer");
    extern __shared__ __align__(16) float smem[];
    float *Hs = smem;                                                      // [RB][D1]
    float *Gs = smem + RB * D1;                                            // [2][RB][32]
    uint32_t *Ms = reinterpret_cast<uint32_t *>(smem + RB * D1 + 2 * RB * kD2);  // [2][RB][P1]
    const int rb = blockIdx.x % a.n_rb, slot = blockIdx.x / a.n_rb;
    const int row0 = rb * RB;
    int k_begin, k_end;
    slot_range(slot, a.n_slots, a.K, k_begin, k_end);
    const int range = threadIdx.x / D1, tr = threadIdx.x % D1, mg = tr >> 2, tn = tr & 3;
    const int m0 = 4 * mg, p = m0 >> 5, sh = m0 & 31;
    const int r_begin = range * RR, r_end = min(r_begin + RR, RB);

    for (int i = threadIdx.x; i < RB * (D1 / 4); i += NT) {
        const int rl = i / (D1 / 4), q4 = i % (D1 / 4), pp = q4 >> 3, q = q4 & 7, c = row0 + rl;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < a.n_j) x = ld4(a.H + ((size_t)pp * a.n_j + c) * 32 + 4 * q);
        *reinterpret_cast<float4 *>(Hs + rl * D1 + 4 * q4) = x;
    }
    for (int i = threadIdx.x; i < 2 * RB * kD2; i += NT) Gs[i] = 0.f;
    for (int i = threadIdx.x; i < 2 * RB * P1; i += NT) Ms[i] = 0xffffffffu;
    __syncthreads();
    auto issue = [&](int k, int buf) {
        for (int i = threadIdx.x; i < RB * 8; i += NT) {
            const int rl = i >> 3, ch = i & 7, c = row0 + rl;
            if (c < a.n_j) cp_async16(Gs + (buf * RB + rl) * kD2 + ch * 4, a.G2 + ((size_t)k * a.n_j + c) * kD2 + ch * 4);
        }
        if (a.mask != nullptr)
            for (int i = threadIdx.x; i < RB * P1; i += NT) {
                const int rl = i / P1, c = row0 + rl;
                if (c < a.n_j) cp_async4(Ms + buf * RB * P1 + i, a.mask + ((size_t)k * a.n_j + c) * P1 + (i % P1));
            }
        cp_async_commit();
    };
    if (k_begin < k_end) issue(k_begin, 0);

    for (int k = k_begin; k < k_end; ++k) {
        const int buf = (k - k_begin) & 1;
        cp_async_wait_all();
        __syncthreads();  // (A) G2_k / keep words landed; the other buffer (last used as scratch) is free
        if (k + 1 < k_end) issue(k + 1, buf ^ 1);
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        const float *hb = Hs + m0;
        const float *gb = Gs + buf * RB * kD2 + tn * 4;
        const uint32_t *mb = Ms + buf * RB * P1 + p;
#pragma unroll 4
        for (int c = r_begin; c < r_end; ++c) {
            const float4 h = ld4(hb + c * D1);
            const float4 g0 = ld4(gb + c * kD2), g1 = ld4(gb + c * kD2 + 16);
            const uint32_t w = mb[c * P1] >> sh;
            fma8_if<1u>(acc[0], h.x, g0, g1, w);
            fma8_if<2u>(acc[1], h.y, g0, g1, w);
            fma8_if<4u>(acc[2], h.z, g0, g1, w);
            fma8_if<8u>(acc[3], h.w, g0, g1, w);
        }
        __syncthreads();  // (B) everyone finished reading Gs[buf]: reuse it for the range partials
        float *scratch = Gs + buf * RB * kD2;  // [(NR - 1)][D1][32]
        if (range > 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float *dst = scratch + ((size_t)(range - 1) * D1 + m0 + i) * kD2 + tn * 4;
                *reinterpret_cast<float4 *>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                *reinterpret_cast<float4 *>(dst + 16) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
            }
        }
        __syncthreads();  // (C)
        if (range == 0) {
            const float sc = a.mask != nullptr ? a.scale : 1.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                for (int r = 1; r < NR; ++r) {
                    const float *src = scratch + ((size_t)(r - 1) * D1 + m0 + i) * kD2 + tn * 4;
                    const float4 x0 = ld4(src), x1 = ld4(src + 16);
                    acc[i][0] += x0.x, acc[i][1] += x0.y, acc[i][2] += x0.z, acc[i][3] += x0.w;
                    acc[i][4] += x1.x, acc[i][5] += x1.y, acc[i][6] += x1.z, acc[i][7] += x1.w;
                }
                float *dst = a.dW2 + ((size_t)k * a.n_rb + rb) * D1 * kD2 + (m0 + i) * kD2 + tn * 4;
                *reinterpret_cast<float4 *>(dst) = make_float4(acc[i][0] * sc, acc[i][1] * sc, acc[i][2] * sc, acc[i][3] * sc);
                *reinterpret_cast<float4 *>(dst + 16) = make_float4(acc[i][4] * sc, acc[i][5] * sc, acc[i][6] * sc, acc[i][7] * sc);
            }
            // the zero rows beyond n_j / the stale partials in Gs[buf] are overwritten by the next
            // cp.async into this buffer only for rows < n_j: clear the tail rows again
        }
        if (row0 + RB > a.n_j) {
            __syncthreads();  // (D) partial block: the scratch overwrote rows that must read as zero
            const int first = max(a.n_j - row0, 0);
            for (int i = first * kD2 + threadIdx.x; i < RB * kD2; i += NT) scratch[i] = 0.f;
        }
    }
}

// out[k][e] = sum over the n_chunks partials, in a fixed order: CTA = (k, block of 32 elements), warp w sums a
// contiguous range of chunks (four loads in flight), the 8 range sums are added in warp order.
__global__ void __launch_bounds__(256) dw2_reduce_kernel(const float *__restrict__ part, float *__restrict__ out, int n_chunks, int elems) {
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int blocks_per_k = elems / 32;
    const long long k = blockIdx.x / blocks_per_k;
    const int e = (blockIdx.x % blocks_per_k) * 32 + lane;
    const int per = (n_chunks + 7) / 8, c0 = warp * per, c1 = min(c0 + per, n_chunks);
    const float *src = part + (k * n_chunks) * elems + e;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int c = c0;
    for (; c + 4 <= c1; c += 4) {
        s0 += src[(size_t)c * elems], s1 += src[(size_t)(c + 1) * elems];
        s2 += src[(size_t)(c + 2) * elems], s3 += src[(size_t)(c + 3) * elems];
    }
    for (; c < c1; ++c) s0 += src[(size_t)c * elems];
    red[warp][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (warp == 0) {
        float s = red[0][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) s += red[w][lane];
        out[k * elems + e] = s;
    }
}

template <int D1>
size_t project_smem() { return (size_t)(DenseCfg<D1>::RB * (D1 + 4) + 2 * D1 * kD2) * sizeof(float); }
template <int D1>
size_t dh_smem() { return (size_t)(2 * DenseCfg<D1>::RB * 36 + 2 * 32 * 36) * sizeof(float); }
template <int D1>
size_t dw2_smem() { return (size_t)(DenseCfg<D1>::RB * D1 + 2 * DenseCfg<D1>::RB * kD2 + 2 * DenseCfg<D1>::RB * (D1 / 32)) * sizeof(float); }

template <typename Kernel>
void set_smem(Kernel kernel, size_t bytes) {
    CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

#define DGN_DISPATCH_D1(D1, D2, ...)                                                                             \
    do {                                                                                                           \
        if ((D2) != kD2) DGN_FAIL(DGN_ERR_UNSUPPORTED, "hidden2 = %d is not supported (must be 32)", (D2));        \
        switch (D1) {                                                                                              \
            case 32: { constexpr int kD1 = 32; __VA_ARGS__; } break;                                                      \
            case 64: { constexpr int kD1 = 64; __VA_ARGS__; } break;                                                      \
            case 128: { constexpr int kD1 = 128; __VA_ARGS__; } break;                                                    \
            default: DGN_FAIL(DGN_ERR_UNSUPPORTED, "hidden1 = %d is not supported (32, 64 or 128)", (D1));         \
        }                                                                                                          \
    } while (0)

}  // namespace



// rows per CTA of project / dh (which == 0) and of dw2 (which == 1)
int dense_row_block(int D1, int which) {
    (void)which;
    return D1 == 128 ? DenseCfg<128>::RB : DenseCfg<64>::RB;
}

void launch_project(const DenseArgs &a, int D1, int D2, cudaStream_t s) {
    if (a.K == 0 || a.n_j == 0) return;
    DGN_DISPATCH_D1(D1, D2, {
        set_smem(project_kernel<kD1>, project_smem<kD1>());
        project_kernel<kD1><<<a.n_rb * a.n_slots, DenseCfg<kD1>::RB, project_smem<kD1>(), s>>>(a);
    });
    CUDA_CHECK(cudaGetLastError());
}

void launch_dw2(const DenseArgs &a, int D1, int D2, cudaStream_t s) {
    if (a.K == 0 || a.n_j == 0) return;
    DGN_DISPATCH_D1(D1, D2, {
        set_smem(dw2_kernel<kD1>, dw2_smem<kD1>());
        dw2_kernel<kD1><<<a.n_rb * a.n_slots, DenseCfg<kD1>::NT2, dw2_smem<kD1>(), s>>>(a);
    });
    CUDA_CHECK(cudaGetLastError());
}

void launch_dw2_reduce(const float *part, float *out, int K, int n_chunks, int elems, cudaStream_t s) {
    if ((long long)K * elems == 0) return;
    DGN_REQUIRE(elems % 32 == 0, "dw2 reduce: %d elements per relation", elems);
    dw2_reduce_kernel<<<(unsigned)(K * (elems / 32)), 256, 0, s>>>(part, out, n_chunks, elems);
    CUDA_CHECK(cudaGetLastError());
}

void launch_dh(const DenseArgs &a, int D1, int D2, cudaStream_t s) {
    if (a.K == 0 || a.n_j == 0) return;
    DGN_DISPATCH_D1(D1, D2, {
        set_smem(dh_kernel<kD1>, dh_smem<kD1>());
        dh_kernel<kD1><<<(D1 / 32) * a.n_rb * a.n_slots, DenseCfg<kD1>::RB, dh_smem<kD1>(), s>>>(a);
    });
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace dgn
