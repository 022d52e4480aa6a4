// Probe of the tcgen05 shared-memory operand layouts on sm_100a: which shared-memory word does the tensor core read
// as element (k, n) of an MN-major B operand / (m, k) of an MN-major A operand (kind::tf32, SWIZZLE_128B)?
// The other operand is a K-major identity slice, the probed region holds its own word index (split in two passes of
// 10 bits, exact in tf32), so D reveals the address map.  nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../decagon_b200/csrc/tc_common.cuh"
using namespace dgn::tc;

// mode 0: probe B (MN-major), A = K-major identity rows; mode 1: probe A (MN-major), B = K-major identity rows
__device__ __forceinline__ uint64_t desc_lt(uint32_t saddr, uint32_t lbo, uint32_t sbo, int lt) {
    return (umma_desc(saddr, lbo, sbo) & ~(7ull << 61)) | ((uint64_t)lt << 61);
}
__global__ void __launch_bounds__(128, 1) probe(int mode, uint32_t lbo, uint32_t sbo, int pass, float *out /* [128][64] */, int lt) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *Reg = smem;           // 64 KB probed region
    unsigned char *Id = smem + 65536;    // 16 KB K-major identity tile [128 rows][128 B]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc<64>(&slot);
    if (tid == 0) mbar_init(&bar, 1);
    for (int w = tid; w < 16384; w += 128) reinterpret_cast<float *>(Reg)[w] = pass == 0 ? (float)(w & 1023) : (float)(w >> 10);
    for (int w = tid; w < 4096; w += 128) reinterpret_cast<float *>(Id)[w] = 0.f;
    __syncthreads();
    if (tid < 8) *reinterpret_cast<float *>(Id + sw128(tid, tid >> 2) + (tid & 3) * 4) = 1.f;  // Id[r][k] = (r == k)
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        if (mode == 0) {
            const uint32_t idesc = idesc_tf32(128, 64, 0, 1);
            mma_tf32(tmem, umma_desc(smem_u32(Id)), desc_lt(smem_u32(Reg), lbo, sbo, lt), idesc, 0);
        } else if (mode == 1) {
            const uint32_t idesc = idesc_tf32(128, 64, 1, 0);
            mma_tf32(tmem, desc_lt(smem_u32(Reg), lbo, sbo, lt), umma_desc(smem_u32(Id)), idesc, 0);
        } else if (mode == 2) {  // control: B K-major
            const uint32_t idesc = idesc_tf32(128, 64, 0, 0);
            mma_tf32(tmem, umma_desc(smem_u32(Id)), umma_desc(smem_u32(Reg), lbo, sbo), idesc, 0);
        } else {  // control: A K-major
            const uint32_t idesc = idesc_tf32(128, 64, 0, 0);
            mma_tf32(tmem, umma_desc(smem_u32(Reg), lbo, sbo), umma_desc(smem_u32(Id)), idesc, 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after();
    float v0[32], v1[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v0);
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 32, v1);
    for (int j = 0; j < 32; ++j) out[tid * 64 + j] = v0[j], out[tid * 64 + 32 + j] = v1[j];
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem);
}

int main(int argc, char **argv) {
    float *out;
    cudaMalloc(&out, 2 * 128 * 64 * sizeof(float));
    static float h[2][128 * 64];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 16384);
    const uint32_t combos[][2] = {{1024, 2048}, {4096, 512}};
    const int lts[] = {1, 0, 4, 6, 2};
    for (int lt : lts)
    for (int mode = 1; mode >= 0; --mode)
        for (auto &c : combos) {
            for (int pass = 0; pass < 2; ++pass) {
                probe<<<1, 128, 65536 + 16384>>>(mode, c[0], c[1], pass, out + pass * 128 * 64, lt);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("mode %d lbo %u sbo %u: %s\n", mode, c[0], c[1], cudaGetErrorString(e));
                    return 1;
                }
            }
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            printf("== layout_type %d mode %d (%s, modes 0/1 MN-major, 2/3 K-major control) LBO %u SBO %u: byte offset read for (mn, k)\n", lt, mode, (mode == 0 || mode == 2) ? "B" : "A", c[0], c[1]);
            const int mns[] = {0, 1, 3, 4, 8, 28, 31, 32, 33, 36, 63, 64, 96, 127};
            for (int mn : mns) {
                if ((mode == 0 || mode == 2) && mn >= 64) continue;
                printf("  mn %3d:", mn);
                for (int k = 0; k < 8; ++k) {
                    // mode 0: D[m = k][n = mn]; mode 1: D[m = mn][n = k]
                    const int idx = (mode == 0 || mode == 2) ? k * 64 + mn : mn * 64 + k;
                    const long w = lroundf(h[0][idx]) + 1024 * lroundf(h[1][idx]);
                    printf(" %6ld", w * 4);
                }
                printf("\n");
            }
        }
    return 0;
}
