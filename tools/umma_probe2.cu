// Microbenchmark 2 (development aid): does issuing tcgen05.mma from several warps raise the small-tile issue rate?
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace srnn::ptx;
namespace srnn { int make_tmap_bf16(CUtensorMap*, const void*, uint64_t, uint64_t, uint64_t, uint32_t) { return 0; } }

__global__ void k_probe(long long* res) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[8];
    __shared__ uint32_t slot;
    __shared__ long long tstart[8], tend[8];
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t idesc = umma_idesc_bf16(64, 32);
    int r = 0;
    uint32_t uses = 0;   // per-warp count of completed phases of bar[warp]
    for (int nw = 1; nw <= 8; nw *= 2) {           // number of issuing warps; 64 MMAs in total
        for (int rep = 0; rep < 3; ++rep) {
            __syncthreads();
            if (warp < nw && lane == 0) {
                const uint64_t dA0 = umma_desc_sw128(smem_u32(smem));
                const uint64_t dB0 = umma_desc_sw128(smem_u32(smem + 131072));
                const int per = 64 / nw;
                long long t0 = clock64();
                for (int j = 0; j < per; ++j) {
                    const int g = warp * per + j, kb = g >> 2;
                    umma_bf16(tmem + warp * 32, dA0 + (kb & 15) * 512 + 2 * (g & 3), dB0 + (kb & 7) * 256 + 2 * (g & 3), idesc, j > 0);
                }
                umma_commit(&bar[warp]);
                mbar_wait(&bar[warp], uses & 1);
                ++uses;
                long long t1 = clock64();
                tstart[warp] = t0; tend[warp] = t1;
            }
            __syncthreads();
            if (threadIdx.x == 0 && rep == 2) {
                long long a = tstart[0], b = tend[0];
                for (int w = 1; w < nw; ++w) { if (tstart[w] < a) a = tstart[w]; if (tend[w] > b) b = tend[w]; }
                res[r] = b - a;
            }
        }
        ++r;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main() {
    long long* d; cudaMalloc(&d, sizeof(long long) * 8);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    k_probe<<<1, 256, 220 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    long long h[8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int nw = 1;
    for (int i = 0; i < 4; ++i, nw *= 2) printf("64 x (M64 N32 K16) from %d issuing warp(s): %lld cycles\n", nw, h[i]);
    return 0;
}
