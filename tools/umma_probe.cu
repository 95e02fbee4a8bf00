// Microbenchmark (development aid, not product code): issue/complete cost of tcgen05.mma kind::f16 for small tiles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I <pkg>/csrc -I include tools/umma_probe.cu -o gpurun_out/umma_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace srnn::ptx;
namespace srnn { int make_tmap_bf16(CUtensorMap*, const void*, uint64_t, uint64_t, uint64_t, uint32_t) { return 0; } }

struct Res { long long issue, total; };

template <int M, int N>
__device__ void probe(uint8_t* smem, uint32_t tmem, uint64_t* bar, int n, int nacc, int interleave, int commit_every, Res* out, uint32_t& phase,
                      int vary = 0, int fence = 0) {
    const uint32_t idesc = umma_idesc_bf16(M, N);
    long long t0 = clock64();
    int commits = 0;
    for (int j = 0; j < n; ++j) {
        const int a = j % nacc;
        const int kb = j >> 2;
        if (fence && (j & 3) == 0) tc_fence_after();
        const uint64_t da = umma_desc_sw128(smem_u32(smem + (vary ? (kb % 16) * 8192 : 0)));
        const uint64_t db = umma_desc_sw128(smem_u32(smem + 131072 + (vary ? (kb % 8) * 4096 : 0)));
        uint32_t d = interleave ? tmem + ((uint32_t)((a & 1) * 16) << 16) + (a >> 1) * N : tmem + a * N;
        umma_bf16(d, da + 2 * (j & 3), db + 2 * (j & 3), idesc, j >= nacc);
        if (commit_every && (j % commit_every) == commit_every - 1 && j != n - 1) { umma_commit(bar + 1); ++commits; }
    }
    long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, phase);
    phase ^= 1;
    long long t2 = clock64();
    out->issue = t1 - t0;
    out->total = t2 - t0;
}

__global__ void k_probe(Res* res) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[2];
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1000000); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        uint32_t ph = 0;
        int r = 0;
        for (int rep = 0; rep < 2; ++rep) {   // rep 0 = warm-up
            r = 0;
            probe<64, 32>(smem, tmem, bar, 64, 1, 0, 0, res + r++, ph);     // 0: M64 N32 same accumulator
            probe<64, 32>(smem, tmem, bar, 64, 8, 0, 0, res + r++, ph);     // 1: 8 accumulators (separate columns)
            probe<64, 32>(smem, tmem, bar, 64, 8, 1, 0, res + r++, ph);     // 2: 8 accumulators interleaved lanes
            probe<64, 32>(smem, tmem, bar, 64, 1, 0, 4, res + r++, ph);     // 3: same acc, commit every 4
            probe<128, 32>(smem, tmem, bar, 64, 1, 0, 0, res + r++, ph);    // 4
            probe<128, 32>(smem, tmem, bar, 64, 8, 0, 0, res + r++, ph);    // 5
            probe<128, 64>(smem, tmem, bar, 64, 1, 0, 0, res + r++, ph);    // 6
            probe<128, 256>(smem, tmem, bar, 64, 1, 0, 0, res + r++, ph);   // 7
            probe<64, 64>(smem, tmem, bar, 64, 1, 0, 0, res + r++, ph);     // 8
            probe<64, 256>(smem, tmem, bar, 64, 1, 0, 0, res + r++, ph);    // 9
            probe<64, 32>(smem, tmem, bar, 1, 1, 0, 0, res + r++, ph);      // 10: single MMA latency
            probe<128, 256>(smem, tmem, bar, 1, 1, 0, 0, res + r++, ph);    // 11
            probe<64, 32>(smem, tmem, bar, 0, 1, 0, 0, res + r++, ph);      // 12: empty commit
            probe<64, 32>(smem, tmem, bar, 64, 1, 0, 4, res + r++, ph, 1, 0);   // 13: vary operand tiles, commit/4
            probe<64, 32>(smem, tmem, bar, 64, 1, 0, 4, res + r++, ph, 1, 1);   // 14: + fence::after_thread_sync per k-block
            probe<128, 32>(smem, tmem, bar, 64, 1, 0, 4, res + r++, ph, 1, 0);  // 15: M128 vary
            probe<64, 32>(smem, tmem, bar, 64, 1, 0, 4, res + r++, ph, 0, 1);   // 16: fixed tiles + fence
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main() {
    Res* d; cudaMalloc(&d, sizeof(Res) * 32);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    k_probe<<<1, 128, 220 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    Res h[32]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[] = {"M64 N32 x64 same acc", "M64 N32 x64 8 acc cols", "M64 N32 x64 8 acc interleaved", "M64 N32 x64 commit/4",
                           "M128 N32 x64 same", "M128 N32 x64 8 acc", "M128 N64 x64 same", "M128 N256 x64 same", "M64 N64 x64 same",
                           "M64 N256 x64 same", "M64 N32 x1", "M128 N256 x1", "empty commit", "M64 N32 vary tiles commit/4",
                           "M64 N32 vary + fence", "M128 N32 vary commit/4", "M64 N32 fixed + fence"};
    for (int i = 0; i < 17; ++i) printf("%-32s issue=%6lld cyc  total=%6lld cyc\n", names[i], h[i].issue, h[i].total);
    return 0;
}
