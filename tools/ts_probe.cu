// Probe (development aid): tcgen05.mma with the A operand in TENSOR MEMORY (the ".ts" form: [d_tmem], [a_tmem], b_desc).
//  1. correctness: D[128 x 16] = A[128 x 64] . B[16 x 64]^T with A written to TMEM by tcgen05.st (lane = row, one 32-bit
//     column = two consecutive K elements), B in shared memory (K-major, 128-byte swizzle), against a host reference.
//  2. issue rate: 64 MMAs of M128 x N16 x K16 from 1 / 2 / 4 issuing warps (private accumulators), A from smem vs TMEM.
// nvcc -gencode arch=compute_100a,code=sm_100a -I jalil-saboorizadeh-multi-speaker-neural-vocoder_b200/csrc -I include -o tools/ts_probe tools/ts_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace srnn::ptx;
namespace srnn { int make_tmap_bf16(CUtensorMap*, const void*, uint64_t, uint64_t, uint64_t, uint32_t) { return 0; } }

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// A (128 x 64) and B (16 x 64) row-major bf16 in global; out D (128 x 16) fp32; res[]: timing results
__global__ void __launch_bounds__(256, 1) k_probe(const __nv_bfloat16* A, const __nv_bfloat16* Bm, float* D, long long* res) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[8];
    __shared__ uint32_t slot;
    __shared__ long long tstart[8], tend[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sB = smem;               // 16 rows x 128 B (one k-block), swizzled; timing: 16 k-blocks x 2 KB
    uint8_t* sA = smem + 32768;       // timing only: 128 rows x 128 B per k-block (zeros)
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    __syncthreads();
    // B tile: row n, K element k -> n*128 + ((k/8) ^ (n & 7)) * 16 + (k % 8) * 2
    for (int e = threadIdx.x; e < 16 * 64; e += blockDim.x) {
        const int n = e / 64, k = e % 64;
        *reinterpret_cast<__nv_bfloat16*>(sB + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = Bm[e];
    }
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
    const uint32_t colA = 0, colD = 448;
    // ---- A -> TMEM: thread (warp w < 4, lane l) owns lane 32w + l = row of A; column c = K elements 2c, 2c+1
    if (warp < 4) {
        const int row = warp * 32 + lane;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(A + (size_t)row * 64);
        uint32_t r0[16], r1[16];
        for (int j = 0; j < 16; ++j) { r0[j] = src[j]; r1[j] = src[16 + j]; }
        const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + colA;
        tmem_st16(ta, r0);
        tmem_st16(ta + 16, r1);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 16);
    uint32_t uses[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (threadIdx.x == 0) {
        const uint64_t dB = umma_desc_sw128(smem_u32(sB));
        for (int j = 0; j < 4; ++j) umma_bf16_ts(tmem + colD, tmem + colA + 8 * j, dB + 2 * j, idesc, j > 0);
        umma_commit(&bar[0]);
        mbar_wait(&bar[0], 0);
    }
    uses[0] = 1;
    __syncthreads();
    tc_fence_after();
    if (warp < 4) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + colD, v);
        for (int n = 0; n < 16; ++n) D[(size_t)(warp * 32 + lane) * 16 + n] = v[n];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // ---- timing: 64 MMAs (16 k-blocks x 4) split over nw issuing warps; mode 0 = A from smem, 1 = A from TMEM,
    //      2 = k-blocks 0..11 from TMEM, 12..15 from smem
    int r = 0;
    for (int nn = 16; nn <= 64; nn *= 2)
    for (int mode = 1; mode < 3; ++mode) {
        const uint32_t idesc_t = umma_idesc_bf16(128, nn);
        for (int nw = 1; nw <= 4; nw *= 2) {
            for (int rep = 0; rep < 3; ++rep) {
                __syncthreads();
                if (warp < nw && lane == 0) {
                    const uint64_t dA0 = umma_desc_sw128(smem_u32(sA));
                    const uint64_t dB0 = umma_desc_sw128(smem_u32(sB));
                    long long t0 = clock64();
                    int cnt = 0;
                    for (int kb = warp; kb < 16; kb += nw) {
                        const bool ts = mode == 1 || (mode == 2 && kb < 12);
                        for (int j = 0; j < 4; ++j, ++cnt) {
                            if (ts)
                                umma_bf16_ts(tmem + colD - 192 + 64 * warp, tmem + colA + (kb % 8) * 32 + 8 * j, dB0 + kb * 128 + 2 * j,
                                             idesc_t, cnt > 0);
                            else
                                umma_bf16(tmem + colD - 192 + 64 * warp, dA0 + (kb & 3) * 1024 + 2 * j, dB0 + kb * 128 + 2 * j, idesc_t,
                                          cnt > 0);
                        }
                    }
                    umma_commit(&bar[warp]);
                    mbar_wait(&bar[warp], uses[warp] & 1);
                    long long t1 = clock64();
                    tstart[warp] = t0; tend[warp] = t1;
                }
                if (warp < nw) ++uses[warp];
                __syncthreads();
                if (threadIdx.x == 0 && rep == 2) {
                    long long a = tstart[0], b = tend[0];
                    for (int w = 1; w < nw; ++w) { if (tstart[w] < a) a = tstart[w]; if (tend[w] > b) b = tend[w]; }
                    res[r] = b - a;
                }
            }
            ++r;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
    const int M = 128, N = 16, K = 64;
    __nv_bfloat16 *hA = new __nv_bfloat16[M * K], *hB = new __nv_bfloat16[N * K];
    float* fa = new float[M * K]; float* fb = new float[N * K];
    srand(1);
    for (int i = 0; i < M * K; ++i) { fa[i] = (float)(rand() % 17 - 8) / 4.f; hA[i] = __float2bfloat16(fa[i]); }
    for (int i = 0; i < N * K; ++i) { fb[i] = (float)(rand() % 13 - 6) / 2.f; hB[i] = __float2bfloat16(fb[i]); }
    __nv_bfloat16 *dA, *dB; float* dD; long long* dres;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dres, 32 * 8);
    cudaMemcpy(dA, hA, M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, N * K * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, M * N * 4);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    k_probe<<<1, 256, 210 * 1024>>>(dA, dB, dD, dres);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    float* hD = new float[M * N];
    cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)fa[m * K + k] * fb[n * K + k];
            double d = s - hD[m * N + n]; if (d < 0) d = -d;
            if (d > maxerr) maxerr = d;
        }
    printf("TS MMA  D = A(tmem) . B(smem)^T  max |err| = %g  (%s)\n", maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
    long long h[32]; cudaMemcpy(h, dres, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[3] = {"A from smem", "A from TMEM", "12 kb TMEM + 4 kb smem"};
    int r = 0;
    for (int nn = 16; nn <= 64; nn *= 2)
        for (int mode = 1; mode < 3; ++mode)
            for (int nw = 1; nw <= 4; nw *= 2) printf("64 x (M128 N%d K16), %-24s %d issuing warp(s): %lld cycles\n", nn, names[mode], nw, h[r++]);
    return 0;
}
