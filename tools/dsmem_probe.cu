// Probe (development aid): the two per-sample exchanges of the sample-level kernel inside an 8-CTA thread-block cluster.
//   A  all-gather : every CTA owns 2 rows x 1024 bf16 (4 KB) and needs all 16 rows (32 KB) as a swizzled UMMA B tile
//   B  reduce-scatter: every CTA holds partial logits for 16 rows x 256 (fp32) and the owner of 2 rows needs all 8 partials
// Variants of A:  0 = st.async (remote store + complete_tx on the destination's mbarrier)
//                 1 = st.shared::cluster.v4 + mbarrier.arrive.release.cluster on every destination
//                 2 = global staging + cp.async.bulk ... .multicast::cluster (16 copies of 256 B per owner)
// B always uses st.async.b64.  Reports co-residency (cudaOccupancyMaxActiveClusters) and cycles per exchange.
// nvcc -gencode arch=compute_100a,code=sm_100a -I jalil-saboorizadeh-multi-speaker-neural-vocoder_b200/csrc -I include -o tools/dsmem_probe tools/dsmem_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace srnn::ptx;
namespace srnn { int make_tmap_bf16(CUtensorMap*, const void*, uint64_t, uint64_t, uint64_t, uint32_t) { return 0; } }

constexpr int CS = 8;
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint4 v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(raddr),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void st_async_b64(uint32_t raddr, uint64_t v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];\n" ::"r"(raddr), "l"(v), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void arrive_remote_release(uint32_t rbar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(rbar) : "memory");
}
__device__ __forceinline__ void wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "h"(mask)
                 : "memory");
}

__device__ __forceinline__ uint32_t pat(int it, int cta, int row, int f) { return (uint32_t)(it * 7919 + cta * 131 + row * 17 + f); }

__global__ void __launch_bounds__(256, 1) k_probe(int variant, int iters, uint8_t* stage, long long* res, int* errs) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sX1 = smem;                       // [16 kb][16 rows][128 B]
    float* sLand = (float*)(smem + 32768);     // [8 src][256 logits][2 rows]
    uint64_t* bars = (uint64_t*)(smem + 32768 + 16384);
    uint64_t* x1_full = bars;                  // tx-based (variants 0, 2) or 128 * 8 arrivals (variant 1)
    uint64_t* part_full = bars + 1;
    const int tid = threadIdx.x;
    const uint32_t c = cluster_ctarank();
    const int cl = blockIdx.x / CS;
    if (tid == 0) {
        mbar_init(x1_full, variant == 1 ? 128 * CS : 1);
        mbar_init(part_full, 1);
        fence_barrier_init();
    }
    __syncthreads();
    cluster_sync_all();
    long long accA = 0, accB = 0, accT = 0;
    int bad = 0;
    if (tid < 128) {
        const int rl = tid >> 6, f0 = (tid & 63) * 16;             // owned row (0/1), 16 features
        const int n = 2 * (int)c + rl;                             // row inside the 16-row tile
        const int kb = f0 >> 6, ch = (f0 & 63) >> 3;               // k-block, first 16-byte chunk (even)
        const uint32_t off0 = kb * 2048 + (n >> 3) * 1024 + (n & 7) * 128 + ((ch ^ (n & 7)) << 4);
        const uint32_t off1 = kb * 2048 + (n >> 3) * 1024 + (n & 7) * 128 + (((ch + 1) ^ (n & 7)) << 4);
        const uint32_t lx1 = smem_u32(sX1), lbar = smem_u32(x1_full), lland = smem_u32(sLand), lpbar = smem_u32(part_full);
        uint8_t* my_stage = stage + (size_t)blockIdx.x * 4096;      // variant 2: [kb][2 rows][128 B]
        for (int it = 0; it < iters; ++it) {
            const long long t0 = clock64();
            // ---------------- A ----------------
            uint4 v0, v1;
            v0.x = pat(it, c, rl, f0); v0.y = v0.x + 1; v0.z = v0.x + 2; v0.w = v0.x + 3;
            v1.x = v0.x + 4; v1.y = v0.x + 5; v1.z = v0.x + 6; v1.w = v0.x + 7;
            if (variant == 0) {
                if (tid == 0) mbar_expect_tx(x1_full, 32768);
#pragma unroll
                for (int d = 0; d < CS; ++d) {
                    const uint32_t rb = mapa(lbar, d);
                    st_async_v4(mapa(lx1 + off0, d), v0, rb);
                    st_async_v4(mapa(lx1 + off1, d), v1, rb);
                }
                wait_cluster(x1_full, it & 1);
            } else if (variant == 1) {
#pragma unroll
                for (int d = 0; d < CS; ++d) {
                    st_cluster_v4(mapa(lx1 + off0, d), v0);
                    st_cluster_v4(mapa(lx1 + off1, d), v1);
                }
#pragma unroll
                for (int d = 0; d < CS; ++d) arrive_remote_release(mapa(lbar, d));
                wait_cluster(x1_full, it & 1);
            } else if (variant == 4) {
                // smem staging [kb][2 rows][128 B] (destination swizzle) -> 128 bulk copies smem -> peer smem (one per thread:
                // dest = tid >> 4, k-block = tid & 15), each completing 256 B on the destination's barrier
                if (tid == 0) mbar_expect_tx(x1_full, 32768);
                uint8_t* stg = smem + 32768 + 16384 + 1024;
                *reinterpret_cast<uint4*>(stg + kb * 256 + rl * 128 + ((ch ^ (n & 7)) << 4)) = v0;
                *reinterpret_cast<uint4*>(stg + kb * 256 + rl * 128 + (((ch + 1) ^ (n & 7)) << 4)) = v1;
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                asm volatile("bar.sync 1, 128;\n" ::: "memory");
                {
                    const int d = tid >> 4, kq = tid & 15, n0 = 2 * (int)c;
                    const uint32_t dst = mapa(lx1 + kq * 2048 + (n0 >> 3) * 1024 + (n0 & 7) * 128, d);
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                                 "r"(smem_u32(stg) + kq * 256), "r"(256), "r"(mapa(lbar, d)) : "memory");
                }
                wait_cluster(x1_full, it & 1);
            } else {   // variants 2, 3
                if (tid == 0) mbar_expect_tx(x1_full, 32768);
                // staging layout [kb][rl][128 B] with the destination row's swizzle
                *reinterpret_cast<uint4*>(my_stage + kb * 256 + rl * 128 + ((ch ^ (n & 7)) << 4)) = v0;
                *reinterpret_cast<uint4*>(my_stage + kb * 256 + rl * 128 + (((ch + 1) ^ (n & 7)) << 4)) = v1;
                asm volatile("fence.proxy.async;\n" ::: "memory");
                asm volatile("bar.sync 1, 128;\n" ::: "memory");
                if (tid < 16) {
                    const int n0 = 2 * (int)c;
                    bulk_mc(lx1 + tid * 2048 + (n0 >> 3) * 1024 + (n0 & 7) * 128, my_stage + tid * 256, 256, lbar, (uint16_t)0xFF);
                }
                wait_cluster(x1_full, it & 1);
            }
            const long long tA = clock64();
            if (it == iters - 1 || it == 3) {        // verify the whole tile
                for (int e = tid; e < 16 * 64; e += 128) {     // (row, 16-byte chunk index over 1024 features / 8)
                    const int nn = e >> 6, f = (e & 63) * 16;
                    for (int h = 0; h < 2; ++h) {
                        const int ff = f + 8 * h;
                        const int kb2 = ff >> 6, ch2 = (ff & 63) >> 3;
                        const uint4 got = *reinterpret_cast<const uint4*>(sX1 + kb2 * 2048 + (nn >> 3) * 1024 + (nn & 7) * 128 + ((ch2 ^ (nn & 7)) << 4));
                        const uint32_t want = pat(it, nn >> 1, nn & 1, f) + 4 * h;
                        if (got.x != want || got.w != want + 3) ++bad;
                    }
                }
            }
            // ---------------- B ----------------
            if (tid == 0) mbar_expect_tx(part_full, 16384);
            if (variant < 3) {
#pragma unroll
                for (int d = 0; d < CS; ++d) {
                    const uint32_t rb = mapa(lpbar, d);
#pragma unroll
                    for (int t2 = 0; t2 < 2; ++t2) {
                        const int o = t2 * 128 + tid;
                        const float a = (float)(it + (int)c + o + 2 * d), b = a + 0.5f;
                        const uint64_t v = (uint64_t)__float_as_uint(a) | ((uint64_t)__float_as_uint(b) << 32);
                        st_async_b64(mapa(lland + (uint32_t)(((int)c * 256 + o) * 8), d), v, rb);
                    }
                }
            } else {
                // stage [dest][256 logits][2 rows] in local smem (re-using the X1 tile, dead at this point), then one bulk copy per dest
                float* stg = (float*)sX1;
#pragma unroll
                for (int d = 0; d < CS; ++d)
#pragma unroll
                    for (int t2 = 0; t2 < 2; ++t2) {
                        const int o = t2 * 128 + tid;
                        const float a = (float)(it + (int)c + o + 2 * d);
                        *reinterpret_cast<float2*>(stg + (d * 256 + o) * 2) = make_float2(a, a + 0.5f);
                    }
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                asm volatile("bar.sync 1, 128;\n" ::: "memory");
                if (tid < CS) {
                    const uint32_t dst = mapa(lland + (uint32_t)((int)c * 2048), tid), rb = mapa(lpbar, tid);
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                                 "r"(lx1 + tid * 2048), "r"(2048), "r"(rb) : "memory");
                }
            }
            wait_cluster(part_full, it & 1);
            const long long tB = clock64();
            {   // reduce: thread handles logits 2*tid, 2*tid+1 of both rows
                float4 s = make_float4(0, 0, 0, 0);
#pragma unroll
                for (int src = 0; src < CS; ++src) {
                    const float4 p = *reinterpret_cast<const float4*>(sLand + (src * 256 + 2 * tid) * 2);
                    s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
                }
                // expected: sum_src (it + src + o + 2c) for o = 2 tid
                const float want = 8.f * (it + 2 * tid + 2 * (int)c) + 28.f;
                if (s.x != want || s.y != want + 4.f) ++bad;
            }
            asm volatile("bar.sync 1, 128;\n" ::: "memory");
            const long long tE = clock64();
            if (it >= 8) { accA += tA - t0; accB += tB - tA; accT += tE - t0; }
        }
    }
    if (bad) atomicAdd(errs, bad);
    if (blockIdx.x == 0 && tid == 0) { res[0] = accA / (iters - 8); res[1] = accB / (iters - 8); res[2] = accT / (iters - 8); }
    if (blockIdx.x == 8 * 9 + 3 && tid == 0) { res[3] = accA / (iters - 8); res[4] = accB / (iters - 8); res[5] = accT / (iters - 8); }
    (void)cl;
    cluster_sync_all();
}

int main() {
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    uint8_t* stage; long long* dres; int* derr;
    cudaMalloc(&stage, 148 * 4096); cudaMalloc(&dres, 64); cudaMalloc(&derr, 4);
    for (int grid : {120}) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute a[1];
        a[0].id = cudaLaunchAttributeClusterDimension;
        a[0].val.clusterDim.x = CS; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        cfg.attrs = a; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_probe, &cfg);
        printf("grid %d: max active 8-CTA clusters at %d KB smem: %d (%s)\n", grid, smem / 1024, n, cudaGetErrorString(e));
        for (int variant = 0; variant < 5; ++variant) {
            cudaMemset(dres, 0, 64); cudaMemset(derr, 0, 4);
            const int iters = 208;
            e = cudaLaunchKernelEx(&cfg, k_probe, variant, iters, stage, dres, derr);
            cudaError_t e2 = cudaDeviceSynchronize();
            long long h[6]; int herr;
            cudaMemcpy(h, dres, sizeof(h), cudaMemcpyDeviceToHost);
            cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
            printf("  variant %d: launch %s / run %s, errors %d; CTA0: all-gather %lld, reduce-scatter %lld, total %lld cycles"
                   " | CTA75: %lld %lld %lld\n", variant, cudaGetErrorString(e), cudaGetErrorString(e2), herr, h[0], h[1], h[2],
                   h[3], h[4], h[5]);
        }
    }
    return 0;
}
