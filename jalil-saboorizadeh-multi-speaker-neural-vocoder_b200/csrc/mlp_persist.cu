// Persistent fused sample-level kernel for generation (SRNN_MODE_BF16): for `nsteps` consecutive audio samples it
// runs  table-gather -> ReLU -> hidden GEMM -> ReLU -> output GEMM -> log-softmax -> inverse-CDF sample  with NO
// host round trip and NO per-step launch (replaces model.py:504-517 executed once per sample by the reference).
//
// Decomposition (H = dim, NS = H/64 feature slices, 32 utterance rows per group, RPC = 32/NS rows owned per CTA):
//   CTA (rg, sl) keeps the bf16 weight slices resident in shared memory for the whole launch:
//       W_hid[sl*64 .. +64, :]   (64 x H,  the UMMA A operand of the hidden GEMM, M = 64)
//       W_out[:, sl*64 .. +64]   (256 x 64, the A operand of a split-K output GEMM, 2 x M = 128)
//   and OWNS rows  rg*32 + sl*RPC .. +RPC  for everything that is per-row (table gathers, softmax, sampling).
//   per step:  owner rows: x1 = relu(P + Tbl[FS-1][newest sample])           -> global X1 (bf16)      [CUDA cores]
//              -- group barrier A (the NS CTAs of a row group) --
//              D1[64 feat x 32 rows]  = W_hid slice . X1(32 rows)^T           (TMA ring -> tcgen05, TMEM)
//              x2 = relu(D1 + b_hid) -> smem (bf16, swizzled B operand)        [epilogue warps]
//              D2[256 x 32 rows]      = W_out[:, slice] . x2^T  (split-K partial logits) -> global Part
//              -- group barrier B --
//              owner rows: logits = b_out + sum_slices Part; log-softmax; defined sampler -> seq      [CUDA cores]
//   The (FS-1)-tap part of the gather for the NEXT step (P) is prefetched by 4 dedicated warps during the GEMMs, so
//   only one table row per utterance is on the serial path.
// Warp roles (416 threads): 0..3 = gather/"G" warps, 4..7 = epilogue/"E" warps, 8 = TMA producer,
// 9..12 = MMA issuers (warp 9 also owns the TMEM allocation).  One thread can only issue a small-tile tcgen05.mma
// every ~45-90 cycles (measured, tools/umma_probe*.cu), far above the 16-cycle tensor floor of an M=64,N=32 tile, so
// the K loop is split over FOUR issuing threads, each with private accumulators (k-block kb belongs to issuer kb%4).
// The SM's issue arbiter favours the highest warp id, so the single-thread latency-critical roles sit on top and
// every wait in the bulk warps is an mbarrier try_wait (hardware back-off), never a hot shared-memory spin.
#include "common.cuh"
#include "sampler.cuh"
#include "umma.cuh"

namespace srnn {

using namespace ptx;

constexpr int MP_THREADS = 416;
constexpr int MP_ISSUERS = 4;
constexpr int MP_MAX_STAGES = 16;   // one stage per k-block at H = 1024: the WHOLE X1 tile of a step is in flight at once
// Back-to-back tcgen05.mma that accumulate into the SAME TMEM tile serialise on the accumulator (~150 cycles each,
// measured), which dominates with N = 32 columns.  The K loop is therefore spread round-robin over NACC1 independent
// accumulators that the epilogue sums.  Two M=64 accumulators share a column range (lanes 0-15 / 16-31 of each
// quadrant, the "interleaved" M=64 TMEM layout), so one 32-lane tcgen05.ld fetches two of them.
constexpr int MP_NACC1 = 8;                      // hidden GEMM accumulators (4 column ranges of 32)
constexpr int MP_D2_COL = (MP_NACC1 / 2) * 32;   // output GEMM: 2 tiles x 2 accumulators x 32 columns
constexpr uint32_t MP_TMEM_COLS = 512;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
// barrier among the NS CTAs of one row group; executed by the 128 E threads (named barrier 1)
__device__ __forceinline__ void group_barrier(unsigned* ctr, unsigned target, int tidE) {
    named_bar_sync(1, 128);
    if (tidE == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(ctr) : "memory");   // publishes the CTA's writes
        while (ld_acquire_gpu(ctr) < target) {
        }
    }
    named_bar_sync(1, 128);
}

__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

#define MP_TRACE(slot)                                                                   \
    do {                                                                                 \
        if (p.trace && blockIdx.x == 0 && tidE == 0) p.trace[k * 64 + (slot)] = clock64(); \
    } while (0)

__global__ void __launch_bounds__(MP_THREADS, 1)
k_mlp_persist(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWo,
              const __grid_constant__ CUtensorMap tmX1, const MlpPersistParams p) {
    const int H = p.H, KB = H >> 6, NS = H >> 6, RPC = 32 / NS, FS = p.FS;
    // The TMA ring holds every k-block of X1 (KB x 4 KB): with 8 stages the second half of the tile could only be
    // requested after the first half had been consumed, i.e. two L2 round trips per step on the serial path (measured
    // 4.4k cycles for 64 KB).  The 32 KB this costs are recovered by NOT keeping the W_out slice resident: it is re-fetched
    // from L2 every step into the first 32 KB of the ring as soon as the hidden GEMM has consumed X1 (during epilogue 1).
    const int NSTG = KB;
    const int nissue = KB < MP_ISSUERS ? KB : MP_ISSUERS;     // MMA-issuing threads in use
    const int nacc = 2 * nissue;                              // hidden-GEMM accumulators in use (2 per issuer)
    const int rg = blockIdx.x / NS, sl = blockIdx.x % NS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sWh = smem;                                  // KB x (64 rows x 128 B)
    uint8_t* sRing = sWh + (size_t)KB * 8192;             // X1: KB x (32 rows x 128 B); then W_out: 2 x (128 rows x 128 B)
    uint8_t* sWo = sRing;                                 // (time-multiplexed with the X1 stages, at least 32 KB)
    const size_t ring_bytes = (size_t)KB * 4096 > 32768 ? (size_t)KB * 4096 : 32768;
    uint8_t* sX2 = sRing + ring_bytes;                    // 32 rows x 128 B
    float* sP = (float*)(sX2 + 4096);                     // 2 x 2048 fp32
    uint8_t* sQ = (uint8_t*)(sP + 4096);                  // 32 owned rows x 32-entry sample ring
    float* sU = (float*)(sQ + 1024);                      // this step's uniform of every owned row (prefetched)
    float* sLogit = sU + 32;                              // RPC owned rows x 256 reduced logits
    uint64_t* bars = (uint64_t*)(sLogit + (size_t)RPC * SRNN_Q);
    uint64_t* w_ready = bars + 0;
    uint64_t* full = bars + 1;                            // [MP_MAX_STAGES]
    uint64_t* empty = full + MP_MAX_STAGES;               // [MP_MAX_STAGES]
    uint64_t* x1_ready = empty + MP_MAX_STAGES;           // (re-used as "W_out slice landed" barrier)
    uint64_t* bar_d1 = x1_ready + 1;
    uint64_t* x2_ready = bar_d1 + 1;
    uint64_t* bar_d2 = x2_ready + 1;
    uint64_t* p_ready = bar_d2 + 1;                       // [2]
    uint64_t* p_free = p_ready + 2;                       // [2]
    uint64_t* q_ready = p_free + 2;                       // [3] sample of step k has been drawn (barrier k % 3)
    uint32_t* tmem_slot = (uint32_t*)(q_ready + 3);

    const int i0 = *p.step_base + p.pos0;                 // absolute index of the first sample of this launch
    const int row0 = rg * 32 + sl * RPC;                  // first owned row (global utterance index)
    // The group counters are never reset between launches (a memset node in front of every launch sat on the serial path):
    // launch m of a generation call -- first sample i0 = lookback + m * nsteps -- starts from the 2 * nsteps * NS arrivals per
    // launch that its predecessors left behind.  The caller zeroes the counters once per generation call.
    const unsigned cbase = (unsigned)((i0 - p.lookback) / p.nsteps) * 2u * (unsigned)p.nsteps * (unsigned)NS;
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[63] = clock64();

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmWh);
        prefetch_tmap(&tmWo);
        prefetch_tmap(&tmX1);
        mbar_init(w_ready, 1);
        for (int s = 0; s < MP_MAX_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(x1_ready, 1);
        mbar_init(bar_d1, nissue);
        mbar_init(x2_ready, 1);
        mbar_init(bar_d2, nissue);
        mbar_init(&p_ready[0], 128);
        mbar_init(&p_ready[1], 128);
        mbar_init(&p_free[0], 1);
        mbar_init(&p_free[1], 1);
        mbar_init(&q_ready[0], 1);
        mbar_init(&q_ready[1], 1);
        mbar_init(&q_ready[2], 1);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc<MP_TMEM_COLS>(tmem_slot);
    // owned rows' sample ring: the FS most recent samples before i0 (written by earlier launches / the q_zero prefix)
    for (int e = threadIdx.x; e < RPC * 32; e += MP_THREADS) {
        const int rl = e >> 5, w = e & 31;
        const int b = row0 + rl;
        const int a = i0 - 32 + w;                        // absolute sample index, slot a & 31
        uint8_t q = 128;
        if (b < p.B && a >= 0 && a >= i0 - FS) q = __ldcg(p.seq + (size_t)b * p.Lseq + a);
        sQ[rl * 32 + (a & 31)] = q;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // warp-uniform copy (shfl from lane 0) so the MMA operands stay in uniform registers: otherwise the compiler wraps
    // every tcgen05.mma of the single issuing thread in an ELECT/R2UR.BROADCAST loop
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t tm_d1 = tmem, tm_d2 = tmem + MP_D2_COL;
    (void)NSTG;
    const int ngrp = KB < MP_ISSUERS ? KB : MP_ISSUERS;       // X1 arrives in ngrp TMA boxes of gsz k-blocks (KB is a power of two)
    const int gsz = KB / ngrp, gshift = 31 - __clz(gsz);

    if (warp == 8) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(w_ready, (uint32_t)(KB * 8192));
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(sWh + (size_t)kb * 8192, &tmWh, w_ready, kb * 64, sl * 64);
            for (int k = 0; k < p.nsteps; ++k) {
                // the ring is free again once the output GEMM of the previous step has read W_out (which also implies that
                // its hidden GEMM is done with the X1 stages)
                if (k) mbar_wait(bar_d2, (k - 1) & 1);
                {   // group barrier A, waiting side: arrivals so far = (2k+1) * NS once every slice has published X1
                    const unsigned target = cbase + (unsigned)(2 * k + 1) * (unsigned)NS;
                    const unsigned* ctr = p.ctr + rg;
                    while (ld_acquire_gpu(ctr) < target) {
                    }
                    fence_proxy_async_all();              // other CTAs' generic-proxy global writes -> visible to TMA reads
                }
                if (p.trace && blockIdx.x == 0) p.trace[k * 64 + 10] = clock64();
                for (int g = 0; g < ngrp; ++g) {          // ngrp TMA instructions of gsz k-blocks each
                    mbar_expect_tx(&full[g], (uint32_t)gsz * 4096);
                    tma_load_3d(sRing + (size_t)g * gsz * 4096, &tmX1, &full[g], 0, rg * 32, g * gsz);
                }
                if (p.trace && blockIdx.x == 0) p.trace[k * 64 + 11] = clock64();
                mbar_wait(bar_d1, k & 1);                 // hidden GEMM complete: X1 stages consumed
                mbar_expect_tx(x1_ready, 32768);
                tma_load_2d(sWo, &tmWo, x1_ready, sl * 64, 0);
                tma_load_2d(sWo + 16384, &tmWo, x1_ready, sl * 64, 128);
            }
        }
    } else if (warp >= 9) {
        // ===================== MMA issuers (one thread each) =====================
        const int w = warp - 9;
        if (lane == 0 && w < nissue) {
            constexpr uint32_t idesc1 = umma_idesc_bf16(64, 32);
            constexpr uint32_t idesc2 = umma_idesc_bf16(128, 32);
            mbar_wait(w_ready, 0);
            const uint64_t dA0 = umma_desc_sw128(smem_u32(sWh));      // + kb * (8192 >> 4)
            const uint64_t dB0 = umma_desc_sw128(smem_u32(sRing));    // + stage * (4096 >> 4)
            const uint64_t dWo0 = umma_desc_sw128(smem_u32(sWo));     // + tile * (16384 >> 4)
            const uint64_t db2 = umma_desc_sw128(smem_u32(sX2));
            const uint32_t d1a = tm_d1 + (uint32_t)w * 32;            // this issuer's two interleaved M=64 accumulators
            const uint32_t d1b = d1a + (16u << 16);
            for (int k = 0; k < p.nsteps; ++k) {
                for (int kb = w; kb < KB; kb += MP_ISSUERS) {
                    mbar_wait(&full[kb >> gshift], k & 1);
                    tc_fence_after();
                    const uint64_t da = dA0 + (uint64_t)(kb * 512);
                    const uint64_t db = dB0 + (uint64_t)(kb * 256);
                    const uint32_t acc = kb >= MP_ISSUERS;
                    umma_bf16(d1a, da, db, idesc1, acc);
                    umma_bf16(d1b, da + 2, db + 2, idesc1, acc);
                    umma_bf16(d1a, da + 4, db + 4, idesc1, 1);
                    umma_bf16(d1b, da + 6, db + 6, idesc1, 1);
                }
                umma_commit(bar_d1);
                if (p.trace && blockIdx.x == 0 && w == 0) p.trace[k * 64 + 14] = clock64();
                mbar_wait(x2_ready, k & 1);
                mbar_wait(x1_ready, k & 1);               // this step's W_out slice has landed in the ring
                tc_fence_after();
                for (int c = w; c < 4; c += nissue) {                 // (output tile, K half) -> its own accumulator
                    const int t2 = c >> 1, h2 = c & 1;
                    const uint64_t da2 = dWo0 + (uint64_t)(t2 * 1024) + (uint64_t)(h2 * 4);
                    const uint32_t d2 = tm_d2 + 64 * t2 + 32 * h2;
                    umma_bf16(d2, da2, db2 + (uint64_t)(h2 * 4), idesc2, 0);
                    umma_bf16(d2, da2 + 2, db2 + (uint64_t)(h2 * 4) + 2, idesc2, 1);
                }
                umma_commit(bar_d2);
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================== E warps: per-row work, epilogues, group barriers =====================
        const int tidE = threadIdx.x - 128;
        const int q4 = warp & 3;                          // TMEM lane quadrant this warp may access
        const int flat = tidE * 16;                       // 16 consecutive features of one owned row
        const int rl = flat / H, f0 = flat % H;
        const int b = row0 + rl;
        unsigned bar_no = 0;
        unsigned* ctr = p.ctr + rg;
        // the owned rows' uniforms are fetched one step ahead (a DRAM miss would otherwise sit on the serial path)
        const int ub = row0 + tidE;
        const bool u_mine = tidE < RPC && ub < p.B;
        float u_next = u_mine ? __ldg(p.uniforms + (size_t)(i0 - p.lookback) * p.u_ld + ub) : 0.f;
        for (int k = 0; k < p.nsteps; ++k) {
            const int i = i0 + k;
            // ---- E1: x1 = relu(P + Tbl[FS-1][newest sample]) for the owned rows -> global X1 ----
            MP_TRACE(0);
            if (tidE < RPC) {
                sU[tidE] = u_next;
                if (u_mine && k + 1 < p.nsteps) u_next = __ldg(p.uniforms + (size_t)(i + 1 - p.lookback) * p.u_ld + ub);
            }
            mbar_wait(&p_ready[k & 1], (k >> 1) & 1);
            MP_TRACE(1);
            {
                const int qn = sQ[rl * 32 + ((i - 1) & 31)], qm = sQ[rl * 32 + ((i - 2) & 31)];
                const uint4* tp = reinterpret_cast<const uint4*>(p.tbl + ((size_t)(FS - 1) * SRNN_Q + qn) * H + f0);
                const uint4* tq = reinterpret_cast<const uint4*>(p.tbl + ((size_t)(FS - 2) * SRNN_Q + qm) * H + f0);
                const uint4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tq), t3 = __ldg(tq + 1);
                float tv[16], tw[16];
                bf16x8_to_f32(t0, tv);
                bf16x8_to_f32(t1, tv + 8);
                bf16x8_to_f32(t2, tw);
                bf16x8_to_f32(t3, tw + 8);
#pragma unroll
                for (int v = 0; v < 16; ++v) tv[v] += tw[v];
                const float4* pp = reinterpret_cast<const float4*>(sP + (k & 1) * 2048 + flat);
                uint32_t o[8];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float4 pv = pp[v];
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(fmaxf(pv.x + tv[4 * v], 0.f), fmaxf(pv.y + tv[4 * v + 1], 0.f));
                    const __nv_bfloat162 hi = __floats2bfloat162_rn(fmaxf(pv.z + tv[4 * v + 2], 0.f), fmaxf(pv.w + tv[4 * v + 3], 0.f));
                    o[2 * v] = *reinterpret_cast<const uint32_t*>(&lo);
                    o[2 * v + 1] = *reinterpret_cast<const uint32_t*>(&hi);
                }
                uint4* xp = reinterpret_cast<uint4*>(p.x1 + (size_t)b * H + f0);   // x1 has RG*32 rows: always in range
                xp[0] = make_uint4(o[0], o[1], o[2], o[3]);
                xp[1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
            // ---- group barrier A: the whole X1 of this row group is in global memory ----
            MP_TRACE(2);
            // E warps only ARRIVE (release); the TMA thread is the one that waits for the other slices' arrivals, so the
            // wake-up hand-off between two warps is off the serial path.  X1 is not rewritten before barrier B.
            named_bar_sync(1, 128);
            ++bar_no;
            if (tidE == 0) {
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(ctr) : "memory");
                mbar_arrive(&p_free[k & 1]);              // P[k&1] consumed by every E thread
            }
            MP_TRACE(3);
            // ---- epilogue 1: D1 (+bias, ReLU) -> bf16 swizzled B operand in smem ----
            mbar_wait(bar_d1, k & 1);
            MP_TRACE(4);
            tc_fence_after();
            {
                float v0[16], v1[16];
                tmem_ld16(tm_d1 + ((uint32_t)(32 * q4) << 16), v0);
                tmem_ld16(tm_d1 + ((uint32_t)(32 * q4) << 16) + 16, v1);
#pragma unroll 1
                for (int c = 1; c < nacc / 2; ++c) {
                    float w0[16], w1[16];
                    tmem_ld16(tm_d1 + ((uint32_t)(32 * q4) << 16) + 32 * c, w0);
                    tmem_ld16(tm_d1 + ((uint32_t)(32 * q4) << 16) + 32 * c + 16, w1);
#pragma unroll
                    for (int n = 0; n < 16; ++n) {
                        v0[n] += w0[n];
                        v1[n] += w1[n];
                    }
                }
#pragma unroll
                for (int n = 0; n < 16; ++n) {            // lanes 16-31 hold the odd accumulators of the same features
                    v0[n] += __shfl_down_sync(0xffffffffu, v0[n], 16);
                    v1[n] += __shfl_down_sync(0xffffffffu, v1[n], 16);
                }
                if (lane < 16) {
                    const int f = 16 * q4 + lane;         // feature inside the slice = K index of the output GEMM
                    const float bv = p.b_hid[sl * 64 + f];
                    uint8_t* base = sX2 + (f & 7) * 2;
                    const int chunk = f >> 3;
#pragma unroll
                    for (int n = 0; n < 16; ++n) {
                        *reinterpret_cast<__nv_bfloat16*>(base + n * 128 + ((chunk ^ (n & 7)) << 4)) =
                            __float2bfloat16(fmaxf(v0[n] + bv, 0.f));
                        *reinterpret_cast<__nv_bfloat16*>(base + (n + 16) * 128 + ((chunk ^ ((n + 16) & 7)) << 4)) =
                            __float2bfloat16(fmaxf(v1[n] + bv, 0.f));
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();                     // generic smem writes -> visible to the UMMA operand reads
            named_bar_sync(1, 128);
            if (tidE == 0) mbar_arrive(x2_ready);
            MP_TRACE(5);
            // ---- epilogue 2: split-K partial logits -> global Part[rg][sl][row][256] ----
            mbar_wait(bar_d2, k & 1);
            MP_TRACE(6);
            tc_fence_after();
            {
                float* dst = p.part + ((size_t)(rg * NS + sl) * 32) * SRNN_Q + 32 * q4 + lane;
#pragma unroll
                for (int t2 = 0; t2 < 2; ++t2) {
                    float v0[16], v1[16], w0[16], w1[16];
                    tmem_ld16(tm_d2 + 64 * t2 + ((uint32_t)(32 * q4) << 16), v0);
                    tmem_ld16(tm_d2 + 64 * t2 + ((uint32_t)(32 * q4) << 16) + 16, v1);
                    tmem_ld16(tm_d2 + 64 * t2 + 32 + ((uint32_t)(32 * q4) << 16), w0);
                    tmem_ld16(tm_d2 + 64 * t2 + 32 + ((uint32_t)(32 * q4) << 16) + 16, w1);
#pragma unroll
                    for (int n = 0; n < 16; ++n) {
                        dst[(size_t)n * SRNN_Q + t2 * 128] = v0[n] + w0[n];
                        dst[(size_t)(n + 16) * SRNN_Q + t2 * 128] = v1[n] + w1[n];
                    }
                }
            }
            tc_fence_before();
            // ---- group barrier B: all slices' partial logits are in global memory ----
            MP_TRACE(7);
            group_barrier(ctr, cbase + (++bar_no) * NS, tidE);
            MP_TRACE(8);
            // ---- reduce the NS split-K partials of the owned rows (all 128 E threads, every load in flight) ----
            {
                const int per_row = SRNN_Q / 4;                                   // 64 threads cover one row's 256 logits
                for (int e = tidE; e < RPC * per_row; e += 128) {
                    const int r2 = e / per_row, o4 = (e % per_row) * 4;
                    const int n = sl * RPC + r2;
                    const float* src = p.part + ((size_t)(rg * NS) * 32 + n) * SRNN_Q + o4;
                    float4 acc = __ldg(reinterpret_cast<const float4*>(p.b_out + o4));
                    for (int s0 = 0; s0 < NS; s0 += 16) {
                        float4 a[16];
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            const int s2 = (s0 + u < NS) ? s0 + u : NS - 1;
                            a[u] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)s2 * 32 * SRNN_Q));
                        }
#pragma unroll
                        for (int u = 0; u < 16; ++u)                              // fixed summation order: slice 0, 1, 2, ...
                            if (s0 + u < NS) {
                                acc.x += a[u].x; acc.y += a[u].y; acc.z += a[u].z; acc.w += a[u].w;
                            }
                    }
                    *reinterpret_cast<float4*>(sLogit + r2 * SRNN_Q + o4) = acc;
                }
            }
            named_bar_sync(1, 128);
            // ---- log-softmax + defined sampler, one warp per owned row ----
            for (int r2 = warp - 4; r2 < RPC; r2 += 4) {
                const int bb = row0 + r2;
                float v[8];
                {
                    const float4* lp = reinterpret_cast<const float4*>(sLogit + r2 * SRNN_Q + lane * 8);
                    const float4 a0 = lp[0], a1 = lp[1];
                    v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
                }
                float m = v[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) m = fmaxf(m, v[j]);
                for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) s += expf(v[j] - m);
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const float lse = m + logf(s);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] -= lse;
                const int t = i - p.lookback;
                if (bb < p.B) {
                    if (p.logp_out) {
                        float4* o4 = reinterpret_cast<float4*>(p.logp_out + ((size_t)bb * p.T + t) * SRNN_Q + lane * 8);
                        o4[0] = make_float4(v[0], v[1], v[2], v[3]);
                        o4[1] = make_float4(v[4], v[5], v[6], v[7]);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = expf(v[j]);
                    const float u = sU[r2];
                    const int idx = sampler_warp(v, u, lane);
                    if (lane == 0) {
                        p.seq[(size_t)bb * p.Lseq + i] = (uint8_t)idx;
                        sQ[r2 * 32 + (i & 31)] = (uint8_t)idx;
                    }
                } else if (lane == 0) {
                    sQ[r2 * 32 + (i & 31)] = 128;
                }
            }
            named_bar_sync(1, 128);                       // sample i of every owned row is in sQ
            if (tidE == 0) mbar_arrive(&q_ready[k % 3]);
            MP_TRACE(9);
        }
    } else if (warp < 4) {
        // ===================== G warps: prefetch P_g = c0 + taps 0..FS-3 for step g (two steps of slack) ==========
        const int tidG = threadIdx.x;
        const int flat = tidG * 16;
        const int rl = flat / H, f0 = flat % H;
        const int b = row0 + rl;
        const int bc = b < p.B ? b : p.B - 1;             // clamp: padded rows compute garbage that is never used
        for (int g = 0; g < p.nsteps; ++g) {
            const int i = i0 + g;
            if (g >= 3)   // sample i-3 (tap FS-3, the newest one this prefetch uses) has been drawn.  Three barriers by
                          // step % 3: E can be up to two steps past the awaited one, which would alias a phase parity.
                mbar_wait(&q_ready[g % 3], (g / 3 - 1) & 1);
            if (g >= 2) mbar_wait(&p_free[g & 1], ((g >> 1) - 1) & 1);
            float acc[16];
            {
                const float4* cp = reinterpret_cast<const float4*>(p.c0 + (size_t)bc * FS * H + (size_t)(i % FS) * H + f0);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float4 c = __ldg(cp + v);
                    acc[4 * v] = c.x; acc[4 * v + 1] = c.y; acc[4 * v + 2] = c.z; acc[4 * v + 3] = c.w;
                }
            }
            const uint8_t* qrow = sQ + rl * 32;
#pragma unroll 6
            for (int j = 0; j < FS - 2; ++j) {
                const int qj = qrow[(i - FS + j) & 31];
                const uint4* tp = reinterpret_cast<const uint4*>(p.tbl + ((size_t)j * SRNN_Q + qj) * H + f0);
                const uint4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
                float tv[16];
                bf16x8_to_f32(t0, tv);
                bf16x8_to_f32(t1, tv + 8);
#pragma unroll
                for (int v = 0; v < 16; ++v) acc[v] += tv[v];
            }
            float4* pp = reinterpret_cast<float4*>(sP + (g & 1) * 2048 + flat);
#pragma unroll
            for (int v = 0; v < 4; ++v) pp[v] = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
            mbar_arrive(&p_ready[g & 1]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc<MP_TMEM_COLS>(tmem);
}

size_t mlp_persist_smem(int H) {
    const int KB = H / 64;
    const int RPC = 32 / (H / 64);
    const size_t ring = (size_t)KB * 4096 > 32768 ? (size_t)KB * 4096 : 32768;
    return (size_t)KB * 8192 + ring + 4096 + 16384 + 1024 + 128 + (size_t)RPC * 1024 + 512 + 1024;
}

bool mlp_persist_supported(int H, int FS, int B, int n_sms) {
    if (H % 64 || H > 1024) return false;
    const int NS = H / 64;
    if (NS & (NS - 1)) return false;                       // k-block groups of the X1 TMA boxes assume a power of two
    if (32 % NS) return false;
    if (FS < 2 || FS > 31) return false;
    const int RG = (B + 31) / 32;
    return RG * NS <= n_sms;
}

// x1 must hold RG*32 rows; part RG*NS*32*256 floats; ctr RG counters, zeroed by the caller once per generation call
// (launches must start at sample lookback + m * nsteps: the kernel derives its counter base from that).
int mlp_persist_launch(const __nv_bfloat16* w_hid16, const __nv_bfloat16* w_out16, const MlpPersistParams& p,
                       cudaStream_t st) {
    const int H = p.H, NS = H / 64, RG = (p.B + 31) / 32;
    CUtensorMap tmWh, tmWo, tmX1;
    SRNN_TRY(make_tmap_bf16(&tmWh, w_hid16, H, H, H, 64));
    SRNN_TRY(make_tmap_bf16(&tmWo, w_out16, SRNN_Q, H, H, 128));
    const int KBh = H / 64, ngrp = KBh < MP_ISSUERS ? KBh : MP_ISSUERS;
    SRNN_TRY(make_tmap_bf16_kb(&tmX1, p.x1, (uint64_t)RG * 32, H, H, 32, KBh / ngrp));
    const size_t smem = mlp_persist_smem(H);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        SRNN_CUDA(cudaFuncSetAttribute(k_mlp_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(RG * NS);
    cfg.blockDim = dim3(MP_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;          // co-residency of all CTAs (they barrier on each other)
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_mlp_persist, tmWh, tmWo, tmX1, p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail(SRNN_ERR_CUDA, "k_mlp_persist launch: %s", cudaGetErrorString(e));
    return SRNN_OK;
}

}  // namespace srnn
