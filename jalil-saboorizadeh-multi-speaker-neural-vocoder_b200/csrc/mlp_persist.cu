// Persistent fused sample-level kernel for generation (SRNN_MODE_BF16): for `nsteps` consecutive audio samples it
// runs  table-gather -> ReLU -> hidden GEMM -> ReLU -> output GEMM -> log-softmax -> inverse-CDF sample  with NO
// host round trip and NO per-step launch (replaces model.py:504-517 executed once per sample by the reference).
//
// Decomposition (H = dim, NS = H/64 feature slices; a row group = 32 utterances = two HALVES of NR = 16 rows):
//   CTA (rg, sl) keeps the bf16 weight slices resident in shared memory for the whole launch:
//       W_hid[sl*64 .. +64, :]   (64 x H,  the UMMA A operand of the hidden GEMM, M = 64)
//       W_out[:, sl*64 .. +64]   (256 x 64, the A operand of a split-K output GEMM, 2 x M = 128)
//   and, in each half, OWNS 16/NS rows for everything that is per-row (table gathers, softmax, sampling).
//   per step and half:
//              owner rows: x1 = relu(P + Tbl[FS-2][.] + Tbl[FS-1][newest sample])  -> global X1 (bf16)   [CUDA cores]
//              -- group barrier A (the NS CTAs of the row group; E warps arrive, the TMA thread waits) --
//              D1[64 feat x 16 rows]  = W_hid slice . X1(16 rows)^T           (TMA ring -> tcgen05, TMEM)
//              x2 = relu(D1 + b_hid) -> smem (bf16, swizzled B operand)        [epilogue warps]
//              D2[256 x 16 rows]      = W_out[:, slice] . x2^T  (split-K partial logits) -> global Part
//              -- group barrier B --
//              owner rows: logits = b_out + sum_slices Part; log-softmax; defined sampler -> seq      [CUDA cores]
//   The serial chain of one sample is dominated by latencies (two cross-CTA exchanges through L2, TMA, MMA issue), so the
//   two halves run as INDEPENDENT warp sets on the same resident weights and the SM interleaves them: while one half
//   waits on a barrier the other one computes.  The taps 0..FS-3 of the table gather for a later step (P) are
//   prefetched by dedicated gather warps two samples ahead.
// Warp roles (704 threads): 0..3 gather "G" (2 per half), 4..7 / 8..11 epilogue "E" of half 0 / 1, 12..13 TMA producers,
// 14..17 / 18..21 MMA issuers of half 0 / 1 (warp 14 owns the TMEM allocation).  One thread can only issue a small-tile
// tcgen05.mma every ~45-90 cycles (tools/umma_probe*.cu), far above the tensor floor of an M=64,N=16 tile, so the K loop
// of a half is split over four issuing threads with private TMEM accumulators (k-block kb belongs to issuer kb % 4).
// The issue arbiter favours the highest warp id, so the single-thread latency-critical roles sit on top; every wait in
// the bulk warps is an mbarrier try_wait (hardware back-off), never a hot shared-memory spin.
#include "common.cuh"
#include "sampler.cuh"
#include "umma.cuh"

namespace srnn {

using namespace ptx;

constexpr int MP_NH = 2;                 // halves per row group
constexpr int MP_NR = 16;                // rows per half = UMMA N
constexpr int MP_THREADS = 704;          // 22 warps
constexpr int MP_ISSUERS = 4;            // MMA-issuing threads per half
constexpr int MP_MAX_STAGES = 8;
constexpr int MP_STAGE_BYTES = MP_NR * 128;          // one K-block of X1: 16 rows x 64 bf16
constexpr int MP_TM_HALF = 128;          // TMEM columns per half: D1 = 4 issuers x 16, D2 = 2 tiles x 2 K-halves x 16
constexpr int MP_TM_D2 = MP_ISSUERS * MP_NR;
constexpr uint32_t MP_TMEM_COLS = 256;
constexpr int MP_NBAR = 2 * MP_MAX_STAGES + 3 + 2 + 2 + 3;   // mbarriers per half

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned* p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(p) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

// per-half resources
struct HalfRes {
    uint8_t* ring;      // NSTG x (16 rows x 128 B) X1 K-blocks (TMA destination, UMMA B operand)
    uint8_t* x2;        // 16 rows x 128 B swizzled B operand of the output GEMM
    float* P;           // 2 x 1024 fp32 prefetched partial gathers of the owned rows
    uint8_t* Q;         // owned rows x 32-entry sample ring
    float* U;           // this step's uniform per owned row
    float* logit;       // owned rows x 256 reduced logits
    uint64_t *full, *empty, *bar_d1, *x2_ready, *bar_d2, *p_ready, *p_free, *q_ready;
    uint32_t tm_d1, tm_d2;
    int row0;           // first owned row (global utterance index)
    int xrow;           // first row of this half in the X1 exchange buffer
    unsigned* ctr;      // group barrier counter of (row group, half)
    float* part;        // (NS, 16, 256) split-K partial logits of (row group, half)
};

#define MP_TRACE(slot)                                                                                 \
    do {                                                                                               \
        if (p.trace && blockIdx.x == 0 && hh == 0 && tidE == 0) p.trace[k * 64 + (slot)] = clock64();  \
    } while (0)

__global__ void __launch_bounds__(MP_THREADS, 1)
k_mlp_persist(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWo,
              const __grid_constant__ CUtensorMap tmX1, const MlpPersistParams p) {
    const int H = p.H, KB = H >> 6, NS = H >> 6, RPC = MP_NR / NS, FS = p.FS;
    const int NSTG = KB < MP_MAX_STAGES ? KB : MP_MAX_STAGES;
    const int nissue = KB < MP_ISSUERS ? KB : MP_ISSUERS;     // MMA-issuing threads in use per half
    const int rg = blockIdx.x / NS, sl = blockIdx.x % NS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sWh = smem;                                  // KB x (64 rows x 128 B)
    uint8_t* sWo = sWh + (size_t)KB * 8192;               // 2 x (128 rows x 128 B)
    uint8_t* sRingAll = sWo + 32768;                      // MP_NH x NSTG x 2 KB
    uint8_t* sX2All = sRingAll + (size_t)MP_NH * NSTG * MP_STAGE_BYTES;
    float* sPall = (float*)(sX2All + MP_NH * MP_STAGE_BYTES);
    float* sLall = sPall + MP_NH * 2 * 1024;
    float* sUall = sLall + (size_t)MP_NH * RPC * SRNN_Q;
    uint8_t* sQall = (uint8_t*)(sUall + MP_NH * MP_NR);
    uint64_t* bars = (uint64_t*)(sQall + MP_NH * MP_NR * 32);   // w_ready + MP_NH * MP_NBAR
    uint64_t* w_ready = bars;
    uint32_t* tmem_slot = (uint32_t*)(bars + 1 + MP_NH * MP_NBAR);

    const int i0 = *p.step_base + p.pos0;                 // absolute index of the first sample of this launch

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmWh);
        prefetch_tmap(&tmWo);
        prefetch_tmap(&tmX1);
        mbar_init(w_ready, 1);
        for (int h2 = 0; h2 < MP_NH; ++h2) {
            uint64_t* b = bars + 1 + h2 * MP_NBAR;
            for (int s = 0; s < 2 * MP_MAX_STAGES; ++s) mbar_init(&b[s], 1);      // full[8], empty[8]
            mbar_init(&b[16], nissue);                                            // bar_d1
            mbar_init(&b[17], 1);                                                 // x2_ready
            mbar_init(&b[18], nissue);                                            // bar_d2
            mbar_init(&b[19], 64);                                                // p_ready[2]: the half's 64 G threads
            mbar_init(&b[20], 64);
            mbar_init(&b[21], 1);                                                 // p_free[2]
            mbar_init(&b[22], 1);
            mbar_init(&b[23], 1);                                                 // q_ready[3]
            mbar_init(&b[24], 1);
            mbar_init(&b[25], 1);
        }
        fence_barrier_init();
    }
    if (warp == 14) tmem_alloc<MP_TMEM_COLS>(tmem_slot);
    // owned rows' sample rings: the FS most recent samples before i0 (earlier launches / the q_zero prefix)
    for (int e = threadIdx.x; e < MP_NH * RPC * 32; e += MP_THREADS) {
        const int h2 = e / (RPC * 32), r = (e / 32) % RPC, w = e & 31;
        const int b = rg * 32 + h2 * MP_NR + sl * RPC + r;
        const int a = i0 - 32 + w;                        // absolute sample index, slot a & 31
        uint8_t q = 128;
        if (b < p.B && a >= 0 && a >= i0 - FS) q = __ldcg(p.seq + (size_t)b * p.Lseq + a);
        sQall[(h2 * MP_NR + r) * 32 + (a & 31)] = q;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // warp-uniform copy (shfl from lane 0) so the MMA operands stay in uniform registers: otherwise the compiler wraps
    // every tcgen05.mma of the single issuing thread in an ELECT/R2UR.BROADCAST loop
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const int stg_mask = NSTG - 1, stg_shift = 31 - __clz(NSTG);   // NSTG is a power of two

    // which half does this warp serve?
    int hh;
    if (warp < 4) hh = warp >> 1;
    else if (warp < 12) hh = (warp - 4) >> 2;
    else if (warp < 14) hh = warp - 12;
    else hh = (warp - 14) >> 2;
    HalfRes R;
    {
        uint64_t* b = bars + 1 + hh * MP_NBAR;
        R.ring = sRingAll + (size_t)hh * NSTG * MP_STAGE_BYTES;
        R.x2 = sX2All + hh * MP_STAGE_BYTES;
        R.P = sPall + hh * 2048;
        R.Q = sQall + hh * MP_NR * 32;
        R.U = sUall + hh * MP_NR;
        R.logit = sLall + (size_t)hh * RPC * SRNN_Q;
        R.full = b; R.empty = b + 8; R.bar_d1 = b + 16; R.x2_ready = b + 17; R.bar_d2 = b + 18;
        R.p_ready = b + 19; R.p_free = b + 21; R.q_ready = b + 23;
        R.tm_d1 = tmem + hh * MP_TM_HALF;
        R.tm_d2 = R.tm_d1 + MP_TM_D2;
        R.xrow = rg * 32 + hh * MP_NR;
        R.row0 = R.xrow + sl * RPC;
        R.ctr = p.ctr + rg * MP_NH + hh;
        R.part = p.part + (size_t)(rg * MP_NH + hh) * NS * MP_NR * SRNN_Q;
    }

    if (warp >= 12 && warp < 14) {
        // ===================== TMA producer of half hh =====================
        if (lane == 0) {
            if (hh == 0) {                                // resident weights, once per launch
                mbar_expect_tx(w_ready, (uint32_t)(KB * 8192 + 32768));
                for (int kb = 0; kb < KB; ++kb) tma_load_2d(sWh + (size_t)kb * 8192, &tmWh, w_ready, kb * 64, sl * 64);
                tma_load_2d(sWo, &tmWo, w_ready, sl * 64, 0);
                tma_load_2d(sWo + 16384, &tmWo, w_ready, sl * 64, 128);
            }
            int it = 0;
            for (int k = 0; k < p.nsteps; ++k) {
                {   // group barrier A, waiting side: arrivals so far = (2k+1) * NS once every slice has published X1
                    const unsigned target = (unsigned)(2 * k + 1) * (unsigned)NS;
                    while (ld_acquire_gpu(R.ctr) < target) {
                    }
                    fence_proxy_async_all();              // other CTAs' generic-proxy global writes -> visible to TMA reads
                }
                if (p.trace && blockIdx.x == 0 && hh == 0) p.trace[k * 64 + 10] = clock64();
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it & stg_mask;
                    const uint32_t ph = (it >> stg_shift) & 1;
                    mbar_wait(&R.empty[s], ph ^ 1);
                    mbar_expect_tx(&R.full[s], MP_STAGE_BYTES);
                    tma_load_2d(R.ring + (size_t)s * MP_STAGE_BYTES, &tmX1, &R.full[s], kb * 64, R.xrow);
                }
                if (p.trace && blockIdx.x == 0 && hh == 0) p.trace[k * 64 + 11] = clock64();
            }
        }
    } else if (warp >= 14) {
        // ===================== MMA issuers of half hh (one thread each) =====================
        const int w = (warp - 14) & 3;
        if (lane == 0 && w < nissue) {
            constexpr uint32_t idesc1 = umma_idesc_bf16(64, MP_NR);
            constexpr uint32_t idesc2 = umma_idesc_bf16(128, MP_NR);
            mbar_wait(w_ready, 0);
            const uint64_t dA0 = umma_desc_sw128(smem_u32(sWh));      // + kb * (8192 >> 4)
            const uint64_t dB0 = umma_desc_sw128(smem_u32(R.ring));   // + stage * (2048 >> 4)
            const uint64_t dWo0 = umma_desc_sw128(smem_u32(sWo));     // + tile * (16384 >> 4)
            const uint64_t db2 = umma_desc_sw128(smem_u32(R.x2));
            const uint32_t d1a = R.tm_d1 + (uint32_t)w * MP_NR;       // this issuer's two interleaved M=64 accumulators
            const uint32_t d1b = d1a + (16u << 16);
            for (int k = 0; k < p.nsteps; ++k) {
                for (int kb = w; kb < KB; kb += MP_ISSUERS) {
                    const int it = k * KB + kb;
                    const int s = it & stg_mask;
                    const uint32_t ph = (it >> stg_shift) & 1;
                    mbar_wait(&R.full[s], ph);
                    tc_fence_after();
                    const uint64_t da = dA0 + (uint64_t)(kb * 512);
                    const uint64_t db = dB0 + (uint64_t)(s * (MP_STAGE_BYTES >> 4));
                    const uint32_t acc = kb >= MP_ISSUERS;
                    umma_bf16(d1a, da, db, idesc1, acc);
                    umma_bf16(d1b, da + 2, db + 2, idesc1, acc);
                    umma_bf16(d1a, da + 4, db + 4, idesc1, 1);
                    umma_bf16(d1b, da + 6, db + 6, idesc1, 1);
                    umma_commit(&R.empty[s]);
                }
                umma_commit(R.bar_d1);
                if (p.trace && blockIdx.x == 0 && hh == 0 && w == 0) p.trace[k * 64 + 14] = clock64();
                mbar_wait(R.x2_ready, k & 1);
                tc_fence_after();
                for (int c = w; c < 4; c += nissue) {                 // (output tile, K half) -> its own accumulator
                    const int t2 = c >> 1, h2 = c & 1;
                    const uint64_t da2 = dWo0 + (uint64_t)(t2 * 1024) + (uint64_t)(h2 * 4);
                    const uint32_t d2 = R.tm_d2 + (uint32_t)c * MP_NR;
                    umma_bf16(d2, da2, db2 + (uint64_t)(h2 * 4), idesc2, 0);
                    umma_bf16(d2, da2 + 2, db2 + (uint64_t)(h2 * 4) + 2, idesc2, 1);
                }
                umma_commit(R.bar_d2);
            }
        }
    } else if (warp >= 4) {
        // ===================== E warps of half hh: per-row work, epilogues, group barriers =====================
        const int tidE = (threadIdx.x - 128) & 127;
        const int q4 = warp & 3;                          // TMEM lane quadrant this warp may access
        const int barid = 1 + hh;
        const int flat = tidE * 8;                        // 8 consecutive features of one owned row
        const int rl = flat / H, f0 = flat % H;
        const int b = R.row0 + rl;
        unsigned bar_no = 0;
        // the owned rows' uniforms are fetched one step ahead (a DRAM miss would otherwise sit on the serial path)
        const int ub = R.row0 + tidE;
        const bool u_mine = tidE < RPC && ub < p.B;
        float u_next = u_mine ? __ldg(p.uniforms + (size_t)(i0 - p.lookback) * p.B + ub) : 0.f;
        for (int k = 0; k < p.nsteps; ++k) {
            const int i = i0 + k;
            // ---- E1: x1 = relu(P + Tbl[FS-2][.] + Tbl[FS-1][newest sample]) for the owned rows -> global X1 ----
            MP_TRACE(0);
            if (tidE < RPC) {
                R.U[tidE] = u_next;
                if (u_mine && k + 1 < p.nsteps) u_next = __ldg(p.uniforms + (size_t)(i + 1 - p.lookback) * p.B + ub);
            }
            mbar_wait(&R.p_ready[k & 1], (k >> 1) & 1);
            MP_TRACE(1);
            {
                const int qn = R.Q[rl * 32 + ((i - 1) & 31)], qm = R.Q[rl * 32 + ((i - 2) & 31)];
                const uint4 t0 = __ldg(reinterpret_cast<const uint4*>(p.tbl + ((size_t)(FS - 1) * SRNN_Q + qn) * H + f0));
                const uint4 t1 = __ldg(reinterpret_cast<const uint4*>(p.tbl + ((size_t)(FS - 2) * SRNN_Q + qm) * H + f0));
                float tv[8], tw[8];
                bf16x8_to_f32(t0, tv);
                bf16x8_to_f32(t1, tw);
                const float4* pp = reinterpret_cast<const float4*>(R.P + (k & 1) * 1024 + flat);
                const float4 p0 = pp[0], p1 = pp[1];
                const float x[8] = {p0.x + tv[0] + tw[0], p0.y + tv[1] + tw[1], p0.z + tv[2] + tw[2], p0.w + tv[3] + tw[3],
                                    p1.x + tv[4] + tw[4], p1.y + tv[5] + tw[5], p1.z + tv[6] + tw[6], p1.w + tv[7] + tw[7]};
                uint32_t o[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const __nv_bfloat162 t = __floats2bfloat162_rn(fmaxf(x[2 * v], 0.f), fmaxf(x[2 * v + 1], 0.f));
                    o[v] = *reinterpret_cast<const uint32_t*>(&t);
                }
                // x1 has RG*32 rows: padded rows are always in range
                *reinterpret_cast<uint4*>(p.x1 + (size_t)b * H + f0) = make_uint4(o[0], o[1], o[2], o[3]);
            }
            // ---- group barrier A: E warps only ARRIVE (release); the TMA thread of this half waits for the other slices,
            // so the wake-up hand-off between warps is off the serial path.  X1 is not rewritten before barrier B. ----
            MP_TRACE(2);
            named_bar_sync(barid, 128);
            ++bar_no;
            if (tidE == 0) {
                red_release_gpu(R.ctr);
                mbar_arrive(&R.p_free[k & 1]);            // P[k&1] consumed by every E thread
            }
            MP_TRACE(3);
            // ---- epilogue 1: D1 (+bias, ReLU) -> bf16 swizzled B operand in smem ----
            mbar_wait(R.bar_d1, k & 1);
            MP_TRACE(4);
            tc_fence_after();
            {
                float v[16];
                tmem_ld16(R.tm_d1 + ((uint32_t)(32 * q4) << 16), v);
#pragma unroll 1
                for (int c = 1; c < nissue; ++c) {        // the other issuers' accumulators
                    float w2[16];
                    tmem_ld16(R.tm_d1 + ((uint32_t)(32 * q4) << 16) + c * MP_NR, w2);
#pragma unroll
                    for (int n = 0; n < 16; ++n) v[n] += w2[n];
                }
#pragma unroll
                for (int n = 0; n < 16; ++n)              // lanes 16-31 hold the second accumulator of the same features
                    v[n] += __shfl_down_sync(0xffffffffu, v[n], 16);
                if (lane < 16) {
                    const int f = 16 * q4 + lane;         // feature inside the slice = K index of the output GEMM
                    const float bv = p.b_hid[sl * 64 + f];
                    uint8_t* base = R.x2 + (f & 7) * 2;
                    const int chunk = f >> 3;
#pragma unroll
                    for (int n = 0; n < 16; ++n)
                        *reinterpret_cast<__nv_bfloat16*>(base + n * 128 + ((chunk ^ (n & 7)) << 4)) =
                            __float2bfloat16(fmaxf(v[n] + bv, 0.f));
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();                     // generic smem writes -> visible to the UMMA operand reads
            named_bar_sync(barid, 128);
            if (tidE == 0) mbar_arrive(R.x2_ready);
            MP_TRACE(5);
            // ---- epilogue 2: split-K partial logits -> global Part[sl][row][256] ----
            mbar_wait(R.bar_d2, k & 1);
            MP_TRACE(6);
            tc_fence_after();
            {
                float* dst = R.part + ((size_t)sl * MP_NR) * SRNN_Q + 32 * q4 + lane;
#pragma unroll
                for (int t2 = 0; t2 < 2; ++t2) {
                    float v0[16], v1[16];
                    tmem_ld16(R.tm_d2 + (2 * t2) * MP_NR + ((uint32_t)(32 * q4) << 16), v0);
                    tmem_ld16(R.tm_d2 + (2 * t2 + 1) * MP_NR + ((uint32_t)(32 * q4) << 16), v1);
#pragma unroll
                    for (int n = 0; n < 16; ++n) dst[(size_t)n * SRNN_Q + t2 * 128] = v0[n] + v1[n];
                }
            }
            tc_fence_before();
            // ---- group barrier B: all slices' partial logits are in global memory ----
            MP_TRACE(7);
            named_bar_sync(barid, 128);
            ++bar_no;
            if (tidE == 0) {
                red_release_gpu(R.ctr);
                const unsigned target = bar_no * (unsigned)NS;
                while (ld_acquire_gpu(R.ctr) < target) {
                }
            }
            named_bar_sync(barid, 128);
            MP_TRACE(8);
            // ---- reduce the NS split-K partials of the owned rows (every load in flight) ----
            for (int e = tidE; e < RPC * (SRNN_Q / 4); e += 128) {
                const int r2 = e / (SRNN_Q / 4), o4 = (e % (SRNN_Q / 4)) * 4;
                const int n = sl * RPC + r2;
                const float* src = R.part + (size_t)n * SRNN_Q + o4;
                float4 acc = __ldg(reinterpret_cast<const float4*>(p.b_out + o4));
                for (int s0 = 0; s0 < NS; s0 += 8) {
                    float4 a[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int s2 = (s0 + u < NS) ? s0 + u : NS - 1;
                        a[u] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)s2 * MP_NR * SRNN_Q));
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)           // fixed summation order: slice 0, 1, 2, ...
                        if (s0 + u < NS) {
                            acc.x += a[u].x; acc.y += a[u].y; acc.z += a[u].z; acc.w += a[u].w;
                        }
                }
                *reinterpret_cast<float4*>(R.logit + r2 * SRNN_Q + o4) = acc;
            }
            named_bar_sync(barid, 128);
            // ---- log-softmax + defined sampler, one warp per owned row ----
            for (int r2 = q4; r2 < RPC; r2 += 4) {
                const int bb = R.row0 + r2;
                float v[8];
                {
                    const float4* lp = reinterpret_cast<const float4*>(R.logit + r2 * SRNN_Q + lane * 8);
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
                    const int idx = sampler_warp(v, R.U[r2], lane);
                    if (lane == 0) {
                        p.seq[(size_t)bb * p.Lseq + i] = (uint8_t)idx;
                        R.Q[r2 * 32 + (i & 31)] = (uint8_t)idx;
                    }
                } else if (lane == 0) {
                    R.Q[r2 * 32 + (i & 31)] = 128;
                }
            }
            named_bar_sync(barid, 128);                   // sample i of every owned row is in Q
            if (tidE == 0) mbar_arrive(&R.q_ready[k % 3]);
            MP_TRACE(9);
        }
    } else {
        // ===================== G warps of half hh: prefetch P_g = c0 + taps 0..FS-3 for step g (two steps of slack) =====
        const int tidG = threadIdx.x & 63;
        const int flat = tidG * 16;
        const int rl = flat / H, f0 = flat % H;
        const int b = R.row0 + rl;
        const int bc = b < p.B ? b : p.B - 1;             // clamp: padded rows compute garbage that is never used
        for (int g = 0; g < p.nsteps; ++g) {
            const int i = i0 + g;
            if (g >= 3)   // sample i-3 (tap FS-3, the newest one this prefetch uses) has been drawn.  Three barriers by
                          // step % 3: E can be up to two steps past the awaited one, which would alias a phase parity.
                mbar_wait(&R.q_ready[g % 3], (g / 3 - 1) & 1);
            if (g >= 2) mbar_wait(&R.p_free[g & 1], ((g >> 1) - 1) & 1);
            float acc[16];
            {
                const float4* cp = reinterpret_cast<const float4*>(p.c0 + (size_t)bc * FS * H + (size_t)(i % FS) * H + f0);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float4 c = __ldg(cp + v);
                    acc[4 * v] = c.x; acc[4 * v + 1] = c.y; acc[4 * v + 2] = c.z; acc[4 * v + 3] = c.w;
                }
            }
            const uint8_t* qrow = R.Q + rl * 32;
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
            float4* pp = reinterpret_cast<float4*>(R.P + (g & 1) * 1024 + flat);
#pragma unroll
            for (int v = 0; v < 4; ++v) pp[v] = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
            mbar_arrive(&R.p_ready[g & 1]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 14) tmem_dealloc<MP_TMEM_COLS>(tmem);
}

size_t mlp_persist_smem(int H) {
    const int KB = H / 64, NSTG = KB < MP_MAX_STAGES ? KB : MP_MAX_STAGES;
    const int RPC = MP_NR / (H / 64);
    return (size_t)KB * 8192 + 32768 + (size_t)MP_NH * NSTG * MP_STAGE_BYTES + MP_NH * MP_STAGE_BYTES +
           MP_NH * 2 * 1024 * sizeof(float) + (size_t)MP_NH * RPC * SRNN_Q * sizeof(float) + MP_NH * MP_NR * sizeof(float) +
           MP_NH * MP_NR * 32 + (1 + MP_NH * MP_NBAR) * sizeof(uint64_t) + 64 + 1024 /*alignment slack*/;
}

bool mlp_persist_supported(int H, int FS, int B, int n_sms) {
    if (H % 64 || H > 1024) return false;
    const int NS = H / 64;
    if (MP_NR % NS) return false;
    if (FS < 2 || FS > 31) return false;
    const int RG = (B + 31) / 32;
    return RG * NS <= n_sms;
}

// x1 must hold RG*32 rows; part RG*NS*32*256 floats; ctr 2*RG counters (zeroed here).
int mlp_persist_launch(const __nv_bfloat16* w_hid16, const __nv_bfloat16* w_out16, const MlpPersistParams& p,
                       cudaStream_t st) {
    const int H = p.H, NS = H / 64, RG = (p.B + 31) / 32;
    CUtensorMap tmWh, tmWo, tmX1;
    SRNN_TRY(make_tmap_bf16(&tmWh, w_hid16, H, H, H, 64));
    SRNN_TRY(make_tmap_bf16(&tmWo, w_out16, SRNN_Q, H, H, 128));
    SRNN_TRY(make_tmap_bf16(&tmX1, p.x1, (uint64_t)RG * 32, H, H, MP_NR));
    const size_t smem = mlp_persist_smem(H);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        SRNN_CUDA(cudaFuncSetAttribute(k_mlp_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    SRNN_CUDA(cudaMemsetAsync(p.ctr, 0, sizeof(unsigned) * RG * MP_NH, st));
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
