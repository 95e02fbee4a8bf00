// Cluster form of the persistent fused sample-level kernel (SRNN_MODE_BF16 generation, H = 1024): the same per-sample chain as
// mlp_persist.cu --  table gather -> ReLU -> hidden GEMM -> ReLU -> output GEMM -> log-softmax -> inverse-CDF sample  (replaces
// model.py:504-517 executed once per sample by the reference) -- but a row group is ONE 8-CTA thread-block cluster instead of 16
// CTAs that meet at two release/acquire counters in L2:
//
//   * CTA c of a cluster owns hidden features [128c, 128c+128) and RPC (2 or 3) of the group's R = 8*RPC utterance rows.
//   * Its 128 x 1024 slice of W_hid (256 KB in bf16, more than shared memory holds) is RESIDENT ACROSS ALL SAMPLES of a launch in
//     TENSOR MEMORY + shared memory: k-blocks 0..11 live in TMEM columns 0..383 (lane = feature, one 32-bit column = two K
//     elements) and feed tcgen05.mma as the A operand straight from TMEM (the ".ts" form: [d_tmem], [a_tmem], b_desc); k-blocks
//     12..15 and the 256 x 128 slice of W_out (split-K output GEMM) stay in shared memory.  Accumulators use columns 384..511.
//   * x1 all-gather: owner rows go to a global exchange buffer; the eight CTAs signal each other with ONE
//     mbarrier.arrive.release.cluster on the peer's shared-memory barrier (no global counter, no ld.acquire spin), then each CTA
//     pulls the R x 1024 tile with TMA (4 boxes, 128B swizzle) as the UMMA B operand (N = 16 or 32 columns).
//   * split-K partial logits: staged in shared memory ([row][256] fp32, in the then idle x1 tile) and pushed to the owner of each
//     row by ONE cp.async.bulk shared::cta -> shared::cluster copy per peer, completing (complete_tx) on the owner's mbarrier: no
//     global partials, no second barrier, no L2 read on the way to the softmax.
//   Every exchange involves all eight CTAs, so the exchanges themselves are the only synchronisation: buffers are single and a
//   fast CTA can never run more than one exchange ahead of a slow one.
// Clusters are independent (different utterances): no cooperative launch, any number of clusters, CUDA-graph capturable.
// Measured primitives behind the design (tools/ts_probe.cu, tools/dsmem_probe.cu; profiles/README.md): TS MMA 64 x (M128 N16 K16)
// from four issuing warps 2302 cycles (A from shared memory: 3609); 8-CTA push of 16 KB partial logits 716 cycles; only 15
// 8-CTA clusters with > 113 KB of shared memory are co-resident on a B200 (GPC floor-sweeping), hence RPC = 3 for 256 utterances.
#include "common.cuh"
#include "sampler.cuh"
#include "umma.cuh"

namespace srnn {

using namespace ptx;

constexpr int MC_THREADS = 384;            // 12 warps = 3 per SM sub-partition: 168 registers per thread (13 warps: 128)
constexpr int MC_CS = 8;                 // CTAs per cluster = feature slices of 128
constexpr int MC_H = 1024;
constexpr int MC_KB = 16;                // k-blocks of 64
constexpr int MC_KB_TMEM = 12;           // k-blocks of the W_hid slice kept in tensor memory (32 columns each)
constexpr uint32_t MC_COL_D = 384;       // accumulator columns
constexpr uint32_t MC_TMEM_COLS = 512;

__device__ __forceinline__ void mc_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ uint32_t mc_mapa(uint32_t a, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void mc_arrive_remote_release(uint32_t rbar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(rbar) : "memory");
}
__device__ __forceinline__ void mc_wait_acq_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mc_bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t rbar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_cluster),
                 "r"(src_cta), "r"(bytes), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mc_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mc_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void mc_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void mc_fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

__device__ __forceinline__ void mc_bf16x8_to_f32(const uint4& u, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

// exact warp-wide float maximum through the order-preserving float -> int map and one redux.sync
__device__ __forceinline__ float mc_warp_max(float m) {
    int k = __float_as_int(m);
    k ^= (k >> 31) & 0x7fffffff;
    k = __reduce_max_sync(0xffffffffu, k);
    k ^= (k >> 31) & 0x7fffffff;
    return __int_as_float(k);
}

template <int RPC>
struct McLayout {
    static constexpr int R = 8 * RPC;                          // rows per cluster
    static constexpr int NT = RPC == 2 ? 16 : 32;              // UMMA N (RPC = 3: 24 real rows, the fourth 8-row atom is overrun)
    static constexpr uint32_t KB_BYTES = R * 128;              // one k-block of the x1 / x2 tiles
    static constexpr uint32_t OFF_WHT = 0;                                  // 4 k-blocks x (128 rows x 128 B)
    static constexpr uint32_t OFF_WO = 65536;                               // [tile 2][k-block 2] x (128 rows x 128 B)
    static constexpr uint32_t OFF_X1 = 131072;                              // 16 k-blocks x (R rows x 128 B); then staging [R][256] fp32
    static constexpr uint32_t OFF_X2 = OFF_X1 + MC_KB * KB_BYTES;           // 2 k-blocks x (R rows x 128 B) (+ overrun room)
    static constexpr uint32_t X2_BYTES = RPC == 2 ? 4096 : 8192;
    static constexpr uint32_t OFF_LAND = OFF_X2 + X2_BYTES;                 // [8 src][RPC rows][256] fp32
    static constexpr uint32_t OFF_P = OFF_LAND + MC_CS * RPC * 1024;        // [RPC][1024] fp32
    static constexpr uint32_t OFF_LOGIT = OFF_P + RPC * 4096;               // [RPC][256] fp32
    static constexpr uint32_t OFF_Q = OFF_LOGIT + RPC * 1024;               // [RPC][32] sample ring
    static constexpr uint32_t OFF_U = OFF_Q + 128;
    static constexpr uint32_t OFF_BAR = OFF_U + 64;
    static constexpr uint32_t BYTES = OFF_BAR + 256 + 1024;                 // + alignment slack
    // prologue staging slots of 16 KB for the W_hid tiles on their way to tensor memory: the x1 tile, the x2 tile and the
    // landing zone are contiguous and idle until the first step
    static constexpr int NSLOT = (OFF_P - OFF_X1) / 16384;
};

#define MC_TRACE(slot)                                                                      \
    do {                                                                                    \
        if (p.trace && cl == 0 && tidE == 0) p.trace[(c * p.nsteps + k) * 64 + (slot)] = clock64();  \
    } while (0)

template <int RPC>
__global__ void __launch_bounds__(MC_THREADS, 1)
k_mlp_cluster(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWo,
              const __grid_constant__ CUtensorMap tmX1, const __nv_bfloat16* __restrict__ w_hid16, const MlpPersistParams p) {
    using Lay = McLayout<RPC>;
    constexpr int R = Lay::R, NT = Lay::NT, H = MC_H;
    const int FS = p.FS;
    const int cl = blockIdx.x / MC_CS;
    const int c = (int)cluster_ctarank();                 // feature slice = rank in the cluster
    // warp index as a warp-uniform value (shfl from lane 0): the MMA issuers' descriptors derive from it, and operands the
    // compiler cannot prove uniform make it wrap EVERY tcgen05.mma in an ELECT / R2UR.BROADCAST loop (measured: +50 % on the
    // hidden GEMM)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sWhT = smem + Lay::OFF_WHT;
    uint8_t* sWo = smem + Lay::OFF_WO;
    uint8_t* sX1 = smem + Lay::OFF_X1;
    uint8_t* sX2 = smem + Lay::OFF_X2;
    float* sLand = (float*)(smem + Lay::OFF_LAND);
    float* sP = (float*)(smem + Lay::OFF_P);
    float* sLogit = (float*)(smem + Lay::OFF_LOGIT);
    uint8_t* sQ = smem + Lay::OFF_Q;
    float* sU = (float*)(smem + Lay::OFF_U);
    uint64_t* bars = (uint64_t*)(smem + Lay::OFF_BAR);
    uint64_t* w_ready = bars + 0;
    uint64_t* x1_flag = bars + 1;                         // 8 remote arrivals per step: every CTA's x1 rows are in global memory
    uint64_t* full = bars + 2;                            // [4] x1 tile k-block groups landed (TMA)
    uint64_t* bar_d1 = bars + 6;
    uint64_t* x2_ready = bars + 7;
    uint64_t* bar_d2 = bars + 8;
    uint64_t* land_full = bars + 9;                       // partial logits of the owned rows from all 8 CTAs (complete_tx)
    uint64_t* p_ready = bars + 10;
    uint64_t* p_free = bars + 11;
    uint64_t* q_ready = bars + 12;                        // [3]
    uint64_t* wl_full = bars + 16;                        // [NSLOT <= 5] prologue: a W_hid k-block tile landed in the staging slot
    uint64_t* wl_free = bars + 21;                        // [NSLOT <= 5] ... and has been moved to tensor memory
    uint32_t* tmem_slot = (uint32_t*)(bars + 31);

    if (p.trace && cl == 0 && threadIdx.x == 0) p.trace[(c * p.nsteps) * 64 + 63] = clock64();

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmWh);
        prefetch_tmap(&tmWo);
        prefetch_tmap(&tmX1);
        mbar_init(w_ready, 1);
        mbar_init(x1_flag, MC_CS);
        for (int g = 0; g < 4; ++g) mbar_init(&full[g], 1);
        mbar_init(bar_d1, 4);
        mbar_init(x2_ready, 1);
        mbar_init(bar_d2, 4);
        mbar_init(land_full, 1);
        mbar_init(p_ready, 128);
        mbar_init(p_free, 1);
        mbar_init(&q_ready[0], 1);
        mbar_init(&q_ready[1], 1);
        mbar_init(&q_ready[2], 1);
        for (int j = 0; j < Lay::NSLOT; ++j) {
            mbar_init(&wl_full[j], 1);
            mbar_init(&wl_free[j], 128);
        }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc<MC_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // The weights are constants of the generation call: their first staging tiles and the shared-memory part are requested
    // BEFORE griddepcontrol.wait, so that as a programmatic dependent of the upsampling kernel this launch's latency, barrier /
    // TMEM set-up and first weight loads overlap that kernel's drain (a no-op in a plain launch).
    constexpr int KB_EARLY = Lay::NSLOT < MC_KB_TMEM ? Lay::NSLOT : MC_KB_TMEM;
    if (warp == 8 && elect_one()) {
        for (int kb = 0; kb < KB_EARLY; ++kb) {           // round 0 of every staging slot: nothing to wait for
            mbar_expect_tx(&wl_full[kb], 16384);
            tma_load_2d(sX1 + (size_t)kb * 16384, &tmWh, &wl_full[kb], kb * 64, c * 128);
        }
    }
    pdl_wait();                                           // seq, c0, step_base below come from the preceding launches
    const int i0 = *p.step_base + p.pos0;                 // absolute index of the first sample of this launch
    const int row0 = cl * R + c * RPC;                    // first owned row (global utterance index)
    // owned rows' sample ring: the FS most recent samples before i0 (written by earlier launches / the q_zero prefix)
    for (int e = threadIdx.x; e < RPC * 32; e += MC_THREADS) {
        const int rl = e >> 5, w = e & 31;
        const int b = row0 + rl;
        const int a = i0 - 32 + w;                        // absolute sample index, slot a & 31
        uint8_t q = 128;
        if (b < p.B && a >= 0 && a >= i0 - FS) q = __ldcg(p.seq + (size_t)b * p.Lseq + a);
        sQ[rl * 32 + (a & 31)] = q;
    }
    __syncthreads();
    if (p.trace && cl == 0 && threadIdx.x == 0) p.trace[(c * p.nsteps) * 64 + 20] = clock64();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform (keeps the MMA operands in uniform registers)
    const uint32_t tm_d = tmem + MC_COL_D;

    // G warps: P_g = c0 + taps 0..FS-3 for step g (two steps of slack); thread t owns features 8t..8t+7 of every owned row
    auto g_step = [&](int g) {
        const int f0 = threadIdx.x * 8;
        const int i = i0 + g;
        const bool carry_out = g >= p.nsteps;             // table part of the NEXT launch's first / second sample -> p.pcarry
        const bool carry_in = g <= 1 && p.pcarry != nullptr;
        float* const pcar = p.pcarry + (size_t)(carry_out ? g - p.nsteps : g) * p.carry_plane;
        if (g >= 3)   // sample i-3 (tap FS-3, the newest one this prefetch uses) has been drawn.  Three barriers by
                      // step % 3: E can be up to two steps past the awaited one, which would alias a phase parity.
            mbar_wait(&q_ready[g % 3], (g / 3 - 1) & 1);
        // The gather runs as far ahead as its inputs allow (up to two steps: sample i-3 is its newest tap).  Earlier versions held
        // it back until the current step's x1 tile had landed, so that the latency-bound gather (110 KB per step) never ran beside
        // the TMA loads the hidden GEMM waits for (+1300 cycles per step then); once the E path had shrunk by ~3 k cycles that rule
        // made the GATHER the critical path (E waited 1.4 k cycles per step for P): without it 1824 -> 1874x real-time.
        // SRNN_MC_DBG=2 restores the old rule.
        if (g >= 1 && g < p.nsteps && (p.dbg & 2)) mbar_wait(&full[3], (g - 1) & 1);
        float acc[RPC][8];
#pragma unroll
        for (int r = 0; r < RPC; ++r) {
            const int b = row0 + r;
            const int bc = b < p.B ? b : p.B - 1;         // clamp: padded rows compute garbage that is never used
            if (carry_out) {
#pragma unroll
                for (int v = 0; v < 8; ++v) acc[r][v] = 0.f;
                continue;
            }
            const float4* cp = reinterpret_cast<const float4*>(p.c0 + (size_t)bc * FS * H + (size_t)(i % FS) * H + f0);
            // c0 is rewritten between launches by the upsampling kernel: read at L2 (like seq), not through the non-coherent path
            const float4 ca = __ldcg(cp), cb = __ldcg(cp + 1);
            acc[r][0] = ca.x; acc[r][1] = ca.y; acc[r][2] = ca.z; acc[r][3] = ca.w;
            acc[r][4] = cb.x; acc[r][5] = cb.y; acc[r][6] = cb.z; acc[r][7] = cb.w;
            if (carry_in) {                               // the previous launch already summed this sample's table rows
                const float4* pc = reinterpret_cast<const float4*>(pcar + (size_t)(row0 + r) * H + f0);
                const float4 pa = __ldcg(pc), pb = __ldcg(pc + 1);
                acc[r][0] += pa.x; acc[r][1] += pa.y; acc[r][2] += pa.z; acc[r][3] += pa.w;
                acc[r][4] += pb.x; acc[r][5] += pb.y; acc[r][6] += pb.z; acc[r][7] += pb.w;
            }
        }
        // latency-bound L2 gather: six taps x RPC rows (18 loads of 16 bytes per thread) in flight at a time (nine: slower, the
        // outstanding-load limit of the SM serialises them)
        constexpr int GB = 6;
        for (int j0 = 0; j0 < FS - 2 && !carry_in; j0 += GB) {
            uint4 tv[GB][RPC];
#pragma unroll
            for (int jj = 0; jj < GB; ++jj) {
                const int j = j0 + jj < FS - 2 ? j0 + jj : FS - 3;      // tail: re-load the last tap, not accumulated
#pragma unroll
                for (int r = 0; r < RPC; ++r) {
                    const int qj = sQ[r * 32 + ((i - FS + j) & 31)];
                    tv[jj][r] = __ldg(reinterpret_cast<const uint4*>(p.tbl + ((size_t)j * SRNN_Q + qj) * H + f0));
                }
            }
#pragma unroll
            for (int jj = 0; jj < GB; ++jj) {
                if (j0 + jj < FS - 2) {
#pragma unroll
                    for (int r = 0; r < RPC; ++r) {
                        float f[8];
                        mc_bf16x8_to_f32(tv[jj][r], f);
#pragma unroll
                        for (int v = 0; v < 8; ++v) acc[r][v] += f[v];
                    }
                }
            }
        }
        if (carry_out) {
#pragma unroll
            for (int r = 0; r < RPC; ++r) {
                float4* pc = reinterpret_cast<float4*>(pcar + (size_t)(row0 + r) * H + f0);
                pc[0] = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                pc[1] = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
            }
            return;
        }
        if (g >= 1) mbar_wait(p_free, (g - 1) & 1);       // E has consumed P of the previous step (single buffer)
#pragma unroll
        for (int r = 0; r < RPC; ++r) {
            float4* pp = reinterpret_cast<float4*>(sP + r * H + f0);
            pp[0] = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            pp[1] = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
        }
        mbar_arrive(p_ready);
    };

    float bo[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // E warps: output bias of the logits this lane reduces (lane * 8 ..)
    float u_first = 0.f;                                    // E threads < RPC: uniform of the first sample
    // ---- prologue, overlapped: (E warps + TMA) W_hid slice k-blocks 0..11 -> tensor memory; (G warps) P of the first sample ----
    // Each thread owns one TMEM lane = one feature row, so reading the weights straight from global memory touches 32 different
    // 128-byte lines per warp instruction (measured: 20 k cycles per launch).  Instead TMA streams 128-row x 128-byte swizzled
    // tiles through NSLOT staging slots in the (still idle) x1 / x2 / landing area and every thread picks its row out of shared memory.
    if (warp == 8) {
        if (elect_one()) {
            for (int kb = KB_EARLY; kb < MC_KB_TMEM; ++kb) {     // first what the E warps are waiting for
                const int slot = kb % Lay::NSLOT, round = kb / Lay::NSLOT;
                mbar_wait(&wl_free[slot], (round - 1) & 1);
                mbar_expect_tx(&wl_full[slot], 16384);
                tma_load_2d(sX1 + (size_t)slot * 16384, &tmWh, &wl_full[slot], kb * 64, c * 128);
            }
            // the shared-memory part (W_hid k-blocks 12..15, then the W_out slice) is first needed by the hidden GEMM of step 0:
            // it streams in under that step's gather / exchange
            mbar_expect_tx(w_ready, 131072);
            for (int j = 0; j < MC_KB - MC_KB_TMEM; ++j)
                tma_load_2d(sWhT + (size_t)j * 16384, &tmWh, w_ready, (MC_KB_TMEM + j) * 64, c * 128);
            for (int t2 = 0; t2 < 2; ++t2)
                for (int kb2 = 0; kb2 < 2; ++kb2)
                    tma_load_2d(sWo + (size_t)(t2 * 2 + kb2) * 16384, &tmWo, w_ready, c * 128 + kb2 * 64, t2 * 128);
        }
    } else if (warp >= 4 && warp < 8) {
        // per-launch constants of the E warps, requested before the staging loop so that their (DRAM) latency is hidden
        {
            const int tidE0 = threadIdx.x - 128, ub0 = row0 + tidE0;
            if (tidE0 < RPC && ub0 < p.B) u_first = __ldg(p.uniforms + (size_t)(i0 - p.lookback) * p.u_ld + ub0);
#pragma unroll
            for (int j = 0; j < 8; ++j) bo[j] = __ldg(p.b_out + lane * 8 + j);
        }
        const int q4 = warp & 3, row = q4 * 32 + lane;
        const uint32_t tbase = tmem + ((uint32_t)(q4 * 32) << 16);
#pragma unroll 1
        for (int kb = 0; kb < MC_KB_TMEM; ++kb) {
            const int slot = kb % Lay::NSLOT, round = kb / Lay::NSLOT;
            mbar_wait(&wl_full[slot], round & 1);
            const uint8_t* tile = sX1 + (size_t)slot * 16384 + row * 128;
            uint32_t r0[16], r1[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {                 // 16-byte chunk j of the row sits at chunk position j ^ (row & 7)
                const uint4 a = *reinterpret_cast<const uint4*>(tile + ((j ^ (row & 7)) << 4));
                const uint4 b2 = *reinterpret_cast<const uint4*>(tile + (((4 + j) ^ (row & 7)) << 4));
                r0[4 * j] = a.x; r0[4 * j + 1] = a.y; r0[4 * j + 2] = a.z; r0[4 * j + 3] = a.w;
                r1[4 * j] = b2.x; r1[4 * j + 1] = b2.y; r1[4 * j + 2] = b2.z; r1[4 * j + 3] = b2.w;
            }
            mc_tmem_st16(tbase + kb * 32, r0);
            mc_tmem_st16(tbase + kb * 32 + 16, r1);
            // Release the slot only AFTER the tensor-memory stores have consumed the loaded registers: an mbarrier arrive issued
            // right behind the LDS instructions can overtake them while they queue behind the G warps' global loads in the LSU,
            // and the next TMA tile then lands in the slot before it has been read (measured: corrupted columns in launches >= 2).
            mbar_arrive(&wl_free[slot]);
        }
        mc_tmem_st_wait();
        if (p.trace && cl == 0 && threadIdx.x == 128) p.trace[(c * p.nsteps) * 64 + 21] = clock64();
    } else if (warp < 4) {
        g_step(0);                                        // P of the first sample while the weights stream in
        if (p.trace && cl == 0 && threadIdx.x == 0) p.trace[(c * p.nsteps) * 64 + 22] = clock64();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                   // every CTA's barriers are initialised before any remote arrive / copy
    tc_fence_after();
    if (p.trace && cl == 0 && threadIdx.x == 0) p.trace[(c * p.nsteps) * 64 + 23] = clock64();

    if (warp >= 8) {
        // ===================== MMA issuers (one thread each; k-block kb belongs to issuer kb % 4) =====================
        const int w = warp - 8;
        if (elect_one()) {                                // elect.sync: the compiler then issues each tcgen05.mma once, no ELECT loop
            constexpr uint32_t idesc = umma_idesc_bf16(128, NT);
            mbar_wait(w_ready, 0);
            const uint64_t dWhT = umma_desc_sw128(smem_u32(sWhT));    // + (kb - 12) * (16384 >> 4)
            const uint64_t dX1 = umma_desc_sw128(smem_u32(sX1));      // + kb * (KB_BYTES >> 4)
            const uint64_t dWo = umma_desc_sw128(smem_u32(sWo)) + (uint64_t)(w * 1024);   // (tile w >> 1, k-block w & 1)
            const uint64_t dX2 = umma_desc_sw128(smem_u32(sX2)) + (uint64_t)((w & 1) * (Lay::KB_BYTES >> 4));
            const uint32_t d = tm_d + (uint32_t)w * NT;               // this issuer's private accumulator (both GEMMs)
            for (int k = 0; k < p.nsteps; ++k) {
                if (w == 0) {                             // issuer 0 is also the TMA producer of the x1 tile
                    mc_wait_acq_cluster(x1_flag, k & 1);  // all 8 CTAs have published their x1 rows of this step
                    mc_fence_proxy_async_all();           // generic-proxy global writes (acquired above) -> visible to TMA reads
                    if (p.trace && cl == 0) p.trace[(c * p.nsteps + k) * 64 + 10] = clock64();
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        mbar_expect_tx(&full[g], 4 * Lay::KB_BYTES);
                        tma_load_3d(sX1 + (size_t)g * 4 * Lay::KB_BYTES, &tmX1, &full[g], 0, cl * R, g * 4);
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int kb = w + 4 * g;
                    mbar_wait(&full[g], k & 1);
                    tc_fence_after();
                    const uint64_t db = dX1 + (uint64_t)(kb * (Lay::KB_BYTES >> 4));
                    if (kb < MC_KB_TMEM) {
                        const uint32_t ta = tmem + (uint32_t)(kb * 32);
#pragma unroll
                        for (int j = 0; j < 4; ++j) mc_umma_ts(d, ta + 8 * j, db + 2 * j, idesc, (g | j) != 0);
                    } else {
                        const uint64_t da = dWhT + (uint64_t)((kb - MC_KB_TMEM) * 1024);
#pragma unroll
                        for (int j = 0; j < 4; ++j) umma_bf16(d, da + 2 * j, db + 2 * j, idesc, 1);
                    }
                }
                umma_commit(bar_d1);
                if (p.trace && cl == 0 && w == 0) p.trace[(c * p.nsteps + k) * 64 + 14] = clock64();
                mbar_wait(x2_ready, k & 1);
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < 4; ++j) umma_bf16(d, dWo + 2 * j, dX2 + 2 * j, idesc, j != 0);
                umma_commit(bar_d2);
                if (p.trace && cl == 0 && w == 0) p.trace[(c * p.nsteps + k) * 64 + 15] = clock64();
            }
        }
    } else if (warp >= 4) {
        // ===================== E warps: per-row work, epilogues, exchanges =====================
        const int tidE = threadIdx.x - 128;
        const int q4 = warp & 3;                          // TMEM lane quadrant this warp may access
        const int f0 = tidE * 8;                          // 8 consecutive features of EVERY owned row
        const uint32_t lane_base = (uint32_t)(32 * q4) << 16;
        const int ub = row0 + tidE;
        const bool u_mine = tidE < RPC && ub < p.B;
        float u_next = u_first;
        const uint32_t l_flag = smem_u32(x1_flag), l_land = smem_u32(sLand), l_landbar = smem_u32(land_full), l_stg = smem_u32(sX1);
        if (p.dbg & 1) {                                  // self-check (SRNN_MC_DBG=1): read the TMEM-resident weights back, compare with global memory
            const int row = q4 * 32 + lane;
            int bad = 0, first = -1;
            for (int kb = 0; kb < MC_KB_TMEM; ++kb) {
                float v0[16], v1[16];
                tmem_ld16(tmem + lane_base + kb * 32, v0);
                tmem_ld16(tmem + lane_base + kb * 32 + 16, v1);
                const uint32_t* g = reinterpret_cast<const uint32_t*>(w_hid16 + (size_t)(c * 128 + row) * H + kb * 64);
                for (int j = 0; j < 16; ++j) {
                    if (__float_as_uint(v0[j]) != g[j]) { ++bad; if (first < 0) first = kb * 32 + j; }
                    if (__float_as_uint(v1[j]) != g[16 + j]) { ++bad; if (first < 0) first = kb * 32 + 16 + j; }
                }
            }
            if (bad) printf("TMEMBAD cl %d c %d row %d: %d words differ, first column %d (i0 %d)\n", cl, c, row, bad, first, i0);
            tc_fence_before();
            mc_bar_sync(1, 128);
            tc_fence_after();
        }
        for (int k = 0; k < p.nsteps; ++k) {
            const int i = i0 + k;
            // the next kernel on the stream (the tier step that consumes this frame) may be scheduled now: it waits in
            // griddepcontrol.wait until this grid has completed, only its launch latency and prologue move forward
            if (k == p.nsteps - 1) pdl_trigger();     // (one step earlier: no difference)
            // ---- E1: x1 = relu(P + Tbl[FS-1][newest] + Tbl[FS-2][second newest]) for the owned rows -> global exchange buffer ----
            MC_TRACE(0);
            if (tidE < RPC) {
                sU[tidE] = u_next;
                if (u_mine && k + 1 < p.nsteps) u_next = __ldg(p.uniforms + (size_t)(i + 1 - p.lookback) * p.u_ld + ub);
            }
            if (tidE == 0) mbar_expect_tx(land_full, MC_CS * RPC * 1024);   // this step's partial logits (cannot arrive earlier)
            uint4 t0[RPC], t1[RPC];
#pragma unroll
            for (int r = 0; r < RPC; ++r) {
                const int qn = sQ[r * 32 + ((i - 1) & 31)], qm = sQ[r * 32 + ((i - 2) & 31)];
                t0[r] = __ldg(reinterpret_cast<const uint4*>(p.tbl + ((size_t)(FS - 1) * SRNN_Q + qn) * H + f0));
                t1[r] = __ldg(reinterpret_cast<const uint4*>(p.tbl + ((size_t)(FS - 2) * SRNN_Q + qm) * H + f0));
            }
            mbar_wait(p_ready, k & 1);
            MC_TRACE(1);
#pragma unroll
            for (int r = 0; r < RPC; ++r) {
                float tv[8], tw[8];
                mc_bf16x8_to_f32(t0[r], tv);
                mc_bf16x8_to_f32(t1[r], tw);
                const float4* pp = reinterpret_cast<const float4*>(sP + r * H + f0);
                const float4 pa = pp[0], pb = pp[1];
                const float pv[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
                uint32_t o[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaxf(pv[2 * v] + (tv[2 * v] + tw[2 * v]), 0.f),
                                                                    fmaxf(pv[2 * v + 1] + (tv[2 * v + 1] + tw[2 * v + 1]), 0.f));
                    o[v] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                *reinterpret_cast<uint4*>(p.x1 + (size_t)(row0 + r) * H + f0) = make_uint4(o[0], o[1], o[2], o[3]);
            }
            MC_TRACE(2);
            // ---- all-gather signal: one release arrive on every peer's barrier (threads 0..7, ordered behind all E stores) ----
            mc_bar_sync(1, 128);
            if (tidE < MC_CS) mc_arrive_remote_release(mc_mapa(l_flag, (uint32_t)tidE));
            if (tidE == 0) mbar_arrive(p_free);           // P consumed by every E thread
            MC_TRACE(3);
            // ---- epilogue 1: D1 (4 accumulators) + bias, ReLU -> bf16 swizzled B operand of the output GEMM ----
            mbar_wait(bar_d1, k & 1);
            MC_TRACE(4);
            tc_fence_after();
            {
                const int f = 32 * q4 + lane;             // feature inside the slice = K index of the output GEMM
                const float bv = __ldg(p.b_hid + c * 128 + f);
                uint8_t* base = sX2 + (f >> 6) * Lay::KB_BYTES + (f & 7) * 2;
                const int chunk = (f & 63) >> 3;
#pragma unroll
                for (int h2 = 0; h2 < NT / 16; ++h2) {
                    if (16 * h2 >= R) break;              // columns beyond the real rows (the overrun atom) are never read
                    float v[16];
                    tmem_ld16(tm_d + lane_base + 16 * h2, v);
#pragma unroll
                    for (int w = 1; w < 4; ++w) {         // fixed order: issuer 0, 1, 2, 3
                        float x[16];
                        tmem_ld16(tm_d + lane_base + w * NT + 16 * h2, x);
#pragma unroll
                        for (int n = 0; n < 16; ++n) v[n] += x[n];
                    }
#pragma unroll
                    for (int n = 0; n < 16; ++n) {
                        const int row = 16 * h2 + n;
                        if (row < R)
                            *reinterpret_cast<__nv_bfloat16*>(base + (row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4)) =
                                __float2bfloat16(fmaxf(v[n] + bv, 0.f));
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();                     // generic smem writes -> visible to the UMMA operand reads
            mc_bar_sync(1, 128);
            if (tidE == 0) mbar_arrive(x2_ready);
            MC_TRACE(5);
            // ---- epilogue 2: split-K partial logits -> staging [row][256] (the idle x1 tile) -> one bulk copy per owner ----
            mbar_wait(bar_d2, k & 1);
            MC_TRACE(6);
            tc_fence_after();
            {
                float* stg = reinterpret_cast<float*>(sX1);
#pragma unroll
                for (int t2 = 0; t2 < 2; ++t2) {
                    const int o = t2 * 128 + 32 * q4 + lane;
#pragma unroll
                    for (int h2 = 0; h2 < NT / 16; ++h2) {
                        if (16 * h2 >= R) break;
                        uint32_t v[16], x[16];
                        tmem_ld16_nowait(tm_d + lane_base + (2 * t2) * NT + 16 * h2, v);
                        tmem_ld16_nowait(tm_d + lane_base + (2 * t2 + 1) * NT + 16 * h2, x);
                        tmem_ld_wait();
#pragma unroll
                        for (int n = 0; n < 16; ++n) {
                            const int row = 16 * h2 + n;
                            if (row < R) stg[row * SRNN_Q + o] = __uint_as_float(v[n]) + __uint_as_float(x[n]);
                        }
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();                     // staging writes -> visible to the bulk-copy engine
            mc_bar_sync(1, 128);
            if (tidE < MC_CS) {                           // rows d*RPC .. +RPC of the staging -> slot c of CTA d's landing zone;
                const uint32_t d = (uint32_t)((c + tidE) & (MC_CS - 1));   // rotated: no owner is last in every sender's queue
                mc_bulk_s2c(mc_mapa(l_land + (uint32_t)(c * RPC * 1024), d), l_stg + d * (RPC * 1024), RPC * 1024,
                            mc_mapa(l_landbar, d));
            }
            MC_TRACE(7);
            // ---- the 8 split-K partials of the owned rows have landed: reduce in fixed order, + bias ----
            mbar_wait(land_full, k & 1);
            MC_TRACE(8);
            // ---- reduce (fixed order: bias, slice 0, 1, ...) + log-softmax + defined sampler, one warp per owned row ----
            if (warp - 4 < RPC) {
                const int r2 = warp - 4;
                const int bb = row0 + r2;
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = bo[j];
                float4 pa[MC_CS], pb[MC_CS];             // all sixteen shared-memory loads in flight before the first add
#pragma unroll
                for (int s = 0; s < MC_CS; ++s) {
                    const float4* lp = reinterpret_cast<const float4*>(sLand + (s * RPC + r2) * SRNN_Q + lane * 8);
                    pa[s] = lp[0];
                    pb[s] = lp[1];
                }
#pragma unroll
                for (int s = 0; s < MC_CS; ++s) {
                    v[0] += pa[s].x; v[1] += pa[s].y; v[2] += pa[s].z; v[3] += pa[s].w;
                    v[4] += pb[s].x; v[5] += pb[s].y; v[6] += pb[s].z; v[7] += pb[s].w;
                }
                MC_TRACE(11);
                float m = v[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) m = fmaxf(m, v[j]);
                m = mc_warp_max(m);
                // MUFU-based exp / log (ex2.approx, lg2.approx; relative error ~2^-21): the libm forms cost ~20 instructions each
                // and sit on the serial path of every sample; far inside the bf16 mode's stated tolerance.
                // The defined sampler normalises by its own total (idx = #{k: cdf[k] <= u * total}), so it is fed the
                // UNNORMALISED e = exp(logit - max): the log-sum-exp (warp sum, log, 8 subtractions, 8 more exponentials) leaves
                // the serial path and is only evaluated -- after the sample has been published -- when log-probs are requested.
                float e[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) e[j] = __expf(v[j] - m);
                const int t = i - p.lookback;
                if (bb < p.B) {
                    const float u = sU[r2];
                    const int idx = sampler_warp(e, u, lane);
                    if (lane == 0) {
                        p.seq[(size_t)bb * p.Lseq + i] = (uint8_t)idx;
                        sQ[r2 * 32 + (i & 31)] = (uint8_t)idx;
                    }
                    if (p.logp_out) {
                        float s = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) s += e[j];
                        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                        const float lse = m + __logf(s);
                        float4* o4 = reinterpret_cast<float4*>(p.logp_out + ((size_t)bb * p.T + t) * SRNN_Q + lane * 8);
                        o4[0] = make_float4(v[0] - lse, v[1] - lse, v[2] - lse, v[3] - lse);
                        o4[1] = make_float4(v[4] - lse, v[5] - lse, v[6] - lse, v[7] - lse);
                    }
                } else if (lane == 0) {
                    sQ[r2 * 32 + (i & 31)] = 128;
                }
            }
            MC_TRACE(12);
            mc_bar_sync(1, 128);                          // sample i of every owned row is in sQ
            if (tidE == 0) mbar_arrive(&q_ready[k % 3]);
            MC_TRACE(9);
        }
    } else {
        // ===================== G warps: P of the following samples (the first one was prefetched during the prologue) ==========
        for (int g = 1; g < p.nsteps; ++g) g_step(g);
        if (p.pcarry) {                                   // table parts of the next launch's first two samples
            g_step(p.nsteps);
            g_step(p.nsteps + 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                   // no CTA leaves while a peer may still signal or copy into it
    if (warp == 8) tmem_dealloc<MC_TMEM_COLS>(tmem);
}

// pcarry of the first frame of a call: every tap of its first sample is the q_zero prefix (model.py:459); same summation order
// as the gather warps (taps 0 .. FS-3)
__global__ void k_mc_carry_init(const __nv_bfloat16* __restrict__ tbl, int FS, int H, int rows, int q_zero, float* __restrict__ pcarry) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= H) return;
    float acc = 0.f;
    for (int j = 0; j < FS - 2; ++j) acc += __bfloat162float(tbl[((size_t)j * SRNN_Q + q_zero) * H + f]);
    for (int r = blockIdx.y; r < 2 * rows; r += gridDim.y) pcarry[(size_t)r * H + f] = acc;     // both planes
}
int mlp_cluster_carry_init(const __nv_bfloat16* tbl, int FS, int H, int rows, int q_zero, float* pcarry, cudaStream_t st) {
    SRNN_LAUNCH(k_mc_carry_init, dim3(cdiv(H, 256), rows < 64 ? rows : 64), 256, 0, st, tbl, FS, H, rows, q_zero, pcarry);
    return SRNN_OK;
}

bool mlp_cluster_supported(int H, int FS, int B, int max_clusters) {
    if (H != MC_H || FS < 3 || FS > 31 || B < 1) return false;
    return (B + 23) / 24 <= max_clusters;
}

int mlp_cluster_rows(int B, int max_clusters) {           // rows per cluster: 16 when that fits the co-resident clusters, else 24
    return (B + 15) / 16 <= max_clusters ? 16 : 24;
}

// number of 8-CTA clusters of this kernel that can be co-resident (15 on a B200: GPC floor-sweeping); 0 = unavailable
int mlp_cluster_max_clusters() {
    static int cached = -1;
    if (cached >= 0) return cached;
    cached = 0;
    const size_t smem = McLayout<3>::BYTES;
    if (cudaFuncSetAttribute(k_mlp_cluster<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(k_mlp_cluster<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)McLayout<2>::BYTES) != cudaSuccess) {
        cudaGetLastError();
        return cached;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(MC_CS * 16);
    cfg.blockDim = dim3(MC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = MC_CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_mlp_cluster<3>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    cached = n;
    return cached;
}

// p.x1 must hold n_clusters * R rows of H bf16 (R = mlp_cluster_rows); p.part / p.ctr are unused
int mlp_cluster_launch(const __nv_bfloat16* w_hid16, const __nv_bfloat16* w_out16, const MlpPersistParams& p, int rows_per_cluster,
                       cudaStream_t st) {
    const int H = p.H, R = rows_per_cluster, n_clusters = (p.B + R - 1) / R;
    if (H != MC_H || (R != 16 && R != 24)) return fail(SRNN_ERR_ARG, "k_mlp_cluster: unsupported shape");
    CUtensorMap tmWh, tmWo, tmX1;
    SRNN_TRY(make_tmap_bf16(&tmWh, w_hid16, H, H, H, 128));
    SRNN_TRY(make_tmap_bf16(&tmWo, w_out16, SRNN_Q, H, H, 128));
    SRNN_TRY(make_tmap_bf16_kb(&tmX1, p.x1, (uint64_t)n_clusters * R, H, H, R, 4));
    const size_t smem = R == 16 ? McLayout<2>::BYTES : McLayout<3>::BYTES;
    if (mlp_cluster_max_clusters() <= 0) return fail(SRNN_ERR_CUDA, "k_mlp_cluster: clusters unavailable");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(n_clusters * MC_CS);
    cfg.blockDim = dim3(MC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = MC_CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 2 : 1;
    cudaError_t e = R == 16 ? cudaLaunchKernelEx(&cfg, k_mlp_cluster<2>, tmWh, tmWo, tmX1, w_hid16, p)
                            : cudaLaunchKernelEx(&cfg, k_mlp_cluster<3>, tmWh, tmWo, tmX1, w_hid16, p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail(SRNN_ERR_CUDA, "k_mlp_cluster launch: %s", cudaGetErrorString(e));
    return SRNN_OK;
}

}  // namespace srnn
