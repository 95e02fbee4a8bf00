// Persistent GRU recurrence kernels for the teacher-forced (training) path in SRNN_MODE_BF16: ONE cooperative launch runs all
// F frames of one GRU layer (model.py:244 `self.rnn(input, hidden)` = torch nn.GRU; its autograd for the backward kernel),
// instead of one GEMM launch + one gate launch per frame.
//
// Decomposition: CTA (c, rs) owns the GP_HS = 16 hidden units u0 = 16c .. u0+15 for the 64 utterances (rows) rs*64 .. +63
// (the two row halves of a 128-utterance batch are independent recurrences with their own frame barrier).  It keeps the
// recurrent weight rows of those units resident in shared memory for the whole launch (UMMA B operand, loaded once by TMA):
//     forward : W_hh[g*H + u0 .. +16, :]  for the three gates g = r, z, n      (48 x H   bf16 = 96 KB at H = 1024)
//     backward: W_hh^T[u0 .. +16, :]                                           (16 x 3H  bf16 = 96 KB at H = 1024)
// and streams the recurrent activations of the frame (h_{f-1}: rows x H, resp. dGH_f: rows x 3H, bf16) through a TMA ring as
// the UMMA A operand (M = 64 rows).  The accumulator D[row][gate*16 + j] puts r, z, n of one (utterance, unit) into the SAME
// thread, and the fp32 state h / dh lives in registers across frames.  An M = 64 accumulator occupies lanes 0..15 of each
// TMEM lane quadrant; a second one sits in lanes 16..31 ("interleaved"), so the four issuers' private accumulators fill
// all 128 lanes, one shuffle exchange combines the halves, and every one of the 128 row threads does the gate math of
// 8 units of one utterance.  Between frames the CTAs exchange the new bf16 state through global memory (L2) behind a release/acquire
// counter barrier (all CTAs are co-resident: cooperative launch).
// Warp roles (288 threads): 0..3 = row/epilogue warps (TMEM lane quadrant = warp), 4 = TMA producer, 5..8 = MMA issuers with
// private accumulators (k-block kb belongs to issuer kb % 4; a single thread cannot issue small tcgen05.mma fast enough and
// a single accumulator chain serialises, see mlp_persist.cu).
#include "common.cuh"
#include "umma.cuh"

namespace srnn {

using namespace ptx;
typedef __nv_bfloat16 bf;

constexpr int GP_HS = 16;
constexpr int GP_THREADS = 288;
constexpr int GP_ISSUERS = 4;
constexpr int GP_ROWS = 64;                   // utterances per CTA (UMMA M)
constexpr int GP_STAGES = 12;                 // TMA ring: 12 x (64 rows x 128 B) = 96 KB in flight per SM
constexpr int GP_STAGE_BYTES = GP_ROWS * 128;
// TMEM map (512 columns): issuer w owns lane half (w & 1) of column group (w >> 1) * 256; inside its group it rotates over
// GP_NACC independent accumulators 64 columns apart.  MMAs that accumulate into the SAME TMEM tile serialise on its
// latency (~290 cycles for these tiny tiles: 48 chained MMAs per frame were 13.8k of the 23.6k cycles of a backward frame).
constexpr uint32_t GP_TMEM_COLS = 512;
constexpr int GP_NACC = 4;

__device__ __forceinline__ void gp_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned gp_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void gp_red_release(unsigned* p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(p) : "memory");
}
__device__ __forceinline__ void gp_fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void gp_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }
// MUFU-based gate functions (ex2 + rcp): ~1e-6 absolute error, a quarter of the instructions of expf / tanhf
__device__ __forceinline__ float gp_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float gp_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

__device__ __forceinline__ void ld16(const float* __restrict__ p, float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 t = reinterpret_cast<const float4*>(p)[i];
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ void st16(float* __restrict__ p, const float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void st16_bf16(bf* __restrict__ p, const float (&v)[16]) {
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        o[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    reinterpret_cast<uint4*>(p)[0] = make_uint4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<uint4*>(p)[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

__device__ __forceinline__ void ld8(const float* __restrict__ p, float (&v)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* __restrict__ p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8_bf16(bf* __restrict__ p, const float (&v)[8]) {
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        o[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(o[0], o[1], o[2], o[3]);
}
// Sum of this half-warp's valid accumulators (column groups 0 and 64) for 16 consecutive columns starting at col, then
// the halves' exchange: lanes 0..15 keep columns 0..7, lanes 16..31 columns 8..15 of (own + partner) sums.
__device__ __forceinline__ void gp_gather8(uint32_t tlane, int col, int ncg, int hh, float (&out)[8]) {
    float lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) lo[i] = hi[i] = 0.f;
#pragma unroll
    for (int g = 0; g < 2; ++g) {                    // column group of issuer (hh + 2g); warp-collective loads, selected after
        uint32_t t[GP_NACC][16];
#pragma unroll
        for (int a = 0; a < GP_NACC; ++a) tmem_ld16_nowait(tlane + g * 256 + a * 64 + col, t[a]);
        tmem_ld_wait();
        if (ncg > g) {
#pragma unroll
            for (int a = 0; a < GP_NACC; ++a)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    lo[i] += __uint_as_float(t[a][i]);
                    hi[i] += __uint_as_float(t[a][8 + i]);
                }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float mine = hh ? hi[i] : lo[i];
        const float send = hh ? lo[i] : hi[i];
        out[i] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
    }
}

struct GruFwdParams {
    int B, F, H, rotate, gsz;
    const float* GI;     // (B*F, 3H)  W_ih x + b_ih of every frame (row b*F + f)
    float* GH;           // (B*F, 3H)  out: W_hh h_{f-1} + b_hh (saved for the backward pass)
    float* Y;            // (B*F, H)   out: h_f
    bf* Y16;             // (B*F, H)   out: bf16 copy = recurrent operand of frame f+1 and input of the next layer
    const float* h0;     // (B, H)     initial state (fp32 master)
    float* h_last;       // (B, H)     out: h_{F-1} (TBPTT carry, model.py:348) or null
    const float* b_hh;   // (3H)
    unsigned* ctr;       // frame barrier counter (zeroed by the launcher)
};

struct GruBwdParams {
    int B, F, H, rotate, gsz;
    const float* GI;     // saved forward projections (B*F, 3H)
    const float* GH;
    const float* Y;      // (B*F, H) forward outputs; h_{f-1} = Y[b, f-1] or h0[b]
    const float* h0;     // (B, H)
    const float* dY;     // (B*F, H)  gradient w.r.t. the layer outputs
    float* dGI;          // (B*F, 3H) out: gradient w.r.t. W_ih x + b_ih   (fp32 copy; may be null)
    float* dGH;          // (B*F, 3H) out: gradient w.r.t. W_hh h + b_hh   (fp32 copy; may be null)
    float* bias_part;    // (row halves, 4, H) out: column sums over (utterance, frame) of dpr, dpz, dpn, dpn*r = the bias
                         //                    gradients (db_ih = r,z,n parts; db_hh = r,z,n*r parts), or null
    bf* dGI16;           // bf16 copies (operands of the weight-gradient / input-gradient GEMMs)
    bf* dGH16;           //            (and the recurrent operand of this kernel)
    float* dh0;          // (B, H)    out: dL/dh_{-1}
    unsigned* ctr;
    long long* trace;    // development aid (SRNN_TRACE_GRU=1): clock64 stamps of CTA (0,0), 8 per frame
};

// ---------------------------------------------------------------------------------------------------------------------
// shared skeleton: TMA producer + MMA issuers.  NB = B-operand rows (48 forward / 16 backward), KB = K / 64.
// Frame index `step` runs 0 .. F-1 in processing order; the A operand of processing step s is fetched at column offset
// a_col0(s) of tmA_first (s == 0 and first_separate) or tmA.
// ---------------------------------------------------------------------------------------------------------------------
template <int NB>
struct GpSmem {
    static constexpr int W_KB_BYTES = NB * 128;          // one k-block of the resident weight slice
};

template <int NB, bool FWD>
__device__ __forceinline__ void gp_producer(const CUtensorMap* tmA0, const CUtensorMap* tmA, uint8_t* sRing, uint64_t* full,
                                            uint64_t* empty, const unsigned* ctr, int F, int KB, int K, int NS, int row0,
                                            int rot, int gsz, long long* trace = nullptr) {
    int it = 0;
    for (int s = 0; s < F; ++s) {
        // forward : frame f = s reads h_{f-1} (h0 for f = 0, published before launch) -> wait for s * NS arrivals
        // backward: frame f = F-1-s reads dGH_f written in THIS step -> wait for (s+1) * NS arrivals
        const unsigned target = (unsigned)(FWD ? s : s + 1) * (unsigned)NS;
        if (target) {
            while (gp_ld_acquire(ctr) < target) {
            }
            gp_fence_proxy_async();                      // other CTAs' generic-proxy writes -> visible to TMA reads
        }
        if (trace) trace[s * 8 + 2] = clock64();
        const CUtensorMap* tm = (FWD && s == 0) ? tmA0 : tmA;
        const int col0 = FWD ? (s == 0 ? 0 : (s - 1) * K) : (F - 1 - s) * K;
        // every CTA of a row half streams the same tiles: each starts at its own k-block (rot) so that at any moment the
        // CTAs pull different tiles from different L2 slices instead of all hitting the same few
        // one TMA instruction fetches gsz consecutive k-blocks (3-D box) into one ring stage of gsz * 8 KB
        const int ngrp = KB / gsz, nst = GP_STAGES / gsz;
        int kk = rot;
        for (int g = 0; g < ngrp; ++g, ++it) {
            const int st = it % nst;
            const uint32_t ph = (it / nst) & 1;
            mbar_wait(&empty[st], ph ^ 1);
            mbar_expect_tx(&full[st], (uint32_t)gsz * GP_STAGE_BYTES);
            tma_load_3d(sRing + (size_t)st * gsz * GP_STAGE_BYTES, tm, &full[st], 0, row0, col0 / 64 + kk);
            kk += gsz;
            if (kk >= KB) kk -= KB;
        }
    }
}

template <int NB>
__device__ __forceinline__ void gp_issuer(int w, int nissue, uint8_t* sW, uint8_t* sRing, uint64_t* full, uint64_t* empty,
                                          uint64_t* w_ready, uint64_t* bar_d, uint32_t tmem, int F, int KB, int rot, int gsz) {
    constexpr uint32_t idesc = umma_idesc_bf16(GP_ROWS, NB);
    mbar_wait(w_ready, 0);
    const uint64_t dW0 = umma_desc_sw128(smem_u32(sW));
    const uint64_t dA0 = umma_desc_sw128(smem_u32(sRing));
    // issuers 0, 2 -> lanes 0..15 of every quadrant (column groups 0 / 64); issuers 1, 3 -> lanes 16..31
    const uint32_t dacc = tmem + ((w & 1) ? (16u << 16) : 0u) + (uint32_t)(w >> 1) * 256;
    const int ngrp = KB / gsz, nst = GP_STAGES / gsz;
    for (int s = 0; s < F; ++s) {
        for (int kb = w; kb < KB; kb += GP_ISSUERS) {
            const int g = kb / gsz, it = s * ngrp + g;
            const int st = it % nst;
            const uint32_t ph = (it / nst) & 1;
            mbar_wait(&full[st], ph);
            tc_fence_after();
            const uint64_t da = dA0 + (uint64_t)((st * gsz + kb % gsz) * (GP_STAGE_BYTES >> 4));
            int kw = kb + rot;                                   // the streamed tile of slot kb is k-block (kb + rot) mod KB
            if (kw >= KB) kw -= KB;
            const uint64_t db = dW0 + (uint64_t)(kw * (GpSmem<NB>::W_KB_BYTES >> 4));
            const uint32_t acc = kb >= GP_ISSUERS;               // each of the four accumulators starts at this issuer's first k-block
            umma_bf16(dacc, da, db, idesc, acc);
            umma_bf16(dacc + 64, da + 2, db + 2, idesc, acc);
            umma_bf16(dacc + 128, da + 4, db + 4, idesc, acc);
            umma_bf16(dacc + 192, da + 6, db + 6, idesc, acc);
            umma_commit(&empty[st]);                             // (a stage of gsz k-blocks collects one arrival per k-block)
        }
        umma_commit(bar_d);
    }
    (void)nissue;
}

struct GpLayout {
    uint8_t* sW;
    uint8_t* sRing;
    float* sBias;
    uint64_t* w_ready;
    uint64_t* full;
    uint64_t* empty;
    uint64_t* bar_d;
    uint32_t* tmem_slot;
};
__device__ __forceinline__ GpLayout gp_layout(uint8_t* smem_raw, size_t w_bytes) {
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    GpLayout L;
    L.sW = smem;
    L.sRing = smem + w_bytes;                                   // w_bytes is a multiple of 1024
    L.sBias = (float*)(L.sRing + (size_t)GP_STAGES * GP_STAGE_BYTES);
    uint64_t* bars = (uint64_t*)(L.sBias + 64);
    L.w_ready = bars;
    L.full = bars + 1;
    L.empty = L.full + GP_STAGES;
    L.bar_d = L.empty + GP_STAGES;
    L.tmem_slot = (uint32_t*)(L.bar_d + 1);
    return L;
}
static size_t gp_smem_bytes(size_t w_bytes) { return 1024 + w_bytes + (size_t)GP_STAGES * GP_STAGE_BYTES + 256 + 8 * (2 * GP_STAGES + 3) + 64; }

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GP_THREADS, 1)
k_gru_persist_fwd(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH0,
                  const __grid_constant__ CUtensorMap tmY, const GruFwdParams p) {
    constexpr int NB = 3 * GP_HS;
    const int H = p.H, F = p.F, KB = H >> 6, NS = gridDim.x, rs = blockIdx.y;
    const int nissue = KB < GP_ISSUERS ? KB : GP_ISSUERS;
    const int u0 = blockIdx.x * GP_HS;
    const int gsz = p.gsz;                                    // k-blocks per TMA box / ring stage
    const int rot = p.rotate ? (int)((blockIdx.x * 5u) % (unsigned)(KB / gsz)) * gsz : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    extern __shared__ uint8_t smem_raw[];
    const GpLayout L = gp_layout(smem_raw, (size_t)KB * GpSmem<NB>::W_KB_BYTES);

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmW);
        prefetch_tmap(&tmH0);
        prefetch_tmap(&tmY);
        mbar_init(L.w_ready, 1);
        for (int s = 0; s < GP_STAGES; ++s) {
            mbar_init(&L.full[s], 1);
            mbar_init(&L.empty[s], gsz);
        }
        mbar_init(L.bar_d, nissue);
        fence_barrier_init();
    }
    if (threadIdx.x < NB) L.sBias[threadIdx.x] = p.b_hh[(threadIdx.x >> 4) * H + u0 + (threadIdx.x & 15)];
    if (warp == 5) tmem_alloc<GP_TMEM_COLS>(L.tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *L.tmem_slot, 0);

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(L.w_ready, (uint32_t)(KB * GpSmem<NB>::W_KB_BYTES));
            for (int kb = 0; kb < KB; ++kb)
                for (int g = 0; g < 3; ++g)
                    tma_load_2d(L.sW + (size_t)kb * GpSmem<NB>::W_KB_BYTES + g * (GP_HS * 128), &tmW, L.w_ready, kb * 64, g * H + u0);
            gp_producer<NB, true>(&tmH0, &tmY, L.sRing, L.full, L.empty, p.ctr + rs, F, KB, H, NS, rs * GP_ROWS, rot, gsz);
        }
    } else if (warp >= 5) {
        if (lane == 0 && warp - 5 < nissue)
            gp_issuer<NB>(warp - 5, nissue, L.sW, L.sRing, L.full, L.empty, L.w_ready, L.bar_d, tmem, F, KB, rot, gsz);
    } else {
        // ===================== row warps: thread = (utterance row, 8 of the 16 units) =====================
        const int hh = lane >> 4;                            // unit half: units u0 + 8*hh .. +7
        const int b = rs * GP_ROWS + 16 * warp + (lane & 15);
        const bool ok = b < p.B;
        const int uq = u0 + 8 * hh;
        const int ncg = (nissue > hh) + (nissue > hh + 2);   // valid accumulators in this lane half (issuers hh, hh + 2)
        const uint32_t tlane = tmem + ((uint32_t)(32 * warp) << 16);
        float h[8];
        if (ok) ld8(p.h0 + (size_t)b * H + uq, h);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = 0.f;
        }
        const size_t row_stride = (size_t)3 * H;
        for (int f = 0; f < F; ++f) {
            const size_t r = (size_t)b * F + f;
            float gi[3][8];
            if (ok) {
                const float* gp = p.GI + r * row_stride + uq;
                ld8(gp, gi[0]);
                ld8(gp + H, gi[1]);
                ld8(gp + 2 * H, gi[2]);
                if (f + 1 < F) {                             // next frame's projections -> L2 while this frame computes
                    gp_prefetch_l2(gp + row_stride);
                    gp_prefetch_l2(gp + row_stride + H);
                    gp_prefetch_l2(gp + row_stride + 2 * H);
                }
            }
            mbar_wait(L.bar_d, f & 1);
            tc_fence_after();
            float a[3][8];
#pragma unroll
            for (int g = 0; g < 3; ++g) gp_gather8(tlane, 16 * g, ncg, hh, a[g]);
            tc_fence_before();
            if (ok) {
                float hn[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    a[0][j] += L.sBias[8 * hh + j];
                    a[1][j] += L.sBias[16 + 8 * hh + j];
                    a[2][j] += L.sBias[32 + 8 * hh + j];
                    const float rr = gp_sigmoid(gi[0][j] + a[0][j]);
                    const float zz = gp_sigmoid(gi[1][j] + a[1][j]);
                    const float nn = gp_tanh(gi[2][j] + rr * a[2][j]);
                    hn[j] = (1.f - zz) * nn + zz * h[j];
                    h[j] = hn[j];
                }
                st8_bf16(p.Y16 + r * H + uq, hn);            // first: this is what the other CTAs wait for
                st8(p.Y + r * H + uq, hn);
                float* ghp = p.GH + r * row_stride + uq;
                st8(ghp, a[0]);
                st8(ghp + H, a[1]);
                st8(ghp + 2 * H, a[2]);
            }
            gp_bar_sync(1, 128);
            if (threadIdx.x == 0) gp_red_release(p.ctr + rs);   // publishes the whole CTA's h_f slice
        }
        if (ok && p.h_last) st8(p.h_last + (size_t)b * H + uq, h);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc<GP_TMEM_COLS>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------------
// backward (BPTT through the frames of one layer): per frame f = F-1 .. 0
//   dh = dY_f + carry;  gates recomputed from the saved projections;  dGI_f, dGH_f;  carry = dh*z + dGH_f . W_hh
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GP_THREADS, 1)
k_gru_persist_bwd(const __grid_constant__ CUtensorMap tmWt, const __grid_constant__ CUtensorMap tmG, const GruBwdParams p) {
    constexpr int NB = GP_HS;
    const int H = p.H, F = p.F, K3 = 3 * H, KB = K3 >> 6, NS = gridDim.x, rs = blockIdx.y;
    const int nissue = KB < GP_ISSUERS ? KB : GP_ISSUERS;
    const int u0 = blockIdx.x * GP_HS;
    const int gsz = p.gsz;                                    // k-blocks per TMA box / ring stage
    const int rot = p.rotate ? (int)((blockIdx.x * 5u) % (unsigned)(KB / gsz)) * gsz : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    extern __shared__ uint8_t smem_raw[];
    const GpLayout L = gp_layout(smem_raw, (size_t)KB * GpSmem<NB>::W_KB_BYTES);

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmWt);
        prefetch_tmap(&tmG);
        mbar_init(L.w_ready, 1);
        for (int s = 0; s < GP_STAGES; ++s) {
            mbar_init(&L.full[s], 1);
            mbar_init(&L.empty[s], gsz);
        }
        mbar_init(L.bar_d, nissue);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc<GP_TMEM_COLS>(L.tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *L.tmem_slot, 0);

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(L.w_ready, (uint32_t)(KB * GpSmem<NB>::W_KB_BYTES));
            for (int kb = 0; kb < KB; ++kb)
                tma_load_2d(L.sW + (size_t)kb * GpSmem<NB>::W_KB_BYTES, &tmWt, L.w_ready, kb * 64, u0);
            gp_producer<NB, false>(&tmG, &tmG, L.sRing, L.full, L.empty, p.ctr + rs, F, KB, K3, NS, rs * GP_ROWS, rot, gsz,
                                   (p.trace && blockIdx.x == 0 && blockIdx.y == 0) ? p.trace : nullptr);
        }
    } else if (warp >= 5) {
        if (lane == 0 && warp - 5 < nissue)
            gp_issuer<NB>(warp - 5, nissue, L.sW, L.sRing, L.full, L.empty, L.w_ready, L.bar_d, tmem, F, KB, rot, gsz);
    } else {
        const int hh = lane >> 4;
        const int b = rs * GP_ROWS + 16 * warp + (lane & 15);
        const bool ok = b < p.B;
        const int uq = u0 + 8 * hh;
        const int ncg = (nissue > hh) + (nissue > hh + 2);
        const uint32_t tlane = tmem + ((uint32_t)(32 * warp) << 16);
        float carry[8], sb[4][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) carry[j] = sb[0][j] = sb[1][j] = sb[2][j] = sb[3][j] = 0.f;
        const size_t row_stride = (size_t)K3;
        long long* tr = (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) ? p.trace : nullptr;
        for (int s = 0; s < F; ++s) {
            const int f = F - 1 - s;
            const size_t r = (size_t)b * F + f;
            float dhz[8];
            if (tr) tr[s * 8 + 0] = clock64();
            if (ok) {
                float gr[8], gz[8], gn[8], hr[8], hz[8], hn[8], hp[8], dy[8];
                const float* gip = p.GI + r * row_stride + uq;
                const float* ghp = p.GH + r * row_stride + uq;
                ld8(gip, gr);
                ld8(gip + H, gz);
                ld8(gip + 2 * H, gn);
                ld8(ghp, hr);
                ld8(ghp + H, hz);
                ld8(ghp + 2 * H, hn);
                ld8(f ? p.Y + (r - 1) * H + uq : p.h0 + (size_t)b * H + uq, hp);
                ld8(p.dY + r * H + uq, dy);
                if (f > 0) {                                 // the previous frame's rows -> L2 while this frame's GEMM runs
                    gp_prefetch_l2(gip - row_stride);
                    gp_prefetch_l2(gip - row_stride + H);
                    gp_prefetch_l2(gip - row_stride + 2 * H);
                    gp_prefetch_l2(ghp - row_stride);
                    gp_prefetch_l2(ghp - row_stride + H);
                    gp_prefetch_l2(ghp - row_stride + 2 * H);
                    gp_prefetch_l2(p.dY + (r - 1) * H + uq);
                    if (f > 1) gp_prefetch_l2(p.Y + (r - 2) * H + uq);
                }
                float o_r[8], o_z[8], o_n[8], o_nr[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float rr = gp_sigmoid(gr[j] + hr[j]);
                    const float zz = gp_sigmoid(gz[j] + hz[j]);
                    const float nn = gp_tanh(gn[j] + rr * hn[j]);
                    const float dh = dy[j] + carry[j];
                    const float dn = dh * (1.f - zz);
                    const float dz = dh * (hp[j] - nn);
                    const float dpn = dn * (1.f - nn * nn);
                    const float dr = dpn * hn[j];
                    o_r[j] = dr * rr * (1.f - rr);
                    o_z[j] = dz * zz * (1.f - zz);
                    o_n[j] = dpn;
                    o_nr[j] = dpn * rr;
                    dhz[j] = dh * zz;
                    sb[0][j] += o_r[j];
                    sb[1][j] += o_z[j];
                    sb[2][j] += o_n[j];
                    sb[3][j] += o_nr[j];
                }
                bf* g16 = p.dGH16 + r * row_stride + uq;     // first: the recurrent operand the other CTAs wait for
                st8_bf16(g16, o_r);
                st8_bf16(g16 + H, o_z);
                st8_bf16(g16 + 2 * H, o_nr);
                if (p.dGH) {
                    float* gh = p.dGH + r * row_stride + uq;
                    st8(gh, o_r);
                    st8(gh + H, o_z);
                    st8(gh + 2 * H, o_nr);
                }
                if (p.dGI) {
                    float* gi = p.dGI + r * row_stride + uq;
                    st8(gi, o_r);
                    st8(gi + H, o_z);
                    st8(gi + 2 * H, o_n);
                }
                bf* i16 = p.dGI16 + r * row_stride + uq;
                st8_bf16(i16, o_r);
                st8_bf16(i16 + H, o_z);
                st8_bf16(i16 + 2 * H, o_n);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) dhz[j] = 0.f;
            }
            if (tr) tr[s * 8 + 1] = clock64();
            gp_bar_sync(1, 128);
            if (threadIdx.x == 0) gp_red_release(p.ctr + rs);   // dGH_f of this CTA's units is published
            mbar_wait(L.bar_d, s & 1);
            if (tr) tr[s * 8 + 3] = clock64();
            tc_fence_after();
            float a[8];
            gp_gather8(tlane, 0, ncg, hh, a);
            tc_fence_before();
#pragma unroll
            for (int j = 0; j < 8; ++j) carry[j] = dhz[j] + a[j];
            if (tr) tr[s * 8 + 4] = clock64();
        }
        if (ok && p.dh0) st8(p.dh0 + (size_t)b * H + uq, carry);
        if (p.bias_part) {      // bias gradients: sum over this CTA's 64 utterances (fixed order: lanes, then warps)
            float* sRed = reinterpret_cast<float*>(L.sRing);             // the ring is idle now: [4 warps][4][16]
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float v = sb[a][j];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    if ((lane & 15) == 0) sRed[(warp * 4 + a) * 16 + 8 * hh + j] = v;
                }
            gp_bar_sync(1, 128);
            if (threadIdx.x < 64) {
                const int a = threadIdx.x >> 4, u = threadIdx.x & 15;
                const float v = (sRed[(0 * 4 + a) * 16 + u] + sRed[(1 * 4 + a) * 16 + u]) +
                                (sRed[(2 * 4 + a) * 16 + u] + sRed[(3 * 4 + a) * 16 + u]);
                p.bias_part[((size_t)rs * 4 + a) * H + u0 + u] = v;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc<GP_TMEM_COLS>(tmem);
}

// db_ih (3H) = [sum dpr | sum dpz | sum dpn], db_hh (3H) = [sum dpr | sum dpz | sum dpn*r] over the row halves
__global__ void k_gru_bias_grads(const float* __restrict__ part, int nrs, int H, float* __restrict__ db_ih, float* __restrict__ db_hh) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= H) return;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < nrs; ++r)
#pragma unroll
        for (int a = 0; a < 4; ++a) s[a] += part[((size_t)r * 4 + a) * H + u];
    if (db_ih) { db_ih[u] = s[0]; db_ih[H + u] = s[1]; db_ih[2 * H + u] = s[2]; }
    if (db_hh) { db_hh[u] = s[0]; db_hh[H + u] = s[1]; db_hh[2 * H + u] = s[3]; }
}

// ---------------------------------------------------------------------------------------------------------------------
// Generation-time GRU cell: ONE launch per layer and tier step does  gi = W_ih x + b_ih  on tcgen05 AND the gate math
// (the recurrent projection gh = W_hh h + b_hh does not depend on the new samples and is computed by an earlier launch):
//     h' = (1 - z) n + z h,   r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r gh_n)        (model.py:244 at F = 1)
// CTA (c, rs): 32 hidden units x 64 utterances; B operand = the r/z/n rows of W_ih for those units (96 x K, streamed by 4-D
// TMA boxes of 4 k-blocks), A operand = x (64 rows x K).  Same interleaved M = 64 accumulators / row threads as above.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GC_US = 32;
constexpr int GC_NB = 3 * GC_US;
constexpr int GC_GKB = 4;                                     // k-blocks per TMA box
constexpr int GC_RING = 2;                                    // boxes in flight
constexpr int GC_A_BYTES = GC_GKB * GP_ROWS * 128;            // 32 KB
constexpr int GC_W_BYTES = GC_GKB * GC_NB * 128;              // 48 KB
constexpr int GC_GROUP_BYTES = GC_A_BYTES + GC_W_BYTES;

struct GruCellParams {
    int B, H;
    const float* b_ih;   // (3H)
    const float* GH;     // (B, 3H)  W_hh h + b_hh of this step
    float* h;            // (B, H)   fp32 state, updated in place
    bf* h16;             // (B, H)   bf16 copy of the new state
};

__global__ void __launch_bounds__(GP_THREADS, 1)
k_gru_cell_gen(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const GruCellParams p) {
    const int H = p.H, NG = (H >> 6) / GC_GKB, rs = blockIdx.y, u0 = blockIdx.x * GC_US;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sRing = smem;
    float* sBias = (float*)(sRing + (size_t)GC_RING * GC_GROUP_BYTES);
    uint64_t* full = (uint64_t*)(sBias + GC_NB);
    uint64_t* empty = full + GC_RING;
    uint64_t* bar_d = empty + GC_RING;
    uint32_t* tmem_slot = (uint32_t*)(bar_d + 1);

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmX);
        prefetch_tmap(&tmW);
        for (int s = 0; s < GC_RING; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], GP_ISSUERS);
        }
        mbar_init(bar_d, GP_ISSUERS);
        fence_barrier_init();
    }
    pdl_trigger();
    if (threadIdx.x < GC_NB) sBias[threadIdx.x] = p.b_ih[(threadIdx.x / GC_US) * H + u0 + (threadIdx.x % GC_US)];
    if (warp == 5) tmem_alloc<GP_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // the weights are constants of the call: the W part of the first ring stages is requested BEFORE griddepcontrol.wait, so
    // that as a programmatic dependent (of the lite cell / the previous layer) the launch already streams them while its
    // predecessor finishes; x, gh and the state come from the preceding kernels and are only touched after the wait
    const int early = NG < GC_RING ? NG : GC_RING;
    if (warp == 4 && lane == 0) {
        for (int g = 0; g < early; ++g) {
            mbar_expect_tx(&full[g], GC_GROUP_BYTES);
            tma_load_4d(sRing + (size_t)g * GC_GROUP_BYTES + GC_A_BYTES, &tmW, &full[g], 0, u0, 0, g * GC_GKB);
        }
    }
    pdl_wait();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 4) {
        if (lane == 0) {
            for (int g = 0; g < NG; ++g) {
                const int st = g % GC_RING;
                const uint32_t ph = (g / GC_RING) & 1;
                uint8_t* dst = sRing + (size_t)st * GC_GROUP_BYTES;
                if (g >= early) {
                    mbar_wait(&empty[st], ph ^ 1);
                    mbar_expect_tx(&full[st], GC_GROUP_BYTES);
                    tma_load_4d(dst + GC_A_BYTES, &tmW, &full[st], 0, u0, 0, g * GC_GKB);
                }
                tma_load_3d(dst, &tmX, &full[st], 0, rs * GP_ROWS, g * GC_GKB);
            }
        }
    } else if (warp >= 5) {
        if (lane == 0) {
            const int w = warp - 5;                                      // issuer w takes k-block w of every box
            constexpr uint32_t idesc = umma_idesc_bf16(GP_ROWS, GC_NB);
            const uint32_t dacc = tmem + ((w & 1) ? (16u << 16) : 0u) + (uint32_t)(w >> 1) * 128;
            for (int g = 0; g < NG; ++g) {
                const int st = g % GC_RING;
                const uint32_t ph = (g / GC_RING) & 1;
                mbar_wait(&full[st], ph);
                tc_fence_after();
                const uint32_t base = smem_u32(sRing + (size_t)st * GC_GROUP_BYTES);
                const uint64_t da = umma_desc_sw128(base + w * (GP_ROWS * 128));
                const uint64_t db = umma_desc_sw128(base + GC_A_BYTES + w * (GC_NB * 128));
                umma_bf16(dacc, da, db, idesc, g > 0);
                umma_bf16(dacc, da + 2, db + 2, idesc, 1);
                umma_bf16(dacc, da + 4, db + 4, idesc, 1);
                umma_bf16(dacc, da + 6, db + 6, idesc, 1);
                umma_commit(&empty[st]);
            }
            umma_commit(bar_d);
        }
    } else {
        const int hh = lane >> 4;                                        // unit half: 16 of the CTA's 32 units
        const int b = rs * GP_ROWS + 16 * warp + (lane & 15);
        const bool ok = b < p.B;
        const int uq = u0 + 16 * hh;
        const uint32_t tlane = tmem + ((uint32_t)(32 * warp) << 16);
        float gh[3][16], h[16];
        if (ok) {
            const float* ghp = p.GH + (size_t)b * 3 * H + uq;
            ld16(ghp, gh[0]);
            ld16(ghp + H, gh[1]);
            ld16(ghp + 2 * H, gh[2]);
            ld16(p.h + (size_t)b * H + uq, h);
        }
        mbar_wait(bar_d, 0);
        tc_fence_after();
        float gi[3][16];
#pragma unroll
        for (int g = 0; g < 3; ++g) {          // one accumulator per issuer (measured: splitting the 16-MMA chain did not pay)
            uint32_t a0[16], a1[16], b0[16], b1[16];
            tmem_ld16_nowait(tlane + 32 * g, a0);
            tmem_ld16_nowait(tlane + 32 * g + 16, a1);
            tmem_ld16_nowait(tlane + 128 + 32 * g, b0);
            tmem_ld16_nowait(tlane + 128 + 32 * g + 16, b1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float lo = __uint_as_float(a0[j]) + __uint_as_float(b0[j]);
                const float hi = __uint_as_float(a1[j]) + __uint_as_float(b1[j]);
                const float mine = hh ? hi : lo;
                const float send = hh ? lo : hi;
                gi[g][j] = mine + __shfl_xor_sync(0xffffffffu, send, 16) + sBias[g * GC_US + 16 * hh + j];
            }
        }
        tc_fence_before();
        if (ok) {
            float hn[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float rr = gp_sigmoid(gi[0][j] + gh[0][j]);
                const float zz = gp_sigmoid(gi[1][j] + gh[1][j]);
                const float nn = gp_tanh(gi[2][j] + rr * gh[2][j]);
                hn[j] = (1.f - zz) * nn + zz * h[j];
            }
            st16(p.h + (size_t)b * H + uq, hn);
            st16_bf16(p.h16 + (size_t)b * H + uq, hn);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc<GP_TMEM_COLS>(tmem);
}

// ---- first GRU layer of a tier step with the input expansion folded in (no tensor cores: K = the frame's sample columns) ----
// Latency-bound by construction (a few hundred bytes per thread from L2), so EVERY global load of a thread is issued before the
// first use: one L2 round trip for the folded weights (CL_KMAX x 3), the partial sums, the recurrent projections and the state.
constexpr int CL_RB = 4;        // utterances per CTA: every folded weight loaded from L2 feeds CL_RB FMAs
constexpr int CL_KMAX = 20;     // sample columns handled with all weights in registers (the frame sizes of every BASELINE config)
template <int NK>
__global__ void __launch_bounds__(256)
k_gru_cell_lite(const float* __restrict__ gipre, long long gipre_ld, const float* __restrict__ g_in_t, int k_lo,
                const uint8_t* __restrict__ seq, int seq_ld, int start_static, const int* __restrict__ step_base,
                const float* __restrict__ lut, const float* __restrict__ GH, float* __restrict__ h, bf* __restrict__ h16, int B,
                int H) {
    __shared__ float cl_a[NK * CL_RB];                               // dequantised samples [k][row]
    pdl_trigger();
    const int b0 = blockIdx.x * CL_RB;
    const int u = blockIdx.y * blockDim.x + threadIdx.x;
    const bool live = u < H;
    // the folded weights are constants of the call: requested before griddepcontrol.wait
    float wv[NK][3];
    if (live) {
        const float* w = g_in_t + (size_t)k_lo * 3 * H + u;
#pragma unroll
        for (int k = 0; k < NK; ++k)
#pragma unroll
            for (int g = 0; g < 3; ++g) wv[k][g] = __ldg(w + (size_t)k * 3 * H + g * H);
    }
    pdl_wait();
    const int start = start_static + (step_base ? *step_base : 0);
    for (int e = threadIdx.x; e < NK * CL_RB; e += blockDim.x) {
        const int r = e % CL_RB, k = k_lo + e / CL_RB;
        const int b = b0 + r < B ? b0 + r : B - 1;
        cl_a[e] = lut[seq[(size_t)b * seq_ld + start + k]];
    }
    float acc[3][CL_RB], gh[3][CL_RB], hp[CL_RB];
    if (live) {
#pragma unroll
        for (int r = 0; r < CL_RB; ++r) {
            const int b = b0 + r < B ? b0 + r : B - 1;
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                acc[g][r] = gipre[(size_t)b * gipre_ld + g * H + u];
                gh[g][r] = GH[(size_t)b * 3 * H + g * H + u];
            }
            hp[r] = h[(size_t)b * H + u];
        }
    }
    __syncthreads();
    if (!live) return;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
#pragma unroll
        for (int r = 0; r < CL_RB; ++r) {
            const float a = cl_a[k * CL_RB + r];
#pragma unroll
            for (int g = 0; g < 3; ++g) acc[g][r] = fmaf(a, wv[k][g], acc[g][r]);
        }
    }
#pragma unroll
    for (int r = 0; r < CL_RB; ++r) {
        const int b = b0 + r;
        if (b >= B) break;
        const float rr = gp_sigmoid(acc[0][r] + gh[0][r]);
        const float zz = gp_sigmoid(acc[1][r] + gh[1][r]);
        const float nn = gp_tanh(acc[2][r] + rr * gh[2][r]);
        const float hn = (1.f - zz) * nn + zz * hp[r];
        h[(size_t)b * H + u] = hn;
        h16[(size_t)b * H + u] = __float2bfloat16(hn);
    }
}
bool gru_cell_lite_supported(int nk) { return nk == 20 || nk == 16 || nk == 4 || nk == 2; }
int gru_cell_lite(int B, int H, const float* gipre, long long gipre_ld, const float* g_in_t, int k_lo, int k_hi, const uint8_t* seq,
                  int seq_ld, int start_static, const int* step_base, const float* lut, const float* GH, float* h, bf* h16,
                  cudaStream_t st) {
    const int threads = H >= 256 ? 256 : 64;
    const dim3 grid(cdiv(B, CL_RB), cdiv(H, threads));
#define CL_CASE(NK)                                                                                                            \
    case NK:                                                                                                                   \
        SRNN_LAUNCH_PDL(k_gru_cell_lite<NK>, grid, threads, 0, st, gipre, gipre_ld, g_in_t, k_lo, seq, seq_ld, start_static, \
                        step_base, lut, GH, h, h16, B, H);                                                                     \
        return SRNN_OK;
    switch (k_hi - k_lo) {
        CL_CASE(20)
        CL_CASE(16)
        CL_CASE(4)
        CL_CASE(2)
    }
#undef CL_CASE
    return fail(SRNN_ERR_UNSUPPORTED, "gru_cell_lite: %d sample columns", k_hi - k_lo);
}

bool gru_cell_gen_supported(int H) { return H % (64 * GC_GKB) == 0 && !getenv("SRNN_NO_GRU_CELL"); }

// x16 (B, H) bf16 input of the layer, w_ih16 (3H, H), b_ih (3H), GH (B, 3H) = W_hh h + b_hh; h / h16 updated in place.
int gru_cell_gen(int B, int H, const bf* x16, const bf* w_ih16, const float* b_ih, const float* GH, float* h, bf* h16,
                 cudaStream_t st) {
    CUtensorMap tmX, tmW;
    SRNN_TRY(make_tmap_bf16_kb(&tmX, x16, B, H, H, GP_ROWS, GC_GKB));
    SRNN_TRY(make_tmap_bf16_gates(&tmW, w_ih16, H, H, H, GC_US, GC_GKB));
    const size_t smem = 1024 + (size_t)GC_RING * GC_GROUP_BYTES + GC_NB * 4 + 8 * (2 * GC_RING + 1) + 64;
    static bool attr_set = false;
    if (!attr_set) {
        SRNN_CUDA(cudaFuncSetAttribute(k_gru_cell_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    GruCellParams p{B, H, b_ih, GH, h, h16};
    SRNN_LAUNCH_PDL(k_gru_cell_gen, dim3(H / GC_US, (B + GP_ROWS - 1) / GP_ROWS), dim3(GP_THREADS), smem, st, tmX, tmW, p);
    return SRNN_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
// k-blocks per TMA box: the largest of {SRNN_GRU_GSZ or 4, 2, 1} that divides both the k-block count and the ring depth
static int gp_box_kblocks(int KB) {
    int want = getenv("SRNN_GRU_GSZ") ? atoi(getenv("SRNN_GRU_GSZ")) : 4;
    for (int g = want; g > 1; --g)
        if (KB % g == 0 && GP_STAGES % g == 0) return g;
    return 1;
}

bool gru_persist_supported(int B, int H, int n_sms) {
    if (getenv("SRNN_NO_GRU_PERSIST")) return false;
    if (B < 1 || B > 2 * GP_ROWS || H % 64 || H < 64) return false;
    if ((H / GP_HS) * ((B + GP_ROWS - 1) / GP_ROWS) > n_sms) return false;   // all CTAs must be co-resident
    const size_t w = (size_t)(3 * H / 64) * GP_HS * 128;                   // both kernels keep 96*H bytes of weights
    return gp_smem_bytes(w) <= 227 * 1024;
}

template <typename K, typename... Args>
static int gp_launch(K kernel, int grid, int row_splits, size_t smem, cudaStream_t st, unsigned* ctr, Args... args) {
    SRNN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SRNN_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned) * 2, st));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid, row_splits);
    cfg.blockDim = dim3(GP_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;      // co-residency: the CTAs wait on each other every frame
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail(SRNN_ERR_CUDA, "gru_persist launch: %s", cudaGetErrorString(e));
    return SRNN_OK;
}

// One GRU layer over F frames.  w_hh16 (3H, H) bf16; h0_16 (B, H) bf16 copy of h0; Y16 doubles as the exchange buffer.
int gru_persist_fwd(int B, int F, int H, const float* GI, const bf* w_hh16, const float* b_hh, const float* h0, const bf* h0_16,
                    float* GH, float* Y, bf* Y16, float* h_last, unsigned* ctr, cudaStream_t st) {
    CUtensorMap tmW, tmH0, tmY;
    SRNN_TRY(make_tmap_bf16(&tmW, w_hh16, (uint64_t)3 * H, H, H, GP_HS));
    const int gsz = gp_box_kblocks(H / 64);
    SRNN_TRY(make_tmap_bf16_kb(&tmH0, h0_16, B, H, H, GP_ROWS, gsz));
    SRNN_TRY(make_tmap_bf16_kb(&tmY, Y16, B, (uint64_t)F * H, (uint64_t)F * H, GP_ROWS, gsz));
    GruFwdParams p{B, F, H, getenv("SRNN_GRU_NO_ROTATE") ? 0 : 1, gsz, GI, GH, Y, Y16, h0, h_last, b_hh, ctr};
    const size_t smem = gp_smem_bytes((size_t)(H / 64) * 3 * GP_HS * 128);
    return gp_launch(k_gru_persist_fwd, H / GP_HS, (B + GP_ROWS - 1) / GP_ROWS, smem, st, ctr, tmW, tmH0, tmY, p);
}

// BPTT of one GRU layer.  w_hh16_t (H, 3H) bf16 = W_hh^T.
// dGI / dGH (fp32 copies) may be null; bias_part = scratch of 2*4*H floats when db_ih / db_hh (3H each) are wanted.
int gru_persist_bwd(int B, int F, int H, const float* GI, const float* GH, const float* Y, const float* h0, const float* dY,
                    const bf* w_hh16_t, float* dGI, float* dGH, bf* dGI16, bf* dGH16, float* dh0, unsigned* ctr,
                    cudaStream_t st, float* bias_part, float* db_ih, float* db_hh) {
    CUtensorMap tmWt, tmG;
    SRNN_TRY(make_tmap_bf16(&tmWt, w_hh16_t, H, (uint64_t)3 * H, (uint64_t)3 * H, GP_HS));
    const int gsz = gp_box_kblocks(3 * H / 64);
    SRNN_TRY(make_tmap_bf16_kb(&tmG, dGH16, B, (uint64_t)F * 3 * H, (uint64_t)F * 3 * H, GP_ROWS, gsz));
    GruBwdParams p{B, F, H, getenv("SRNN_GRU_NO_ROTATE") ? 0 : 1, gsz, GI, GH, Y, h0, dY, dGI, dGH, (db_ih || db_hh) ? bias_part : nullptr,
                   dGI16, dGH16, dh0, ctr, nullptr};
    const size_t smem = gp_smem_bytes((size_t)(3 * H / 64) * GP_HS * 128);
    if (getenv("SRNN_TRACE_GRU")) SRNN_CUDA(cudaMalloc((void**)&p.trace, sizeof(long long) * 8 * F));
    int rc = gp_launch(k_gru_persist_bwd, H / GP_HS, (B + GP_ROWS - 1) / GP_ROWS, smem, st, ctr, tmWt, tmG, p);
    if (p.trace) {      // average phase durations (SM cycles) of CTA (0,0) over the frames of this launch
        std::vector<long long> h(8 * F);
        cudaStreamSynchronize(st);
        cudaMemcpy(h.data(), p.trace, sizeof(long long) * 8 * F, cudaMemcpyDeviceToHost);
        cudaFree(p.trace);
        double ew = 0, bar = 0, mm = 0, ep = 0, tot = 0;
        for (int s = 1; s < F; ++s) {
            ew += (double)(h[s * 8 + 1] - h[s * 8 + 0]);
            bar += (double)(h[s * 8 + 2] - h[s * 8 + 1]);
            mm += (double)(h[s * 8 + 3] - h[s * 8 + 2]);
            ep += (double)(h[s * 8 + 4] - h[s * 8 + 3]);
            tot += (double)(h[s * 8 + 0] - h[(s - 1) * 8 + 0]);
        }
        const double n = F > 1 ? F - 1 : 1;
        fprintf(stderr, "[gru bwd trace F=%d] gates+stores=%.0f barrier(all CTAs published)=%.0f stream+MMA=%.0f tmem epilogue=%.0f | frame=%.0f cycles\n",
                F, ew / n, bar / n, mm / n, ep / n, tot / n);
    }
    if (rc == SRNN_OK && p.bias_part)
        SRNN_LAUNCH(k_gru_bias_grads, cdiv(H, 128), 128, 0, st, bias_part, (B + GP_ROWS - 1) / GP_ROWS, H, db_ih, db_hh);
    return rc;
}

}  // namespace srnn
