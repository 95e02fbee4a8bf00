// Persistent GRU recurrence kernels for the teacher-forced (training) path in SRNN_MODE_BF16: ONE cooperative launch runs all
// F frames of one GRU layer (model.py:244 `self.rnn(input, hidden)` = torch nn.GRU; its autograd for the backward kernel),
// instead of one GEMM launch + one gate launch per frame.
//
// Decomposition: CTA c owns the GP_HS = 16 hidden units u0 = 16c .. u0+15 for ALL utterances (rows).  It keeps the recurrent
// weight rows of those units resident in shared memory for the whole launch (UMMA B operand, loaded once by TMA):
//     forward : W_hh[g*H + u0 .. +16, :]  for the three gates g = r, z, n      (48 x H   bf16 = 96 KB at H = 1024)
//     backward: W_hh^T[u0 .. +16, :]                                           (16 x 3H  bf16 = 96 KB at H = 1024)
// and streams the recurrent activations of the frame (h_{f-1}: rows x H, resp. dGH_f: rows x 3H, bf16) through a TMA ring as
// the UMMA A operand (M = 128 rows on the TMEM lanes).  The accumulator D[row][gate*16 + j] therefore puts r, z, n of one
// (utterance, unit) into the SAME thread: the gate math needs no shuffles and the fp32 state h / dh lives in registers
// across frames.  Between frames the CTAs exchange the new bf16 state through global memory (L2) behind a release/acquire
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
constexpr int GP_STAGES = 6;                  // TMA ring: 6 x (128 rows x 128 B) = 96 KB in flight per SM
constexpr int GP_STAGE_BYTES = 128 * 128;
constexpr uint32_t GP_TMEM_COLS = 256;

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
__device__ __forceinline__ float gp_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

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

struct GruFwdParams {
    int B, F, H;
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
    int B, F, H;
    const float* GI;     // saved forward projections (B*F, 3H)
    const float* GH;
    const float* Y;      // (B*F, H) forward outputs; h_{f-1} = Y[b, f-1] or h0[b]
    const float* h0;     // (B, H)
    const float* dY;     // (B*F, H)  gradient w.r.t. the layer outputs
    float* dGI;          // (B*F, 3H) out: gradient w.r.t. W_ih x + b_ih
    float* dGH;          // (B*F, 3H) out: gradient w.r.t. W_hh h + b_hh
    bf* dGI16;           // bf16 copies (operands of the weight-gradient / input-gradient GEMMs)
    bf* dGH16;           //            (and the recurrent operand of this kernel)
    float* dh0;          // (B, H)    out: dL/dh_{-1}
    unsigned* ctr;
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
                                            uint64_t* empty, const unsigned* ctr, int F, int KB, int K, int NS) {
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
        const CUtensorMap* tm = (FWD && s == 0) ? tmA0 : tmA;
        const int col0 = FWD ? (s == 0 ? 0 : (s - 1) * K) : (F - 1 - s) * K;
        for (int kb = 0; kb < KB; ++kb, ++it) {
            const int st = it % GP_STAGES;
            const uint32_t ph = (it / GP_STAGES) & 1;
            mbar_wait(&empty[st], ph ^ 1);
            mbar_expect_tx(&full[st], GP_STAGE_BYTES);
            tma_load_2d(sRing + (size_t)st * GP_STAGE_BYTES, tm, &full[st], col0 + kb * 64, 0);
        }
    }
}

template <int NB>
__device__ __forceinline__ void gp_issuer(int w, int nissue, uint8_t* sW, uint8_t* sRing, uint64_t* full, uint64_t* empty,
                                          uint64_t* w_ready, uint64_t* bar_d, uint32_t tmem, int F, int KB) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, NB);
    mbar_wait(w_ready, 0);
    const uint64_t dW0 = umma_desc_sw128(smem_u32(sW));
    const uint64_t dA0 = umma_desc_sw128(smem_u32(sRing));
    const uint32_t dacc = tmem + (uint32_t)w * 64;
    for (int s = 0; s < F; ++s) {
        for (int kb = w; kb < KB; kb += GP_ISSUERS) {
            const int it = s * KB + kb;
            const int st = it % GP_STAGES;
            const uint32_t ph = (it / GP_STAGES) & 1;
            mbar_wait(&full[st], ph);
            tc_fence_after();
            const uint64_t da = dA0 + (uint64_t)(st * (GP_STAGE_BYTES >> 4));
            const uint64_t db = dW0 + (uint64_t)(kb * (GpSmem<NB>::W_KB_BYTES >> 4));
            umma_bf16(dacc, da, db, idesc, kb >= GP_ISSUERS);
            umma_bf16(dacc, da + 2, db + 2, idesc, 1);
            umma_bf16(dacc, da + 4, db + 4, idesc, 1);
            umma_bf16(dacc, da + 6, db + 6, idesc, 1);
            umma_commit(&empty[st]);
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
    const int H = p.H, F = p.F, KB = H >> 6, NS = gridDim.x;
    const int nissue = KB < GP_ISSUERS ? KB : GP_ISSUERS;
    const int u0 = blockIdx.x * GP_HS;
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
            mbar_init(&L.empty[s], 1);
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
            gp_producer<NB, true>(&tmH0, &tmY, L.sRing, L.full, L.empty, p.ctr, F, KB, H, NS);
        }
    } else if (warp >= 5) {
        if (lane == 0 && warp - 5 < nissue)
            gp_issuer<NB>(warp - 5, nissue, L.sW, L.sRing, L.full, L.empty, L.w_ready, L.bar_d, tmem, F, KB);
    } else {
        // ===================== row warps: gates, state, stores =====================
        const int b = threadIdx.x;                           // utterance row = TMEM lane
        const bool ok = b < p.B;
        const uint32_t tlane = tmem + ((uint32_t)(32 * warp) << 16);
        float h[16];
        if (ok) ld16(p.h0 + (size_t)b * H + u0, h);
        else {
#pragma unroll
            for (int j = 0; j < 16; ++j) h[j] = 0.f;
        }
        const size_t row_stride = (size_t)3 * H;
        for (int f = 0; f < F; ++f) {
            const size_t r = (size_t)b * F + f;
            float gi[3][16];
            if (ok) {
                const float* gp = p.GI + r * row_stride + u0;
                ld16(gp, gi[0]);
                ld16(gp + H, gi[1]);
                ld16(gp + 2 * H, gi[2]);
                if (f + 1 < F) {                             // next frame's projections -> L2 while this frame computes
                    gp_prefetch_l2(gp + row_stride);
                    gp_prefetch_l2(gp + row_stride + H);
                    gp_prefetch_l2(gp + row_stride + 2 * H);
                }
            }
            mbar_wait(L.bar_d, f & 1);
            tc_fence_after();
            float a[3][16];
#pragma unroll
            for (int g = 0; g < 3; ++g) tmem_ld16(tlane + 16 * g, a[g]);
            for (int w = 1; w < nissue; ++w) {
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    float t[16];
                    tmem_ld16(tlane + 64 * w + 16 * g, t);
#pragma unroll
                    for (int j = 0; j < 16; ++j) a[g][j] += t[j];
                }
            }
            tc_fence_before();
            if (ok) {
                float hn[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    a[0][j] += L.sBias[j];
                    a[1][j] += L.sBias[16 + j];
                    a[2][j] += L.sBias[32 + j];
                    const float rr = gp_sigmoid(gi[0][j] + a[0][j]);
                    const float zz = gp_sigmoid(gi[1][j] + a[1][j]);
                    const float nn = tanhf(gi[2][j] + rr * a[2][j]);
                    hn[j] = (1.f - zz) * nn + zz * h[j];
                    h[j] = hn[j];
                }
                st16_bf16(p.Y16 + r * H + u0, hn);           // first: this is what the other CTAs wait for
                st16(p.Y + r * H + u0, hn);
                float* ghp = p.GH + r * row_stride + u0;
                st16(ghp, a[0]);
                st16(ghp + H, a[1]);
                st16(ghp + 2 * H, a[2]);
            }
            gp_bar_sync(1, 128);
            if (threadIdx.x == 0) gp_red_release(p.ctr);     // publishes the whole CTA's h_f slice
        }
        if (ok && p.h_last) st16(p.h_last + (size_t)b * H + u0, h);
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
    const int H = p.H, F = p.F, K3 = 3 * H, KB = K3 >> 6, NS = gridDim.x;
    const int nissue = KB < GP_ISSUERS ? KB : GP_ISSUERS;
    const int u0 = blockIdx.x * GP_HS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    extern __shared__ uint8_t smem_raw[];
    const GpLayout L = gp_layout(smem_raw, (size_t)KB * GpSmem<NB>::W_KB_BYTES);

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmWt);
        prefetch_tmap(&tmG);
        mbar_init(L.w_ready, 1);
        for (int s = 0; s < GP_STAGES; ++s) {
            mbar_init(&L.full[s], 1);
            mbar_init(&L.empty[s], 1);
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
            gp_producer<NB, false>(&tmG, &tmG, L.sRing, L.full, L.empty, p.ctr, F, KB, K3, NS);
        }
    } else if (warp >= 5) {
        if (lane == 0 && warp - 5 < nissue)
            gp_issuer<NB>(warp - 5, nissue, L.sW, L.sRing, L.full, L.empty, L.w_ready, L.bar_d, tmem, F, KB);
    } else {
        const int b = threadIdx.x;
        const bool ok = b < p.B;
        const uint32_t tlane = tmem + ((uint32_t)(32 * warp) << 16);
        float carry[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) carry[j] = 0.f;
        const size_t row_stride = (size_t)K3;
        for (int s = 0; s < F; ++s) {
            const int f = F - 1 - s;
            const size_t r = (size_t)b * F + f;
            float dhz[16];
            if (ok) {
                float gr[16], gz[16], gn[16], hr[16], hz[16], hn[16], hp[16], dy[16];
                const float* gip = p.GI + r * row_stride + u0;
                const float* ghp = p.GH + r * row_stride + u0;
                ld16(gip, gr);
                ld16(gip + H, gz);
                ld16(gip + 2 * H, gn);
                ld16(ghp, hr);
                ld16(ghp + H, hz);
                ld16(ghp + 2 * H, hn);
                ld16(f ? p.Y + (r - 1) * H + u0 : p.h0 + (size_t)b * H + u0, hp);
                ld16(p.dY + r * H + u0, dy);
                if (f > 0) {                                 // the previous frame's rows -> L2 while this frame's GEMM runs
                    gp_prefetch_l2(gip - row_stride);
                    gp_prefetch_l2(gip - row_stride + H);
                    gp_prefetch_l2(gip - row_stride + 2 * H);
                    gp_prefetch_l2(ghp - row_stride);
                    gp_prefetch_l2(ghp - row_stride + H);
                    gp_prefetch_l2(ghp - row_stride + 2 * H);
                    gp_prefetch_l2(p.dY + (r - 1) * H + u0);
                    if (f > 1) gp_prefetch_l2(p.Y + (r - 2) * H + u0);
                }
                float o_r[16], o_z[16], o_n[16], o_nr[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float rr = gp_sigmoid(gr[j] + hr[j]);
                    const float zz = gp_sigmoid(gz[j] + hz[j]);
                    const float nn = tanhf(gn[j] + rr * hn[j]);
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
                }
                bf* g16 = p.dGH16 + r * row_stride + u0;     // first: the recurrent operand the other CTAs wait for
                st16_bf16(g16, o_r);
                st16_bf16(g16 + H, o_z);
                st16_bf16(g16 + 2 * H, o_nr);
                float* gh = p.dGH + r * row_stride + u0;
                st16(gh, o_r);
                st16(gh + H, o_z);
                st16(gh + 2 * H, o_nr);
                float* gi = p.dGI + r * row_stride + u0;
                st16(gi, o_r);
                st16(gi + H, o_z);
                st16(gi + 2 * H, o_n);
                bf* i16 = p.dGI16 + r * row_stride + u0;
                st16_bf16(i16, o_r);
                st16_bf16(i16 + H, o_z);
                st16_bf16(i16 + 2 * H, o_n);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) dhz[j] = 0.f;
            }
            gp_bar_sync(1, 128);
            if (threadIdx.x == 0) gp_red_release(p.ctr);     // dGH_f of this CTA's units is published
            mbar_wait(L.bar_d, s & 1);
            tc_fence_after();
            float a[16];
            tmem_ld16(tlane, a);
            for (int w = 1; w < nissue; ++w) {
                float t[16];
                tmem_ld16(tlane + 64 * w, t);
#pragma unroll
                for (int j = 0; j < 16; ++j) a[j] += t[j];
            }
            tc_fence_before();
#pragma unroll
            for (int j = 0; j < 16; ++j) carry[j] = dhz[j] + a[j];
        }
        if (ok && p.dh0) st16(p.dh0 + (size_t)b * H + u0, carry);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc<GP_TMEM_COLS>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
bool gru_persist_supported(int B, int H, int n_sms) {
    if (getenv("SRNN_NO_GRU_PERSIST")) return false;
    if (B < 1 || B > 128 || H % 64 || H < 64) return false;
    if (H / GP_HS > n_sms) return false;                                   // all CTAs must be co-resident
    const size_t w = (size_t)(3 * H / 64) * GP_HS * 128;                   // both kernels keep 96*H bytes of weights
    return gp_smem_bytes(w) <= 227 * 1024;
}

template <typename K, typename... Args>
static int gp_launch(K kernel, int grid, size_t smem, cudaStream_t st, unsigned* ctr, Args... args) {
    SRNN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SRNN_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned), st));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
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
    SRNN_TRY(make_tmap_bf16(&tmH0, h0_16, B, H, H, 128));
    SRNN_TRY(make_tmap_bf16(&tmY, Y16, B, (uint64_t)F * H, (uint64_t)F * H, 128));
    GruFwdParams p{B, F, H, GI, GH, Y, Y16, h0, h_last, b_hh, ctr};
    const size_t smem = gp_smem_bytes((size_t)(H / 64) * 3 * GP_HS * 128);
    return gp_launch(k_gru_persist_fwd, H / GP_HS, smem, st, ctr, tmW, tmH0, tmY, p);
}

// BPTT of one GRU layer.  w_hh16_t (H, 3H) bf16 = W_hh^T.
int gru_persist_bwd(int B, int F, int H, const float* GI, const float* GH, const float* Y, const float* h0, const float* dY,
                    const bf* w_hh16_t, float* dGI, float* dGH, bf* dGI16, bf* dGH16, float* dh0, unsigned* ctr,
                    cudaStream_t st) {
    CUtensorMap tmWt, tmG;
    SRNN_TRY(make_tmap_bf16(&tmWt, w_hh16_t, H, (uint64_t)3 * H, (uint64_t)3 * H, GP_HS));
    SRNN_TRY(make_tmap_bf16(&tmG, dGH16, B, (uint64_t)F * 3 * H, (uint64_t)F * 3 * H, 128));
    GruBwdParams p{B, F, H, GI, GH, Y, h0, dY, dGI, dGH, dGI16, dGH16, dh0, ctr};
    const size_t smem = gp_smem_bytes((size_t)(3 * H / 64) * GP_HS * 128);
    return gp_launch(k_gru_persist_bwd, H / GP_HS, smem, st, ctr, tmWt, tmG, p);
}

}  // namespace srnn
