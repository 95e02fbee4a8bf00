// tcgen05 / TMA GEMM for the weight-streaming contractions of the frame tiers (GRU input + recurrent projections,
// learned upsampling) in SRNN_MODE_BF16.
//
// "Swap-AB" orientation: the WEIGHT matrix (features x K, bf16, K-major) is the UMMA A operand (M = 128 or 64
// features per CTA) and the ACTIVATIONS (rows x K, bf16, K-major) are the B operand (N = BN batch rows), so the
// accumulator in TMEM is D[feature lane][batch-row column] and the big streamed operand (the weights) is read once
// per row tile.  Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..9 = epilogue (tcgen05.ld -> bias/ReLU -> coalesced stores of out[row][feature]); two warps share each
// TMEM lane quadrant and split the batch-row columns between them (the epilogue, not the MMA, bounds small-K tiles).
#include "common.cuh"
#include "umma.cuh"
#include <stdlib.h>

namespace srnn {

using namespace ptx;

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_tmapEncodeTiled)p;
    }
    return fn;
}

// bf16 row-major (rows x cols, leading dimension ld elements) -> 2-D tensor map with a {64, box_rows} box, 128B swizzle
int make_tmap_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    PFN_tmapEncodeTiled enc = get_encode();
    if (!enc) return fail(SRNN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (((uintptr_t)base & 15) || (ld * 2) % 16) return fail(SRNN_ERR_ARG, "tensor map: base/stride must be 16-byte aligned");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * sizeof(__nv_bfloat16)};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SRNN_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SRNN_OK;
}

// Same matrix viewed as [k-block][row][64]: ONE TMA instruction fetches box_kb consecutive k-blocks of box_rows rows and lays
// them out as box_kb consecutive {box_rows x 128 B} swizzled stages (issuing a TMA costs ~80 cycles of the producer thread).
int make_tmap_bf16_kb(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                      uint32_t box_kb) {
    PFN_tmapEncodeTiled enc = get_encode();
    if (!enc) return fail(SRNN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (((uintptr_t)base & 15) || (ld * 2) % 16 || cols % 64) return fail(SRNN_ERR_ARG, "tensor map: alignment");
    cuuint64_t gdim[3] = {64, rows, cols / 64};
    cuuint64_t gstride[2] = {ld * sizeof(__nv_bfloat16), 128};
    cuuint32_t box[3] = {64, box_rows, box_kb};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SRNN_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r);
    return SRNN_OK;
}

// GRU weight matrix (3H, K) viewed as [k-block][gate][unit][64]: one TMA box = box_kb k-blocks x 3 gates x box_units units,
// landing as box_kb consecutive {3*box_units rows x 128 B} swizzled tiles (rows ordered gate-major): the B operand of a
// fused gate GEMM for one slice of hidden units (gru_persist.cu, k_gru_cell_gen).
int make_tmap_bf16_gates(CUtensorMap* tm, const void* base, uint64_t H, uint64_t K, uint64_t ld, uint32_t box_units,
                         uint32_t box_kb) {
    PFN_tmapEncodeTiled enc = get_encode();
    if (!enc) return fail(SRNN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (((uintptr_t)base & 15) || (ld * 2) % 16 || K % 64) return fail(SRNN_ERR_ARG, "tensor map: alignment");
    cuuint64_t gdim[4] = {64, H, 3, K / 64};
    cuuint64_t gstride[3] = {ld * sizeof(__nv_bfloat16), H * ld * sizeof(__nv_bfloat16), 128};
    cuuint32_t box[4] = {64, box_units, 3, box_kb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SRNN_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed (%d)", (int)r);
    return SRNN_OK;
}

// fp32 row-major output (rows x cols, leading dimension ld) -> 2-D map with a {32 floats = 128 B, box_rows} box, 128B swizzle:
// the TMA-store epilogue of k_gemm_umma_pair_wide
int make_tmap_f32_out(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    PFN_tmapEncodeTiled enc = get_encode();
    if (!enc) return fail(SRNN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (((uintptr_t)base & 15) || (ld * 4) % 16) return fail(SRNN_ERR_ARG, "tensor map: base/stride must be 16-byte aligned");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SRNN_ERR_CUDA, "cuTensorMapEncodeTiled (fp32 out) failed (%d)", (int)r);
    return SRNN_OK;
}

constexpr int GEMM_THREADS = 320;   // 2 control warps + 8 epilogue warps

template <int BM, int BN>
struct GemmSmem {
    static constexpr int A_BYTES = BM * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE = A_BYTES + B_BYTES;
    // as many TMA stages as fit in ~192 KB (max 8): small-N tiles are latency-bound on the load pipeline, not on the MMA
    static constexpr int NSTAGE = (196608 / STAGE) < 8 ? (196608 / STAGE) : 8;
    static constexpr int TOTAL = NSTAGE * STAGE + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

// one launch can carry two independent problems that share the activation row count and K (blockIdx.z selects):
// the GRU input projection gi = W_ih x and the recurrent projection gh = W_hh h of one layer run side by side.
struct alignas(64) GemmProb {
    CUtensorMap tmA, tmB;
    const float* bias;
    const float* addend;
    float* out_f32;
    __nv_bfloat16* out_bf16;
    const __nv_bfloat16* mask;     // optional (rows, ld_out) bf16: the value is kept where mask > 0, else 0 (ReLU backward)
    int n_feat, ld_add, ld_out, relu;
};
struct alignas(64) GemmArgs {
    GemmProb p[2];
    int n_rows, K;
    int gx, gy, gz;        // tile grid (A-operand tiles, B-operand tiles, problems or K splits); CTAs stride over it
    int nbuf;              // TMEM accumulator buffers: 2 = epilogue of tile i overlaps the main loop of tile i+1
    int nstage;            // TMA ring depth in use (<= GemmSmem::NSTAGE): a shallow ring lets two CTAs share an SM
    int ksplit;            // > 1: single problem, blockIdx.z = K split; split z writes out_f32 + z * split_stride (partials)
    long long split_stride;
    long long* trace;      // development aid (SRNN_TRACE_GEMM=1): clock64 stamps of CTA (0,0,0), else null
};

// 16 accumulator columns (= 16 consecutive batch rows) of one feature -> global memory.  The flag combination is a
// compile-time parameter: a branchy per-element epilogue (~50 SASS instructions per value) was measured to cost more
// than the whole K loop of a 128x256 tile.
template <bool ADD, bool RELU, bool F32, bool B16, bool MASK = false>
__device__ __forceinline__ void epi_store16(const float (&v)[16], float bv, int nn, size_t row, int m,
                                            const float* __restrict__ addend, int ld_add, float* __restrict__ out_f32,
                                            __nv_bfloat16* __restrict__ out_bf16, int ld_out,
                                            const __nv_bfloat16* __restrict__ mask = nullptr) {
    float* pf = F32 ? out_f32 + row * ld_out + m : nullptr;
    __nv_bfloat16* pb = B16 ? out_bf16 + row * ld_out + m : nullptr;
    const float* pa = ADD ? addend + row * ld_add + m : nullptr;
    const __nv_bfloat16* pm = MASK ? mask + row * ld_out + m : nullptr;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (nn == 16 || i < nn) {
            float x = v[i] + bv;
            if (ADD) x += pa[(size_t)i * ld_add];
            if (RELU) x = fmaxf(x, 0.f);
            if (MASK) x = __bfloat162float(pm[(size_t)i * ld_out]) > 0.f ? x : 0.f;
            if (F32) pf[(size_t)i * ld_out] = x;
            if (B16) pb[(size_t)i * ld_out] = __float2bfloat16(x);
        }
    }
}

// ROWS orientation (activation rows on the TMEM lanes, features on the columns): one thread owns 16 CONSECUTIVE features of
// one output row, so bias / addend / mask loads and the output stores are 16-byte vectors.
template <bool ADD, bool RELU, bool F32, bool B16, bool MASK>
__device__ __forceinline__ void epi_row16(const float (&v)[16], const float* __restrict__ bias, const float* __restrict__ pa,
                                          float* __restrict__ pf, __nv_bfloat16* __restrict__ pb,
                                          const __nv_bfloat16* __restrict__ pm) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = v[i];
    if (bias) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + i);
            x[4 * i] += b4.x; x[4 * i + 1] += b4.y; x[4 * i + 2] += b4.z; x[4 * i + 3] += b4.w;
        }
    }
    if (ADD) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 a4 = reinterpret_cast<const float4*>(pa)[i];
            x[4 * i] += a4.x; x[4 * i + 1] += a4.y; x[4 * i + 2] += a4.z; x[4 * i + 3] += a4.w;
        }
    }
    if (RELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaxf(x[i], 0.f);
    }
    if (MASK) {
        const uint4 m0 = reinterpret_cast<const uint4*>(pm)[0], m1 = reinterpret_cast<const uint4*>(pm)[1];
        const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {     // bf16 > 0  <=>  sign bit clear and magnitude non-zero (NaN never occurs: ReLU output)
            const uint32_t lo = mw[i] & 0xffffu, hi = mw[i] >> 16;
            if (!(lo != 0 && lo < 0x8000u)) x[2 * i] = 0.f;
            if (!(hi != 0 && hi < 0x8000u)) x[2 * i + 1] = 0.f;
        }
    }
    if (F32) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(pf)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
    }
    if (B16) {
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
            o[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        reinterpret_cast<uint4*>(pb)[0] = make_uint4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<uint4*>(pb)[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// ROWS = false: "swap-AB" (weights = A operand on the TMEM lanes; the generation-time / per-frame orientation).
// ROWS = true : activations = A operand (128 rows per CTA), weights = B operand (BN features): the orientation of the
//               big teacher-forced GEMMs, whose epilogue then stores 16 consecutive features per thread.
// MNMAJ = true (ROWS only): both operands are given TRANSPOSED in memory -- tmA over X (K_total x M_total) and tmB over
//               Y (K_total x N_total), row-major -- and out = X^T . Y: the weight-gradient GEMMs dW = dOut^T . In read
//               dOut and In as they are, no transposed copies.
// Epilogue of one 128-row x BN-feature accumulator tile in the ROWS orientation (lane = output row, 16 consecutive features
// per tcgen05.ld): shared by the single-CTA and the CTA-pair kernels.
template <int BN>
__device__ __forceinline__ void rows_epilogue(const GemmProb& P, float* __restrict__ out_f32, uint32_t tmem_d, int m0, int n0,
                                              int n_rows, int q, int half, int lane) {
    const int n_feat = P.n_feat;
    const float* __restrict__ bias = P.bias;
    const float* __restrict__ addend = P.addend;
    __nv_bfloat16* __restrict__ out_bf16 = P.out_bf16;
    const __nv_bfloat16* __restrict__ mask = P.mask;
    const int ld_out = P.ld_out, ld_add = P.ld_add, relu = P.relu;
    const int r = m0 + 32 * q + lane;
    const bool r_ok = r < n_rows;
    float* __restrict__ of = out_f32;
#pragma unroll 1
    for (int c = 16 * half; c < BN; c += 32) {
        const int f0 = n0 + c;
        if (f0 >= n_feat) break;               // warp-uniform
        float v[16];
        tmem_ld16(tmem_d + ((uint32_t)(32 * q) << 16) + c, v);
        if (!r_ok) continue;
        const size_t o = (size_t)r * ld_out + f0;
        const float* bp = bias ? bias + f0 : nullptr;
        if (f0 + 16 <= n_feat) {
            if (mask)
                epi_row16<false, false, false, true, true>(v, bp, nullptr, nullptr, out_bf16 + o, mask + o);
            else if (!addend && !relu && of && !out_bf16)
                epi_row16<false, false, true, false, false>(v, bp, nullptr, of + o, nullptr, nullptr);
            else if (!addend && relu && !of && out_bf16)
                epi_row16<false, true, false, true, false>(v, bp, nullptr, nullptr, out_bf16 + o, nullptr);
            else if (addend && !relu && of && !out_bf16)
                epi_row16<true, false, true, false, false>(v, bp, addend + (size_t)r * ld_add + f0, of + o, nullptr, nullptr);
            else if (!addend && !relu && of && out_bf16)
                epi_row16<false, false, true, true, false>(v, bp, nullptr, of + o, out_bf16 + o, nullptr);
            else if (!addend && !relu && !of && out_bf16)
                epi_row16<false, false, false, true, false>(v, bp, nullptr, nullptr, out_bf16 + o, nullptr);
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float x = v[i] + (bp ? bp[i] : 0.f);
                    if (addend) x += addend[(size_t)r * ld_add + f0 + i];
                    if (relu) x = fmaxf(x, 0.f);
                    if (of) of[o + i] = x;
                    if (out_bf16) out_bf16[o + i] = __float2bfloat16(x);
                }
            }
        } else {                               // ragged feature tail
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (f0 + i < n_feat) {
                    float x = v[i] + (bp ? bp[i] : 0.f);
                    if (addend) x += addend[(size_t)r * ld_add + f0 + i];
                    if (relu) x = fmaxf(x, 0.f);
                    if (mask) x = __bfloat162float(mask[o + i]) > 0.f ? x : 0.f;
                    if (of) of[o + i] = x;
                    if (out_bf16) out_bf16[o + i] = __float2bfloat16(x);
                }
            }
        }
    }
}

// Epilogue of one BM-feature x BN-row accumulator tile in the swap-AB orientation (lane = feature, columns = batch rows).
template <int BM, int BN>
__device__ __forceinline__ void swap_epilogue(const GemmProb& P, float* __restrict__ out_f32, uint32_t tmem_d, int m0, int n0,
                                              int n_rows, int q, int half, int lane) {
    const int n_feat = P.n_feat;
    const float* __restrict__ bias = P.bias;
    const float* __restrict__ addend = P.addend;
    __nv_bfloat16* __restrict__ out_bf16 = P.out_bf16;
    const __nv_bfloat16* __restrict__ mask = P.mask;
    const int ld_out = P.ld_out, ld_add = P.ld_add, relu = P.relu;
    int m;
    bool lane_ok;
    if (BM == 128) {
        m = m0 + 32 * q + lane;
        lane_ok = true;
    } else {                                   // M = 64: rows 16q..16q+15 live in lanes 0..15 of quadrant q
        m = m0 + 16 * q + lane;
        lane_ok = lane < 16;
    }
    const bool m_ok = lane_ok && m < n_feat;
    const float bv = (bias && m_ok) ? bias[m] : 0.f;
#pragma unroll 1
    for (int c = 16 * half; c < BN; c += 32) {
        if (n0 + c >= n_rows) break;           // warp-uniform: nothing but padding rows beyond this point
        float v[16];
        tmem_ld16(tmem_d + ((uint32_t)(32 * q) << 16) + c, v);
        if (!m_ok) continue;
        const int nn = n_rows - (n0 + c) < 16 ? n_rows - (n0 + c) : 16;
        const size_t row = (size_t)(n0 + c);
        if (mask)                                              // ReLU backward: dpre = dx where the forward value > 0
            epi_store16<false, false, false, true, true>(v, bv, nn, row, m, addend, ld_add, out_f32, out_bf16, ld_out, mask);
        else if (!addend && !relu && out_f32 && !out_bf16)     // GRU projections, upsampling, logits, weight gradients
            epi_store16<false, false, true, false>(v, bv, nn, row, m, addend, ld_add, out_f32, out_bf16, ld_out);
        else if (!addend && relu && !out_f32 && out_bf16)      // MLP hidden layer feeding the next GEMM
            epi_store16<false, true, false, true>(v, bv, nn, row, m, addend, ld_add, out_f32, out_bf16, ld_out);
        else if (addend && !relu && out_f32 && !out_bf16)      // BPTT carry: dh*z + dGH.W_hh
            epi_store16<true, false, true, false>(v, bv, nn, row, m, addend, ld_add, out_f32, out_bf16, ld_out);
        else if (!addend && !relu && !out_f32 && out_bf16)     // generation: top-tier upsampling feeding tier 0's folded first layer
            epi_store16<false, false, false, true>(v, bv, nn, row, m, addend, ld_add, out_f32, out_bf16, ld_out);
        else {                                                 // generic (test hook)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float x = v[i] + bv;
                if (i < nn) {
                    if (addend) x += addend[(row + i) * ld_add + m];
                    if (relu) x = fmaxf(x, 0.f);
                    if (out_f32) out_f32[(row + i) * ld_out + m] = x;
                    if (out_bf16) out_bf16[(row + i) * ld_out + m] = __float2bfloat16(x);
                }
            }
        }
    }
}

struct TileInfo {
    int m0, n0, kb0, KB, prob, split;
    bool valid;
};
// tile index -> (A-operand tile, B-operand tile, problem / K split); the same enumeration as the former 3-D grid
template <int BM, int BN, bool ROWS>
__device__ __forceinline__ TileInfo tile_info(const GemmArgs& args, int tile) {
    TileInfo t;
    const int bx = tile % args.gx, by = (tile / args.gx) % args.gy, bz = tile / (args.gx * args.gy);
    t.m0 = bx * BM;
    t.n0 = by * BN;
    t.prob = args.ksplit > 1 ? 0 : bz;
    t.split = args.ksplit > 1 ? bz : 0;
    t.kb0 = 0;
    t.KB = args.K / 64;
    if (args.ksplit > 1) {                         // this tile's K range (every split is non-empty by construction)
        const int per = (t.KB + args.ksplit - 1) / args.ksplit;
        t.kb0 = bz * per;
        t.KB = t.KB - t.kb0 < per ? t.KB - t.kb0 : per;
    }
    t.valid = t.m0 < (ROWS ? args.n_rows : args.p[t.prob].n_feat);   // the two problems of a launch may differ in feature count
    return t;
}

template <int BM, int BN, bool ROWS, bool MNMAJ = false>
__global__ void __launch_bounds__(GEMM_THREADS, BN <= 256 ? 2 : 1)
k_gemm_umma(const __grid_constant__ GemmArgs args) {
    // PERSISTENT: CTA b processes tiles b, b + gridDim.x, ...  With args.nbuf == 2 the accumulator is double-buffered in
    // TMEM, so the epilogue of tile i (TMEM -> registers -> global) overlaps the TMA/MMA main loop of tile i+1; the
    // producer / MMA / epilogue roles only meet at mbarriers (full/empty per smem stage, full/empty per TMEM buffer).
    using S = GemmSmem<BM, BN>;
    constexpr uint32_t TCOLS = BN < 32 ? 32 : BN;
    const int n_rows = args.n_rows, nbuf = args.nbuf, ntiles = args.gx * args.gy * args.gz;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int UMMA_STAGES = args.nstage;
    uint64_t* full = (uint64_t*)(smem + UMMA_STAGES * S::STAGE);
    uint64_t* empty = full + UMMA_STAGES;
    uint64_t* tmem_full = empty + UMMA_STAGES;     // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long* tr = (args.trace && blockIdx.x == 0) ? args.trace : nullptr;
    if (tr && threadIdx.x == 0) tr[0] = clock64();

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&args.p[0].tmA);
        prefetch_tmap(&args.p[0].tmB);
        for (int s = 0; s < UMMA_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], GEMM_THREADS / 32 - 2);     // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    pdl_trigger();
    if (warp == 1) tmem_alloc_rt(tmem_slot, nbuf == 2 ? 2 * TCOLS : TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform (keeps UMMA operands in uniform regs)
    pdl_wait();                                    // programmatic dependent launch: operands come from the preceding kernel

    if (warp == 0) {
        if (lane == 0) {
            int s = 0, first = 1;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const TileInfo ti = tile_info<BM, BN, ROWS>(args, tile);
                if (!ti.valid) continue;
                const GemmProb& P = args.p[ti.prob];
                const int m0 = ti.m0, n0 = ti.n0, kb0 = ti.kb0;
                for (int kb = 0; kb < ti.KB; ++kb) {
                    if (!first) { if (++s == UMMA_STAGES) { s = 0; ph ^= 1; } }
                    first = 0;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (tr && tile == 0 && kb < 24) tr[8 + kb] = clock64();
                    mbar_expect_tx(&full[s], S::STAGE);
                    if constexpr (MNMAJ) {     // boxes of {64 MN elements, 64 K rows}: one per 64-wide MN chunk
#pragma unroll
                        for (int c = 0; c < BM / 64; ++c)
                            tma_load_2d(smem + s * S::STAGE + c * 8192, &P.tmA, &full[s], m0 + c * 64, (kb0 + kb) * 64);
#pragma unroll
                        for (int c = 0; c < BN / 64; ++c)
                            tma_load_2d(smem + s * S::STAGE + S::A_BYTES + c * 8192, &P.tmB, &full[s], n0 + c * 64, (kb0 + kb) * 64);
                    } else {
                        tma_load_2d(smem + s * S::STAGE, &P.tmA, &full[s], (kb0 + kb) * 64, m0);
                        tma_load_2d(smem + s * S::STAGE + S::A_BYTES, &P.tmB, &full[s], (kb0 + kb) * 64, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = MNMAJ ? umma_idesc_bf16_mn(BM, BN) : umma_idesc_bf16(BM, BN);
            const uint64_t d0 = MNMAJ ? umma_desc_sw128_mn(smem_u32(smem), 8192) : umma_desc_sw128(smem_u32(smem));
            constexpr int KSTEP = MNMAJ ? 128 : 2;   // descriptor advance per K = 16: 16 K rows x 128 B, or 32 B inside the row
            int s = 0, first = 1, tl = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const TileInfo ti = tile_info<BM, BN, ROWS>(args, tile);
                if (!ti.valid) continue;
                const int buf = nbuf == 2 ? (tl & 1) : 0, use = nbuf == 2 ? (tl >> 1) : tl;
                mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);        // the epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)buf * TCOLS;
                for (int kb = 0; kb < ti.KB; ++kb) {
                    if (!first) { if (++s == UMMA_STAGES) { s = 0; ph ^= 1; } }
                    first = 0;
                    mbar_wait(&full[s], ph);
                    if (tr && tile == 0 && kb < 24) tr[32 + kb] = clock64();
                    tc_fence_after();
                    const uint64_t da = d0 + (uint64_t)(s * (S::STAGE >> 4));
                    const uint64_t db = da + (uint64_t)(S::A_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)       // 4 x (K = 16 bf16 = 32 B) inside the 128-byte swizzle row
                        umma_bf16(tmem_d, da + KSTEP * k, db + KSTEP * k, idesc, (kb | k) != 0);
                    umma_commit(&empty[s]);            // frees the smem stage when these MMAs retire
                }
                umma_commit(&tmem_full[buf]);
                ++tl;
            }
        }
    } else {
        // epilogue: a warp may only touch TMEM lanes 32*(warp%4) .. +31; warps 2..5 take the even 16-column chunks,
        // warps 6..9 the odd ones
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        int tl = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileInfo ti = tile_info<BM, BN, ROWS>(args, tile);
        if (!ti.valid) continue;
        const GemmProb& P = args.p[ti.prob];
        const int n_feat = P.n_feat, m0 = ti.m0, n0 = ti.n0;
        const int buf = nbuf == 2 ? (tl & 1) : 0, use = nbuf == 2 ? (tl >> 1) : tl;
        const uint32_t tmem_d = tmem_base + (uint32_t)buf * TCOLS;
        const float* __restrict__ bias = P.bias;
        const float* __restrict__ addend = P.addend;
        float* __restrict__ out_f32 =
            (P.out_f32 && args.ksplit > 1) ? P.out_f32 + (size_t)ti.split * args.split_stride : P.out_f32;
        __nv_bfloat16* __restrict__ out_bf16 = P.out_bf16;
        const __nv_bfloat16* __restrict__ mask = P.mask;
        const int ld_out = P.ld_out, ld_add = P.ld_add, relu = P.relu;
        mbar_wait(&tmem_full[buf], use & 1);
        if (tr && threadIdx.x == 64 && tile == 0) tr[1] = clock64();
        tc_fence_after();
        if constexpr (ROWS) {
            rows_epilogue<BN>(P, out_f32, tmem_d, m0, n0, n_rows, q, half, lane);
        } else {
            swap_epilogue<BM, BN>(P, out_f32, tmem_d, m0, n0, n_rows, q, half, lane);
        }   // !ROWS
        // this warp has read its part of the accumulator: hand the TMEM buffer back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        if (tr && threadIdx.x == 64 && tile == 0) tr[2] = clock64();
        ++tl;
        }   // tile loop
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc_rt(tmem_base, nbuf == 2 ? 2 * TCOLS : TCOLS);
}

static int g_gemm_sms = -1;
// > 0: the next single-CTA launches use at most this many CTAs (persistent tile loop, deep ring) -- a GEMM that is to run on the
// SMs another, co-resident kernel leaves free (the recurrent projections beside the persistent sample kernel)
static thread_local int g_gemm_cta_cap = 0;
thread_local int g_pdl = 0;
void gemm_umma_set_cta_cap(int cap) { g_gemm_cta_cap = cap; }
static constexpr int TCOLS_OF(int bn) { return bn < 32 ? 32 : bn; }
template <int BM, int BN, bool ROWS, bool MNMAJ = false>
static int launch_gemm_umma(const GemmArgs& args, int nprob, int max_feat, cudaStream_t st) {
    using S = GemmSmem<BM, BN>;
    static bool attr_set = false;
    if (!attr_set) {
        SRNN_CUDA(cudaFuncSetAttribute((k_gemm_umma<BM, BN, ROWS, MNMAJ>), cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        attr_set = true;
    }
    const int nz = args.ksplit > 1 ? args.ksplit : nprob;
    const dim3 grid = ROWS ? dim3(cdiv(args.n_rows, BM), cdiv(max_feat, BN), nz) : dim3(cdiv(max_feat, BM), cdiv(args.n_rows, BN), nz);
    // Ring depth: the deepest that fits one CTA per SM, unless the grid is a little larger than one wave of SMs -- then a
    // shallow ring (<= 110 KB) lets two CTAs share an SM, so the whole grid is resident at once and one CTA's epilogue
    // overlaps the other's loads (the generation-time upsampling GEMM has 160 tiles on 148 SMs).
    GemmArgs a = args;
    a.nstage = S::NSTAGE;
    const long long ctas = (long long)grid.x * grid.y * grid.z;
    const int shallow = (110 * 1024 - 1280) / S::STAGE;
    static const bool shallow_all = getenv("SRNN_GEMM_SHALLOW_ALL") != nullptr;
    const int cap = g_gemm_cta_cap;
    if (!cap && g_gemm_sms > 0 && ctas > g_gemm_sms && (ctas <= 2 * g_gemm_sms || shallow_all) && shallow >= 2 && TCOLS_OF(BN) <= 256 &&
        !getenv("SRNN_GEMM_DEEP_RING"))
        a.nstage = shallow < S::NSTAGE ? shallow : S::NSTAGE;
    const size_t smem = (size_t)a.nstage * S::STAGE + 1024 + 256;
    a.gx = (int)grid.x;
    a.gy = (int)grid.y;
    a.gz = (int)grid.z;
    // Deep ring = one CTA per SM: run persistent (one CTA per SM striding over the tiles) with two accumulator buffers when
    // they fit in TMEM; shallow ring = two CTAs per SM, one tile each, one buffer (512 TMEM columns are shared by both).
    const bool deep = a.nstage == S::NSTAGE;
    a.nbuf = (deep && 2 * TCOLS_OF(BN) <= 512 && ctas > 1 && !getenv("SRNN_GEMM_SINGLE_BUF")) ? 2 : 1;
    long long launch_ctas = ctas;
    if (deep && g_gemm_sms > 0 && ctas > g_gemm_sms) launch_ctas = g_gemm_sms;
    if (cap > 0 && launch_ctas > cap) launch_ctas = cap;
    SRNN_LAUNCH_PDL((k_gemm_umma<BM, BN, ROWS, MNMAJ>), dim3((unsigned)launch_ctas), dim3(GEMM_THREADS), smem, st, a);
    return SRNN_OK;
}

// sum of `splits` partial matrices (fixed order) -> out
__global__ void k_sum_splits(const float* __restrict__ part, int splits, size_t n, size_t stride, float* __restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += part[(size_t)z * stride + i];
        out[i] = s;
    }
}
int sum_splits(const float* part, int splits, size_t n, size_t stride, float* out, cudaStream_t st) {
    int grid = (int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256);
    SRNN_LAUNCH(k_sum_splits, grid < 1 ? 1 : grid, 256, 0, st, part, splits, n, stride, out);
    return SRNN_OK;
}

// nprob (1 or 2) problems  out_i (rows, feat_i) = act_i (rows, K) . W_i (feat_i, K)^T + bias_i [+ addend_i] [relu];
// W / act are bf16 with K % 64 == 0.  bn selects the batch-row tile (32..256); bm = 128 (default) or 64.
// ---------------------------------------------------------------------------------------------------------------------
// CTA-PAIR variant (cta_group::2) of the ROWS / K-major GEMM: the two CTAs of a 2-cluster (one TPC) execute ONE UMMA with
// M = 256 (128 output rows per CTA) x N = 256.  Each CTA streams only its 128 activation rows and HALF of the weight tile
// (128 feature rows) per k-block -- 32 KB instead of 48 KB -- because the tensor core of each SM reads the other half of B
// from its peer's shared memory.  The single-CTA kernel's main loop is TMA-ingest bound (48 KB ~ 1000 cycles per k-block
// against 512 cycles of MMA), so this is worth ~1.4x on the long-M teacher-forced GEMMs.
//   * both CTAs run a TMA producer; the byte counts of BOTH land on the EVEN CTA's full barrier (peer-bit-masked address);
//   * one thread of the even CTA issues tcgen05.mma.cta_group::2 and commits with .multicast::cluster to the empty / tmem_full
//     barriers of both CTAs; the epilogue warps of both CTAs arrive on the even CTA's tmem_empty barrier;
//   * persistent over 256 x 256 pair tiles, accumulators double-buffered in TMEM (2 x 256 columns in each CTA).
// PTX forms follow cute/arch/copy_sm100_tma.hpp, mma_sm100_umma.hpp, cutlass/arch/barrier.h of the vendored CUTLASS headers.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int PAIR_HALF = 16384;                 // 128 rows x 128 B
constexpr int PAIR_STAGE = 2 * PAIR_HALF;        // A rows + B half per CTA
constexpr int PAIR_NSTAGE = 6;
constexpr int PAIR_SMEM = PAIR_NSTAGE * PAIR_STAGE + 1024 + 256;

// MNMAJ: both operands transposed in memory (weight gradients dW = dOut^T . In), as in k_gemm_umma; gz = K splits.
// SWAP: swap-AB orientation (weights = A operand, 256 features per pair tile on the lanes; activations = B operand, 128 of
// 256 batch rows per CTA): the generation-time upsampling GEMM (256 utterances x 20480 features).
template <bool MNMAJ, bool SWAP = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
k_gemm_umma_pair(const __grid_constant__ GemmArgs args) {
    constexpr uint32_t TCOLS = 256;
    const GemmProb& P = args.p[0];
    const int n_rows = args.n_rows, KBT = args.K / 64, nxy = args.gx * args.gy, ntiles = nxy * args.gz;
    const int kper = (KBT + args.gz - 1) / args.gz;             // k-blocks per split (every split non-empty by construction)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + PAIR_NSTAGE * PAIR_STAGE);
    uint64_t* empty = full + PAIR_NSTAGE;
    uint64_t* tmem_full = empty + PAIR_NSTAGE;     // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2] (only the even CTA's are waited on)
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&P.tmA);
        prefetch_tmap(&P.tmB);
        for (int s = 0; s < PAIR_NSTAGE; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 2 * (GEMM_THREADS / 32 - 2));   // every epilogue warp of BOTH CTAs
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc2_rt(tmem_slot, 2 * TCOLS);
    tc_fence_before();
    cluster_sync_all();                            // barriers of both CTAs initialised before any remote arrive / TMA
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        if (lane == 0) {                           // TMA producer (both CTAs)
            int s = 0, first = 1;
            uint32_t ph = 0;
            for (int tile = cluster; tile < ntiles; tile += nclusters) {
                const int txy = tile % nxy, sp = tile / nxy;
                const int m0 = (txy % args.gx) * 256 + rank * 128, nb = (txy / args.gx) * 256 + rank * 128;
                const int kb0 = sp * kper, KB = KBT - kb0 < kper ? KBT - kb0 : kper;
                for (int kb = 0; kb < KB; ++kb) {
                    if (!first) { if (++s == PAIR_NSTAGE) { s = 0; ph ^= 1; } }
                    first = 0;
                    mbar_wait(&empty[s], ph ^ 1);
                    if (rank == 0) mbar_expect_tx(&full[s], 2 * PAIR_STAGE);      // this CTA's 32 KB + the peer's 32 KB
                    uint8_t* dst = smem + s * PAIR_STAGE;
                    if constexpr (MNMAJ) {                                        // boxes {64 MN elements, 64 K rows}
                        tma_load_2d_pair(dst, &P.tmA, &full[s], m0, (kb0 + kb) * 64);
                        tma_load_2d_pair(dst + 8192, &P.tmA, &full[s], m0 + 64, (kb0 + kb) * 64);
                        tma_load_2d_pair(dst + PAIR_HALF, &P.tmB, &full[s], nb, (kb0 + kb) * 64);
                        tma_load_2d_pair(dst + PAIR_HALF + 8192, &P.tmB, &full[s], nb + 64, (kb0 + kb) * 64);
                    } else {
                        tma_load_2d_pair(dst, &P.tmA, &full[s], (kb0 + kb) * 64, m0);
                        tma_load_2d_pair(dst + PAIR_HALF, &P.tmB, &full[s], (kb0 + kb) * 64, nb);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {              // MMA issuer: one thread of the even CTA drives both tensor cores
            constexpr uint32_t idesc = MNMAJ ? umma_idesc_bf16_mn(256, 256) : umma_idesc_bf16(256, 256);
            const uint64_t d0 = MNMAJ ? umma_desc_sw128_mn(smem_u32(smem), 8192) : umma_desc_sw128(smem_u32(smem));
            constexpr int KSTEP = MNMAJ ? 128 : 2;
            int s = 0, first = 1, tl = 0;
            uint32_t ph = 0;
            for (int tile = cluster; tile < ntiles; tile += nclusters) {
                const int sp = tile / nxy, kb0 = sp * kper, KB = KBT - kb0 < kper ? KBT - kb0 : kper;
                const int buf = tl & 1, use = tl >> 1;
                mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)buf * TCOLS;
                for (int kb = 0; kb < KB; ++kb) {
                    if (!first) { if (++s == PAIR_NSTAGE) { s = 0; ph ^= 1; } }
                    first = 0;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint64_t da = d0 + (uint64_t)(s * (PAIR_STAGE >> 4));
                    const uint64_t db = da + (uint64_t)(PAIR_HALF >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma2_bf16(tmem_d, da + KSTEP * k, db + KSTEP * k, idesc, (kb | k) != 0);
                    umma2_commit_both(&empty[s]);
                }
                umma2_commit_both(&tmem_full[buf]);
                ++tl;
            }
        }
    } else {                                       // epilogue warps (both CTAs): this CTA's 128 rows x 256 features
        const int q = warp & 3, half = (warp - 2) >> 2;
        int tl = 0;
        for (int tile = cluster; tile < ntiles; tile += nclusters) {
            const int txy = tile % nxy, sp = tile / nxy;
            const int m0 = (txy % args.gx) * 256 + rank * 128, n0 = (txy / args.gx) * 256;
            const int buf = tl & 1, use = tl >> 1;
            mbar_wait(&tmem_full[buf], use & 1);
            tc_fence_after();
            float* of = (P.out_f32 && args.gz > 1) ? P.out_f32 + (size_t)sp * args.split_stride : P.out_f32;
            if constexpr (SWAP) swap_epilogue<128, 256>(P, of, tmem_base + (uint32_t)buf * TCOLS, m0, n0, n_rows, q, half, lane);
            else rows_epilogue<256>(P, of, tmem_base + (uint32_t)buf * TCOLS, m0, n0, n_rows, q, half, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_even_cta(&tmem_empty[buf]);
            ++tl;
        }
    }
    tc_fence_before();
    cluster_sync_all();                            // the peer may still read this CTA's shared memory / signal its barriers
    if (warp == 1) tmem_dealloc2_rt(tmem_base, 2 * TCOLS);
}

// CTA-pair launch.  mn = false: out (n_rows, n_feat) = act (n_rows, K) . W (n_feat, K)^T [+bias][relu][mask];
// mn = true: o.act = X (K, n_rows), o.W = Y (K, n_feat) row-major, out = X^T . Y (fp32, optional split-K into split_scratch).
static int launch_gemm_pair(const GemmOperands& o, int n_rows, int K, bool mn, int ksplit, float* split_scratch, cudaStream_t st,
                            bool swap = false) {
    GemmArgs args;
    memset(&args, 0, sizeof(args));
    const int Kp = (K + 63) / 64 * 64;
    args.n_rows = n_rows;
    args.K = Kp;
    args.ksplit = 1;
    args.gz = 1;
    if (ksplit > 1) {
        const int KB = Kp / 64, per = (KB + ksplit - 1) / ksplit;
        ksplit = (KB + per - 1) / per;                 // no empty split
    }
    if (ksplit > 1) {
        if (!split_scratch || !o.out_f32) return fail(SRNN_ERR_ARG, "gemm pair: split-K needs scratch and an fp32 output");
        args.gz = ksplit;
        args.split_stride = (long long)n_rows * o.ld_out;
    }
    if (mn) {
        SRNN_TRY(make_tmap_bf16(&args.p[0].tmA, o.act, K, n_rows, o.ld_act, 64));
        SRNN_TRY(make_tmap_bf16(&args.p[0].tmB, o.W, K, o.n_feat, o.ld_w, 64));
    } else if (swap) {                                 // lanes = features: A = weights, B = activations
        SRNN_TRY(make_tmap_bf16(&args.p[0].tmA, o.W, o.n_feat, K, o.ld_w, 128));
        SRNN_TRY(make_tmap_bf16(&args.p[0].tmB, o.act, n_rows, K, o.ld_act, 128));
    } else {
        SRNN_TRY(make_tmap_bf16(&args.p[0].tmA, o.act, n_rows, K, o.ld_act, 128));
        SRNN_TRY(make_tmap_bf16(&args.p[0].tmB, o.W, o.n_feat, K, o.ld_w, 128));
    }
    args.p[0].bias = o.bias;
    args.p[0].addend = o.addend;
    args.p[0].out_f32 = args.gz > 1 ? split_scratch : o.out_f32;
    args.p[0].out_bf16 = o.out_bf16;
    args.p[0].mask = o.mask;
    args.p[0].n_feat = o.n_feat;
    args.p[0].ld_add = o.ld_add;
    args.p[0].ld_out = o.ld_out;
    args.p[0].relu = o.relu;
    args.gx = swap ? cdiv(o.n_feat, 256) : cdiv(n_rows, 256);      // pair tiles along the A operand
    args.gy = swap ? cdiv(n_rows, 256) : cdiv(o.n_feat, 256);
    static bool attr_set = false;
    if (!attr_set) {
        SRNN_CUDA(cudaFuncSetAttribute(k_gemm_umma_pair<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
        SRNN_CUDA(cudaFuncSetAttribute(k_gemm_umma_pair<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
        SRNN_CUDA(cudaFuncSetAttribute((k_gemm_umma_pair<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
        attr_set = true;
    }
    const int ntiles = args.gx * args.gy * args.gz;
    int nclusters = g_gemm_sms > 1 ? g_gemm_sms / 2 : 1;
    if (nclusters > ntiles) nclusters = ntiles;
    if (mn) SRNN_LAUNCH(k_gemm_umma_pair<true>, dim3(2 * nclusters), GEMM_THREADS, PAIR_SMEM, st, args);
    else if (swap) SRNN_LAUNCH((k_gemm_umma_pair<false, true>), dim3(2 * nclusters), GEMM_THREADS, PAIR_SMEM, st, args);
    else SRNN_LAUNCH(k_gemm_umma_pair<false>, dim3(2 * nclusters), GEMM_THREADS, PAIR_SMEM, st, args);
    if (args.gz > 1) return sum_splits(split_scratch, args.gz, (size_t)n_rows * o.ld_out, (size_t)args.split_stride, o.out_f32, st);
    return SRNN_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// ONE-WAVE CTA-pair GEMM for the generation-time upsampling (<= 256 batch rows x fs.H features, C2: 256 x 20480 x 1024).
// With 256-feature pair tiles that problem has 80 tiles for 74 CTA pairs, i.e. a second wave for 6 of them; the per-SM TMA
// ingest (every CTA streams all K of its 128 activation rows + its half of the weight tile) sets the duration, so the second
// wave nearly doubles it.  Here a pair tile is 256 rows x (256 + WIDE_NX) features -- two UMMAs per K step, N = 256 and
// N = WIDE_NX, into TMEM columns [0, 256) and [256, 256 + WIDE_NX) -- which makes 72 tiles at C2: one wave, 34 KB per CTA
// and k-block.  One tile per cluster, single accumulator buffer.
// With K = 1024 only, the accumulator drain is as long as the K loop if it goes through per-lane global stores (each lane owns
// a row, 80 KB apart: 32 partial sectors per instruction; measured 2/3 of the kernel), so plain fp32 outputs take the other
// road: TMEM -> registers (+bias) -> 128B-swizzled staging tiles in the idle TMA ring -> cp.async.bulk.tensor stores of
// {32 features x 128 rows} boxes, one per 32-feature group as soon as the eight epilogue warps have filled it.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int WIDE_NX = 32;
constexpr int WIDE_NW = 256 + WIDE_NX;
constexpr int WIDE_BX = WIDE_NX / 2 * 128;                 // this CTA's half of the extra weight rows
constexpr int WIDE_STAGE = 2 * PAIR_HALF + WIDE_BX;       // 34 KB, a multiple of the 1024-byte swizzle atom
constexpr int WIDE_NSTAGE = 6;
constexpr int WIDE_SMEM = WIDE_NSTAGE * WIDE_STAGE + 1024 + 256 + WIDE_NW * 4 /*bias slice*/;
static_assert(WIDE_NW / 32 * 16384 <= WIDE_NSTAGE * WIDE_STAGE, "the output staging tiles reuse the TMA ring");
static_assert(WIDE_STAGE % 1024 == 0, "stage must keep the 1024-byte alignment of the swizzled tiles");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
k_gemm_umma_pair_wide(const __grid_constant__ GemmArgs args) {
    const GemmProb& P = args.p[0];
    const CUtensorMap* tmBX = &args.p[1].tmB;      // same weight matrix, {64, WIDE_NX / 2}-row boxes
    const int n_rows = args.n_rows, KB = args.K / 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + WIDE_NSTAGE * WIDE_STAGE);
    uint64_t* empty = full + WIDE_NSTAGE;
    uint64_t* tmem_full = empty + WIDE_NSTAGE;
    uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);
    float* sbias = (float*)(smem + WIDE_NSTAGE * WIDE_STAGE + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int tile = blockIdx.x >> 1;
    const int m0 = (tile % args.gx) * 256 + rank * 128, n0 = (tile / args.gx) * WIDE_NW;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&P.tmA);
        prefetch_tmap(&P.tmB);
        prefetch_tmap(tmBX);
        prefetch_tmap(&args.p[1].tmA);
        for (int s = 0; s < WIDE_NSTAGE; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    pdl_trigger();
    if (warp == 1) tmem_alloc2_rt(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {                           // TMA producer (both CTAs): own activation rows + own halves of the weights
            int s = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(&empty[s], ph ^ 1);
                if (rank == 0) mbar_expect_tx(&full[s], 2 * WIDE_STAGE);
                uint8_t* dst = smem + s * WIDE_STAGE;
                tma_load_2d_pair(dst, &P.tmA, &full[s], kb * 64, m0);
                tma_load_2d_pair(dst + PAIR_HALF, &P.tmB, &full[s], kb * 64, n0 + rank * 128);
                tma_load_2d_pair(dst + 2 * PAIR_HALF, tmBX, &full[s], kb * 64, n0 + 256 + rank * (WIDE_NX / 2));
                if (++s == WIDE_NSTAGE) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(256, 256), idescx = umma_idesc_bf16(256, WIDE_NX);
            const uint64_t d0 = umma_desc_sw128(smem_u32(smem));
            int s = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint64_t da = d0 + (uint64_t)(s * (WIDE_STAGE >> 4));
                const uint64_t db = da + (uint64_t)(PAIR_HALF >> 4), dbx = da + (uint64_t)((2 * PAIR_HALF) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma2_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    umma2_bf16(tmem_base + 256, da + 2 * k, dbx + 2 * k, idescx, (kb | k) != 0);
                }
                umma2_commit_both(&empty[s]);
                if (++s == WIDE_NSTAGE) { s = 0; ph ^= 1; }
            }
            umma2_commit_both(tmem_full);
        }
    } else {                                       // epilogue warps (both CTAs): this CTA's 128 rows x WIDE_NW features
        const int q = warp & 3, half = (warp - 2) >> 2, et = (int)threadIdx.x - 64;
        const bool tma_epi = P.out_f32 && !P.out_bf16 && !P.addend && !P.mask;      // uniform
        if (tma_epi) {                             // bias slice -> shared memory while the K loop runs
            for (int i = et; i < WIDE_NW; i += GEMM_THREADS - 64) sbias[i] = (P.bias && n0 + i < P.n_feat) ? P.bias[n0 + i] : 0.f;
            epi_bar_sync(1, GEMM_THREADS - 64);
        }
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (tma_epi) {
            const CUtensorMap* tmOut = &args.p[1].tmA;
            const int r = 32 * q + lane, relu = P.relu;
            const int left = P.n_feat - n0, G = left >= WIDE_NW ? WIDE_NW / 32 : (left + 31) / 32;
            for (int g = 0; g < G; ++g) {
                const int c = 32 * g + 16 * half;
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + c, v);
                uint8_t* row = smem + g * 16384 + r * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 o;
                    o.x = v[4 * j] + sbias[c + 4 * j];
                    o.y = v[4 * j + 1] + sbias[c + 4 * j + 1];
                    o.z = v[4 * j + 2] + sbias[c + 4 * j + 2];
                    o.w = v[4 * j + 3] + sbias[c + 4 * j + 3];
                    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    *reinterpret_cast<float4*>(row + (((4 * half + j) ^ (r & 7)) << 4)) = o;      // 128B swizzle: chunk ^= row % 8
                }
                fence_proxy_async_smem();
                epi_bar_sync(1, GEMM_THREADS - 64);
                if (et == 0) {
                    tma_store_2d(tmOut, smem + g * 16384, n0 + 32 * g, m0);
                    bulk_commit();
                }
            }
            if (et == 0) bulk_wait_read0();        // the staging tiles must outlive the bulk reads
        } else {
            rows_epilogue<WIDE_NW>(P, P.out_f32, tmem_base, m0, n0, n_rows, q, half, lane);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc2_rt(tmem_base, 512);
}

static void ensure_gemm_sms() {
    if (g_gemm_sms < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&g_gemm_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_gemm_sms = 0;
    }
}
// worthwhile exactly when 256-feature pair tiles need a second wave and (256 + WIDE_NX)-feature tiles do not
bool gemm_umma_pair_wide_ok(int n_feat, int n_rows, bool force) {
    static const int pair_mode = getenv("SRNN_GEMM_PAIR") ? atoi(getenv("SRNN_GEMM_PAIR")) : 1;
    static const bool off = getenv("SRNN_UP_WIDE") && atoi(getenv("SRNN_UP_WIDE")) == 0;
    ensure_gemm_sms();
    const int nclusters = g_gemm_sms / 2;
    if (!pair_mode || n_rows < 1 || nclusters < 1) return false;
    const long long tiles = (long long)cdiv(n_rows, 256) * cdiv(n_feat, WIDE_NW);
    if (force) return tiles <= nclusters;
    return !off && n_rows > 128 && n_rows <= 256 && cdiv(n_feat, 256) > nclusters && tiles <= nclusters;
}
int gemm_umma_pair_wide(const GemmOperands& o, int n_rows, int K, cudaStream_t st) {
    ensure_gemm_sms();
    if (K % 64 || o.ld_out % 8 || (o.addend && o.ld_add % 4)) return fail(SRNN_ERR_ARG, "gemm pair (wide): K %% 64, ld_out %% 8, ld_add %% 4");
    GemmArgs args;
    memset(&args, 0, sizeof(args));
    args.n_rows = n_rows;
    args.K = K;
    args.ksplit = 1;
    args.gx = cdiv(n_rows, 256);
    args.gy = cdiv(o.n_feat, WIDE_NW);
    args.gz = 1;
    if ((long long)args.gx * args.gy > g_gemm_sms / 2) return fail(SRNN_ERR_UNSUPPORTED, "gemm pair (wide): more tiles than CTA pairs");
    SRNN_TRY(make_tmap_bf16(&args.p[0].tmA, o.act, n_rows, K, o.ld_act, 128));
    SRNN_TRY(make_tmap_bf16(&args.p[0].tmB, o.W, o.n_feat, K, o.ld_w, 128));
    SRNN_TRY(make_tmap_bf16(&args.p[1].tmB, o.W, o.n_feat, K, o.ld_w, WIDE_NX / 2));
    if (o.out_f32 && !o.out_bf16 && !o.addend && !o.mask)
        SRNN_TRY(make_tmap_f32_out(&args.p[1].tmA, o.out_f32, n_rows, o.n_feat, o.ld_out, 128));
    args.p[0].bias = o.bias;
    args.p[0].addend = o.addend;
    args.p[0].out_f32 = o.out_f32;
    args.p[0].out_bf16 = o.out_bf16;
    args.p[0].mask = o.mask;
    args.p[0].n_feat = o.n_feat;
    args.p[0].ld_add = o.ld_add;
    args.p[0].ld_out = o.ld_out;
    args.p[0].relu = o.relu;
    static bool attr_set = false;
    if (!attr_set) {
        SRNN_CUDA(cudaFuncSetAttribute(k_gemm_umma_pair_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, WIDE_SMEM));
        attr_set = true;
    }
    SRNN_LAUNCH_PDL(k_gemm_umma_pair_wide, dim3(2 * args.gx * args.gy), dim3(GEMM_THREADS), (size_t)WIDE_SMEM, st, args);
    return SRNN_OK;
}

// rows = false: swap-AB orientation, tile bm (128|64) features x bn (32..256) rows.
// rows = true : activation rows on the lanes, tile 128 rows x bn (128|256) features (vector epilogue); needs ld_out % 8 == 0.
// ksplit > 1 (single problem, fp32 output only): the K loop is cut into ksplit ranges, each CTA writes its partial tile
// into split_scratch (ksplit x n_rows x ld_out floats) and a second kernel sums them in fixed order into out_f32.
int gemm_umma_ex(const GemmOperands* ops, int nprob, int n_rows, int K, int bm, int bn, bool rows, int ksplit,
                 float* split_scratch, cudaStream_t st) {
    if (K % 64 || K <= 0) return fail(SRNN_ERR_ARG, "gemm_umma: K=%d must be a positive multiple of 64", K);
    if (nprob < 1 || nprob > 2) return fail(SRNN_ERR_ARG, "gemm_umma: 1 or 2 problems per launch");
    if (g_gemm_sms < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&g_gemm_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_gemm_sms = 0;
    }
    GemmArgs args;
    memset(&args, 0, sizeof(args));
    args.n_rows = n_rows;
    args.K = K;
    args.ksplit = 1;
    if (ksplit > 1) {
        const int KB = K / 64, per = (KB + ksplit - 1) / ksplit;
        ksplit = (KB + per - 1) / per;                 // no empty split
    }
    if (ksplit > 1) {
        if (nprob != 1 || !ops[0].out_f32 || ops[0].out_bf16 || ops[0].bias || ops[0].addend || ops[0].relu || ops[0].mask ||
            !split_scratch)
            return fail(SRNN_ERR_ARG, "gemm_umma: split-K needs a single plain fp32-output problem and a scratch buffer");
        args.ksplit = ksplit;
        args.split_stride = (long long)n_rows * ops[0].ld_out;
    }
    int max_feat = 0;
    for (int i = 0; i < nprob; ++i) {
        const GemmOperands& o = ops[i];
        if (rows) {
            if (o.ld_out % 8 || (o.addend && o.ld_add % 4)) return fail(SRNN_ERR_ARG, "gemm_umma rows: ld_out %% 8, ld_add %% 4");
            SRNN_TRY(make_tmap_bf16(&args.p[i].tmA, o.act, n_rows, K, o.ld_act, 128));
            SRNN_TRY(make_tmap_bf16(&args.p[i].tmB, o.W, o.n_feat, K, o.ld_w, bn));
        } else {
            SRNN_TRY(make_tmap_bf16(&args.p[i].tmA, o.W, o.n_feat, K, o.ld_w, bm));
            SRNN_TRY(make_tmap_bf16(&args.p[i].tmB, o.act, n_rows, K, o.ld_act, bn));
        }
        args.p[i].bias = o.bias;
        args.p[i].addend = o.addend;
        args.p[i].out_f32 = args.ksplit > 1 ? split_scratch : o.out_f32;
        args.p[i].out_bf16 = o.out_bf16;
        args.p[i].mask = o.mask;
        args.p[i].n_feat = o.n_feat;
        args.p[i].ld_add = o.ld_add;
        args.p[i].ld_out = o.ld_out;
        args.p[i].relu = o.relu;
        if (o.n_feat > max_feat) max_feat = o.n_feat;
    }
    static long long* g_trace = nullptr;
    if (getenv("SRNN_TRACE_GEMM")) {
        if (!g_trace) cudaMallocManaged((void**)&g_trace, sizeof(long long) * 64);
        if (g_trace) memset(g_trace, 0, sizeof(long long) * 64);
        args.trace = g_trace;
    }
    int rc = SRNN_ERR_UNSUPPORTED;
#define SRNN_GEMM_CASE(BM_, BN_, ROWS_) \
    if (bm == BM_ && bn == BN_ && rows == ROWS_) rc = launch_gemm_umma<BM_, BN_, ROWS_>(args, nprob, max_feat, st);
    SRNN_GEMM_CASE(128, 256, false)
    SRNN_GEMM_CASE(128, 128, false)
    SRNN_GEMM_CASE(128, 64, false)
    SRNN_GEMM_CASE(128, 32, false)
    SRNN_GEMM_CASE(64, 32, false)
    SRNN_GEMM_CASE(64, 64, false)
    SRNN_GEMM_CASE(128, 256, true)
    SRNN_GEMM_CASE(128, 128, true)
#undef SRNN_GEMM_CASE
    if (rc == SRNN_ERR_UNSUPPORTED) return fail(SRNN_ERR_UNSUPPORTED, "gemm_umma: tile %dx%d rows=%d not instantiated", bm, bn, (int)rows);
    if (rc == SRNN_OK && args.trace) {
        cudaStreamSynchronize(st);
        const long long t0 = g_trace[0];
        fprintf(stderr, "[gemm trace %dx%d K=%d] mma done=%lld epilogue done=%lld | producer issue:", bm, bn, K,
                g_trace[1] - t0, g_trace[2] - t0);
        for (int kb = 0; kb < K / 64 && kb < 24; ++kb) fprintf(stderr, " %lld", g_trace[8 + kb] - t0);
        fprintf(stderr, " | consumer full:");
        for (int kb = 0; kb < K / 64 && kb < 24; ++kb) fprintf(stderr, " %lld", g_trace[32 + kb] - t0);
        fprintf(stderr, "\n");
    }
    if (rc == SRNN_OK && args.ksplit > 1)
        rc = sum_splits(split_scratch, args.ksplit, (size_t)n_rows * ops[0].ld_out, (size_t)args.split_stride, ops[0].out_f32, st);
    return rc;
}

int gemm_umma_multi(const GemmOperands* ops, int nprob, int n_rows, int K, int bm, int bn, cudaStream_t st) {
    return gemm_umma_ex(ops, nprob, n_rows, K, bm, bn, false, 1, nullptr, st);
}

// out (M, N) fp32 = X^T . Y with X (Ktot, M; ld_x) and Y (Ktot, N; ld_y) bf16 row-major: both operands MN-major, read in place.
// Ktot needs no padding (TMA zero-fills the tail k-block).  Optional split-K as in gemm_umma_ex.
int gemm_umma_tn(const __nv_bfloat16* X, int ld_x, const __nv_bfloat16* Y, int ld_y, int M, int N, int Ktot, float* out,
                 int ld_out, int ksplit, float* split_scratch, cudaStream_t st) {
    if (ld_out % 8 || M < 1 || N < 1 || Ktot < 1) return fail(SRNN_ERR_ARG, "gemm_umma_tn: bad shape");
    if (g_gemm_sms < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&g_gemm_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_gemm_sms = 0;
    }
    {
        static const int pair_mode = getenv("SRNN_GEMM_PAIR") ? atoi(getenv("SRNN_GEMM_PAIR")) : 1;
        // CTA-pair kernel when the 256 x 256 pair tiles (times the K splits) can fill the GPU
        if (pair_mode && ((M >= 256 && N >= 256) || pair_mode == 2)) {
            GemmOperands o{Y, X, nullptr, nullptr, out, nullptr, N, ld_y, ld_x, 0, ld_out, 0, nullptr};
            int ks = ksplit;
            if (pair_mode != 2) {      // re-derive the split for 256-wide tiles: ~2 work items per cluster
                const int tiles = cdiv(M, 256) * cdiv(N, 256);
                ks = cdiv(g_gemm_sms, tiles);
                if (ks > Ktot / 512) ks = Ktot / 512;
                if (ks > ksplit) ks = ksplit;          // the caller's split respects the capacity of split_scratch
                if (!split_scratch) ks = 1;
            }
            return launch_gemm_pair(o, M, Ktot, true, ks < 2 ? 1 : ks, split_scratch, st);
        }
    }
    const int bn = N <= 128 ? 128 : 256;
    const int Kp = (Ktot + 63) / 64 * 64;
    GemmArgs args;
    memset(&args, 0, sizeof(args));
    args.n_rows = M;
    args.K = Kp;
    args.ksplit = 1;
    if (ksplit > 1) {
        const int KB = Kp / 64, per = (KB + ksplit - 1) / ksplit;
        ksplit = (KB + per - 1) / per;
    }
    if (ksplit > 1) {
        if (!split_scratch) return fail(SRNN_ERR_ARG, "gemm_umma_tn: split-K needs scratch");
        args.ksplit = ksplit;
        args.split_stride = (long long)M * ld_out;
    }
    SRNN_TRY(make_tmap_bf16(&args.p[0].tmA, X, Ktot, M, ld_x, 64));
    SRNN_TRY(make_tmap_bf16(&args.p[0].tmB, Y, Ktot, N, ld_y, 64));
    args.p[0].out_f32 = args.ksplit > 1 ? split_scratch : out;
    args.p[0].n_feat = N;
    args.p[0].ld_out = ld_out;
    int rc = bn == 128 ? launch_gemm_umma<128, 128, true, true>(args, 1, N, st) : launch_gemm_umma<128, 256, true, true>(args, 1, N, st);
    if (rc == SRNN_OK && args.ksplit > 1) rc = sum_splits(split_scratch, args.ksplit, (size_t)M * ld_out, (size_t)args.split_stride, out, st);
    return rc;
}

// The big teacher-forced contractions (hundreds of rows or more): ROWS orientation, 128 x 256 (or 128 x 128) tiles.
int gemm_umma_rows(const GemmOperands& o, int n_rows, int K, int ksplit, float* split_scratch, cudaStream_t st) {
    if (g_gemm_sms < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&g_gemm_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_gemm_sms = 0;
    }
    // SRNN_GEMM_PAIR: 0 = never, 1 = when there is at least a wave of pair tiles (default), 2 = always (tests)
    static const int pair_mode = getenv("SRNN_GEMM_PAIR") ? atoi(getenv("SRNN_GEMM_PAIR")) : 1;
    // CTA-pair kernel: long-M problems with at least a wave of 256 x 256 pair tiles
    if (pair_mode && ksplit <= 1 && K % 64 == 0 && (o.n_feat >= 256 || pair_mode == 2) && o.ld_out % 8 == 0 && (!o.addend || o.ld_add % 4 == 0) &&
        (pair_mode == 2 || (long long)cdiv(n_rows, 256) * cdiv(o.n_feat, 256) >= g_gemm_sms / 2))
        return launch_gemm_pair(o, n_rows, K, false, 1, nullptr, st);
    const int bn = o.n_feat <= 128 ? 128 : 256;
    return gemm_umma_ex(&o, 1, n_rows, K, 128, bn, true, ksplit, split_scratch, st);
}

// swap-AB through the CTA-pair kernel: worthwhile when the 256-feature pair tiles fill the clusters (>= SMs / 2 tiles) and
// there are more than 128 batch rows (each CTA of the pair takes 128 of them)
bool gemm_umma_swap_pair_ok(int n_feat, int n_rows) {
    static const int pair_mode = getenv("SRNN_GEMM_PAIR") ? atoi(getenv("SRNN_GEMM_PAIR")) : 1;
    if (g_gemm_sms < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&g_gemm_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_gemm_sms = 0;
    }
    // measured at the C2 upsampling (256 x 20480 x 1024): 31 us against 30 us for the single-CTA kernel with two CTAs per SM
    // -- the launch is not ingest bound -- so this form is opt-in (SRNN_SWAP_PAIR=1) and kept for larger batches / tests
    return pair_mode && getenv("SRNN_SWAP_PAIR") && n_rows > 128 && n_rows <= 256 && cdiv(n_feat, 256) >= g_gemm_sms / 2;
}
int gemm_umma_swap_pair(const GemmOperands& o, int n_rows, int K, cudaStream_t st) {
    return launch_gemm_pair(o, n_rows, K, false, 1, nullptr, st, true);
}

int gemm_umma(const __nv_bfloat16* W, int n_feat, const __nv_bfloat16* act, int n_rows, int K, int ld_w, int ld_act,
              const float* bias, const float* addend, int ld_add, float* out_f32, __nv_bfloat16* out_bf16,
              int ld_out, int relu, int bm, int bn, cudaStream_t st) {
    GemmOperands o{W, act, bias, addend, out_f32, out_bf16, n_feat, ld_w, ld_act, ld_add, ld_out, relu, nullptr};
    return gemm_umma_multi(&o, 1, n_rows, K, bm, bn, st);
}

// ---- fp32 -> bf16 with zero padding (test hook + weight packing) ------------------------------------------------
__global__ void k_f32_to_bf16_pad(const float* __restrict__ src, int rows, int cols, int ld_src,
                                  __nv_bfloat16* __restrict__ dst, int rows_p, int cols_p) {
    const size_t total = (size_t)rows_p * cols_p;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols_p), c = (int)(i % cols_p);
        const float v = (r < rows && c < cols) ? src[(size_t)r * ld_src + c] : 0.f;
        dst[i] = __float2bfloat16(v);
    }
}
int f32_to_bf16_pad(const float* src, int rows, int cols, int ld_src, __nv_bfloat16* dst, int rows_p, int cols_p,
                    cudaStream_t st) {
    const size_t total = (size_t)rows_p * cols_p;
    int grid = (int)((total + 255) / 256 > 8192 ? 8192 : (total + 255) / 256);
    if (grid < 1) grid = 1;
    SRNN_LAUNCH(k_f32_to_bf16_pad, grid, 256, 0, st, src, rows, cols, ld_src, dst, rows_p, cols_p);
    return SRNN_OK;
}

// ---- multi-matrix fp32 -> bf16 (+ transposed bf16) pack: one launch for every tcgen05 operand copy of the weights ---------
struct PackBf16Args {
    PackBf16Item it[PACK_BF16_MAX];
    int n, total_tiles;
};
// tile = 32 rows x 64 columns; block (32, 8): thread (tx, ty) owns columns 2tx, 2tx+1 of rows ty, ty+8, ...
__global__ void __launch_bounds__(256) k_pack_bf16_multi(const __grid_constant__ PackBf16Args a) {
    __shared__ float tile[32][65];
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
        int m = 0;
        while (m + 1 < a.n && a.it[m + 1].tile0 <= t) ++m;
        const PackBf16Item& I = a.it[m];
        const int tcols = (I.cols + 63) >> 6;
        const int lt = t - I.tile0, r0 = (lt / tcols) * 32, c0 = (lt % tcols) * 64;
        const bool full = r0 + 32 <= I.rows && c0 + 64 <= I.cols && (I.cols & 1) == 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + ty + 8 * i, c = c0 + 2 * tx;
            float2 v = make_float2(0.f, 0.f);
            if (full) {
                v = *reinterpret_cast<const float2*>(I.src + (size_t)r * I.cols + c);
            } else if (r < I.rows) {
                if (c < I.cols) v.x = I.src[(size_t)r * I.cols + c];
                if (c + 1 < I.cols) v.y = I.src[(size_t)r * I.cols + c + 1];
            }
            if (I.dst32) {
                if (full) {
                    *reinterpret_cast<float2*>(I.dst32 + (size_t)r * I.cols + c) = v;
                } else if (r < I.rows) {
                    if (c < I.cols) I.dst32[(size_t)r * I.cols + c] = v.x;
                    if (c + 1 < I.cols) I.dst32[(size_t)r * I.cols + c + 1] = v.y;
                }
            }
            if (I.dst) {
                if (full) {
                    *reinterpret_cast<__nv_bfloat162*>(I.dst + (size_t)r * I.cols + c) = __floats2bfloat162_rn(v.x, v.y);
                } else if (r < I.rows) {
                    if (c < I.cols) I.dst[(size_t)r * I.cols + c] = __float2bfloat16(v.x);
                    if (c + 1 < I.cols) I.dst[(size_t)r * I.cols + c + 1] = __float2bfloat16(v.y);
                }
            }
            tile[ty + 8 * i][2 * tx] = v.x;
            tile[ty + 8 * i][2 * tx + 1] = v.y;
        }
        if (I.dst_t) {
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 8; ++i) {                  // dst_t[c][r]: a warp writes 32 consecutive r of one column c
                const int c = c0 + ty + 8 * i, r = r0 + tx;
                if (c < I.cols && r < I.rows) I.dst_t[(size_t)c * I.rows + r] = __float2bfloat16(tile[tx][ty + 8 * i]);
            }
        }
        __syncthreads();
    }
}
int pack_bf16_multi(PackBf16Item* items, int n, cudaStream_t st) {
    if (n < 1 || n > PACK_BF16_MAX) return fail(SRNN_ERR_ARG, "pack_bf16_multi: 1..%d matrices", PACK_BF16_MAX);
    PackBf16Args a;
    memset(&a, 0, sizeof(a));
    int tiles = 0;
    for (int i = 0; i < n; ++i) {
        items[i].tile0 = tiles;
        tiles += cdiv(items[i].rows, 32) * cdiv(items[i].cols, 64);
        a.it[i] = items[i];
    }
    a.n = n;
    a.total_tiles = tiles;
    const int grid = tiles < 148 * 16 ? tiles : 148 * 16;
    SRNN_LAUNCH(k_pack_bf16_multi, grid, dim3(32, 8), 0, st, a);
    return SRNN_OK;
}

// ---- fp32 -> split bf16 [hi | hi | lo] / [hi | lo | hi] along K (SRNN_MODE_BF16X3 operands; common.cuh) ------------------
__global__ void k_split3_bf16(const float* __restrict__ src, long long rows, int K, long long ld_src,
                              __nv_bfloat16* __restrict__ dst, int weight_order) {
    const int K4 = K >> 2;
    const long long total = rows * K4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / K4;
        const int c = (int)(i - r * K4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(src + r * ld_src + c);
        const float x[4] = {v.x, v.y, v.z, v.w};
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hi[j] = __float2bfloat16(x[j]);
            lo[j] = __float2bfloat16(x[j] - __bfloat162float(hi[j]));
        }
        const uint2 h2 = *reinterpret_cast<const uint2*>(hi), l2 = *reinterpret_cast<const uint2*>(lo);
        __nv_bfloat16* d = dst + r * 3 * K + c;
        *reinterpret_cast<uint2*>(d) = h2;
        *reinterpret_cast<uint2*>(d + K) = weight_order ? l2 : h2;
        *reinterpret_cast<uint2*>(d + 2 * K) = weight_order ? h2 : l2;
    }
}
int split3_bf16(const float* src, long long rows, int K, long long ld_src, __nv_bfloat16* dst, int weight_order, cudaStream_t st) {
    if (K % 4 || ld_src % 4) return fail(SRNN_ERR_ARG, "split3_bf16: K and the leading dimension must be multiples of 4");
    const long long total = rows * (K >> 2);
    int grid = (int)((total + 255) / 256 > 16384 ? 16384 : (total + 255) / 256);
    if (grid < 1) grid = 1;
    SRNN_LAUNCH(k_split3_bf16, grid, 256, 0, st, src, rows, K, ld_src, dst, weight_order);
    return SRNN_OK;
}

__global__ void k_split3_planes_bf16(const float* __restrict__ src, size_t n4, size_t n, __nv_bfloat16* __restrict__ dst, int weight_order) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        const float x[4] = {v.x, v.y, v.z, v.w};
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hi[j] = __float2bfloat16(x[j]);
            lo[j] = __float2bfloat16(x[j] - __bfloat162float(hi[j]));
        }
        const uint2 h2 = *reinterpret_cast<const uint2*>(hi), l2 = *reinterpret_cast<const uint2*>(lo);
        reinterpret_cast<uint2*>(dst)[i] = h2;
        reinterpret_cast<uint2*>(dst + n)[i] = weight_order ? l2 : h2;
        reinterpret_cast<uint2*>(dst + 2 * n)[i] = weight_order ? h2 : l2;
    }
}
int split3_planes_bf16(const float* src, size_t n, __nv_bfloat16* dst, int weight_order, cudaStream_t st) {
    if (n % 4) return fail(SRNN_ERR_ARG, "split3_planes_bf16: element count must be a multiple of 4");
    const size_t n4 = n / 4;
    int grid = (int)((n4 + 255) / 256 > 16384 ? 16384 : (n4 + 255) / 256);
    if (grid < 1) grid = 1;
    SRNN_LAUNCH(k_split3_planes_bf16, grid, 256, 0, st, src, n4, n, dst, weight_order);
    return SRNN_OK;
}

int gemm_x3(int rows, int n_feat, int K, const float* A, long long lda, const __nv_bfloat16* W3, const float* bias, int relu,
            float* C, int ldc, __nv_bfloat16* s3, int bm, int bn, cudaStream_t st) {
    SRNN_TRY(split3_bf16(A, rows, K, lda, s3, 0, st));
    GemmOperands o{W3, s3, bias, nullptr, C, nullptr, n_feat, 3 * K, 3 * K, 0, ldc, relu, nullptr};
    if (!bm && rows >= 256 && ldc % 8 == 0 && n_feat % 16 == 0) return gemm_umma_rows(o, rows, 3 * K, 1, nullptr, st);
    const int pick = rows <= 32 ? 32 : (rows <= 64 ? 64 : (rows <= 128 ? 128 : 256));
    return gemm_umma_multi(&o, 1, rows, 3 * K, bm ? bm : 128, bn ? bn : pick, st);
}

// src (rows, cols; fp32 or bf16; leading dimension ld_src) -> dst (cols, ld_dst) bf16 with dst[c][r] = src[r][c];
// columns rows..ld_dst-1 of dst are zero-filled (K padding of the transposed operand)
template <typename T>
__global__ void k_transpose_to_bf16(const T* __restrict__ src, int rows, int cols, long long ld_src,
                                    __nv_bfloat16* __restrict__ dst, int ld_dst) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? (float)src[(size_t)r * ld_src + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < ld_dst) dst[(size_t)c * ld_dst + r] = __float2bfloat16(tile[threadIdx.x][i]);
    }
}
template <typename T>
static int transpose_to_bf16_t(const T* src, int rows, int cols, long long ld_src, __nv_bfloat16* dst, int ld_dst, cudaStream_t st) {
    dim3 grid(cdiv(ld_dst, 32), cdiv(cols, 32));
    SRNN_LAUNCH((k_transpose_to_bf16<T>), grid, dim3(32, 8), 0, st, src, rows, cols, ld_src, dst, ld_dst);
    return SRNN_OK;
}
int transpose_to_bf16(const float* src, int rows, int cols, long long ld_src, __nv_bfloat16* dst, int ld_dst, cudaStream_t st) {
    return transpose_to_bf16_t(src, rows, cols, ld_src, dst, ld_dst, st);
}
int transpose_to_bf16(const __nv_bfloat16* src, int rows, int cols, long long ld_src, __nv_bfloat16* dst, int ld_dst, cudaStream_t st) {
    return transpose_to_bf16_t(src, rows, cols, ld_src, dst, ld_dst, st);
}

}  // namespace srnn
