// C-ABI entry points (include/srnn_b200.h): context, weight packing, Predictor.forward, Generator.__call__.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

namespace srnn {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int Arena::alloc(void** p, size_t bytes) {
    if (bytes < 256) bytes = 256;
    SRNN_CUDA(cudaMalloc(p, bytes));
    ptrs.push_back(*p);
    return SRNN_OK;
}
void Arena::release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
}

int ensure_ws(srnn_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return SRNN_OK;
    if (ctx->ws) {
        SRNN_CUDA(cudaDeviceSynchronize());
        SRNN_CUDA(cudaFree(ctx->ws));
        ctx->ws = nullptr;
        ctx->ws_bytes = 0;
    }
    bytes += bytes / 8;
    SRNN_CUDA(cudaMalloc(&ctx->ws, bytes));
    ctx->ws_bytes = bytes;
    return SRNN_OK;
}

// two-pass bump allocator over the scratch buffer: pass 1 (base == nullptr) sizes it, pass 2 hands out pointers
struct Bump {
    char* base;
    size_t off = 0;
    explicit Bump(void* b) : base((char*)b) {}
    template <typename T>
    T* take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? (T*)(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

static int check_ready(const srnn_ctx* ctx) {
    if (!ctx) return fail(SRNN_ERR_ARG, "null context");
    if (!ctx->packed) return fail(SRNN_ERR_STATE, "srnn_pack_weights has not been called");
    return SRNN_OK;
}

}  // namespace srnn

using namespace srnn;

extern "C" {

const char* srnn_last_error(void) { return g_err; }
int srnn_version(void) { return 100; }
int64_t srnn_launch_count(void) { return (int64_t)g_launches.load(); }
int64_t srnn_graph_reuse_count(const srnn_ctx* ctx) { return ctx ? (int64_t)ctx->gen_graph.reuses : 0; }

int srnn_create(const srnn_config* cfg, srnn_ctx** out) {
    if (!cfg || !out) return fail(SRNN_ERR_ARG, "null argument");
    if (cfg->n_tiers < 1 || cfg->n_tiers > SRNN_MAX_TIERS) return fail(SRNN_ERR_ARG, "n_tiers must be in [1,%d]", SRNN_MAX_TIERS);
    if (cfg->n_rnn < 1 || cfg->n_rnn > SRNN_MAX_RNN) return fail(SRNN_ERR_ARG, "n_rnn must be in [1,%d]", SRNN_MAX_RNN);
    if (cfg->q_levels != SRNN_Q) return fail(SRNN_ERR_UNSUPPORTED, "q_levels must be %d", SRNN_Q);
    if (cfg->dim < 1 || cfg->cond_dim < 1 || cfg->spk_dim < 1) return fail(SRNN_ERR_ARG, "dim/cond_dim/spk_dim must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(SRNN_ERR_CUDA, "no CUDA device: this library has no CPU fallback (%s)", cudaGetErrorString(e));
    }
    srnn_ctx* c = new srnn_ctx();
    c->cfg = *cfg;
    c->H = cfg->dim;
    c->FS0 = cfg->frame_sizes[0];
    int n = 1;
    for (int i = 0; i < cfg->n_tiers; ++i) {
        if (cfg->frame_sizes[i] < 1) { delete c; return fail(SRNN_ERR_ARG, "frame_sizes[%d] must be positive", i); }
        n *= cfg->frame_sizes[i];                       // np.cumprod(frame_sizes)  model.py:34
        c->tiers[i].fs = cfg->frame_sizes[i];
        c->tiers[i].n = n;
        c->tiers[i].top = (i == cfg->n_tiers - 1);      // only the top tier is conditioned  model.py:46-47
        c->tiers[i].kin = n + (c->tiers[i].top ? cfg->cond_dim + cfg->spk_dim : 0);
    }
    c->lookback = n;
    cudaGetDevice(&c->device);
    cudaDeviceGetAttribute(&c->n_sms, cudaDevAttrMultiProcessorCount, c->device);
    for (auto& e : c->ev_stage) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    {   // a few paths take stream-ordered scratch (cudaMallocAsync: the table fold-back of every training step, test hooks): keep
        // the device pool's memory across synchronisation points instead of returning it to the driver (default threshold 0),
        // or a host sync inside a training loop turns the next step's allocation into a cudaMalloc
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    *out = c;
    return SRNN_OK;
}

int srnn_destroy(srnn_ctx* ctx) {
    if (!ctx) return SRNN_OK;
    cudaDeviceSynchronize();
    ctx->weights.release();
    for (auto& e : ctx->ev_stage) if (e) cudaEventDestroy(e);
    if (ctx->gen_graph.exec) cudaGraphExecDestroy(ctx->gen_graph.exec);
    if (ctx->gen_graph.graph) cudaGraphDestroy(ctx->gen_graph.graph);
    if (ctx->ws) cudaFree(ctx->ws);
    delete ctx;
    return SRNN_OK;
}

int srnn_lookback(const srnn_ctx* ctx) { return ctx ? ctx->lookback : fail(SRNN_ERR_ARG, "null context"); }

int srnn_pack_weights(srnn_ctx* ctx, const srnn_params* P, void* stream) {
    if (!ctx || !P) return fail(SRNN_ERR_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const srnn_config& c = ctx->cfg;
    const int H = ctx->H, Q = ctx->Q, FS0 = ctx->FS0, L = c.n_rnn;
    if (!ctx->tbl) {   // first call: allocate the packed buffers
        for (int i = 0; i < c.n_tiers; ++i) {
            TierPacked& t = ctx->tiers[i];
            SRNN_TRY(ctx->weights.alloc((void**)&t.w_in, sizeof(float) * H * t.kin));
            SRNN_TRY(ctx->weights.alloc((void**)&t.w_in_t, sizeof(float) * H * t.kin));
            SRNN_TRY(ctx->weights.alloc((void**)&t.b_in, sizeof(float) * H));
            for (int l = 0; l < L; ++l) {
                SRNN_TRY(ctx->weights.alloc((void**)&t.w_ih[l], sizeof(float) * 3 * H * H));
                SRNN_TRY(ctx->weights.alloc((void**)&t.w_hh[l], sizeof(float) * 3 * H * H));
                SRNN_TRY(ctx->weights.alloc((void**)&t.b_ih[l], sizeof(float) * 3 * H));
                SRNN_TRY(ctx->weights.alloc((void**)&t.b_hh[l], sizeof(float) * 3 * H));
            }
            SRNN_TRY(ctx->weights.alloc((void**)&t.w_up, sizeof(float) * (size_t)t.fs * H * H));
            SRNN_TRY(ctx->weights.alloc((void**)&t.b_up, sizeof(float) * t.fs * H));
            SRNN_TRY(ctx->weights.alloc((void**)&t.h0, sizeof(float) * L * H));
        }
        SRNN_TRY(ctx->weights.alloc((void**)&ctx->w_hid, sizeof(float) * H * H));
        SRNN_TRY(ctx->weights.alloc((void**)&ctx->b_hid, sizeof(float) * H));
        SRNN_TRY(ctx->weights.alloc((void**)&ctx->w_out, sizeof(float) * Q * H));
        SRNN_TRY(ctx->weights.alloc((void**)&ctx->b_out, sizeof(float) * Q));
        SRNN_TRY(ctx->weights.alloc((void**)&ctx->lut, sizeof(float) * Q));
        SRNN_TRY(ctx->weights.alloc((void**)&ctx->tbl, sizeof(float) * (size_t)FS0 * Q * H));
        ctx->has_bf16 = (H % 64 == 0);
        if (ctx->has_bf16) {
            typedef __nv_bfloat16 bf;
            for (int i = 0; i < c.n_tiers; ++i) {
                TierPacked& t = ctx->tiers[i];
                for (int l = 0; l < L; ++l) {
                    SRNN_TRY(ctx->weights.alloc((void**)&t.w_ih16[l], sizeof(bf) * 3 * H * H));
                    SRNN_TRY(ctx->weights.alloc((void**)&t.w_hh16[l], sizeof(bf) * 3 * H * H));
                }
                SRNN_TRY(ctx->weights.alloc((void**)&t.w_up16, sizeof(bf) * (size_t)t.fs * H * H));
                for (int l = 0; l < L; ++l) {
                    SRNN_TRY(ctx->weights.alloc((void**)&t.w_ih16_t[l], sizeof(bf) * 3 * H * H));
                    SRNN_TRY(ctx->weights.alloc((void**)&t.w_hh16_t[l], sizeof(bf) * 3 * H * H));
                }
                SRNN_TRY(ctx->weights.alloc((void**)&t.w_up16_t, sizeof(bf) * (size_t)t.fs * H * H));
            }
            SRNN_TRY(ctx->weights.alloc((void**)&ctx->w_hid16_t, sizeof(bf) * H * H));
            SRNN_TRY(ctx->weights.alloc((void**)&ctx->w_out16_t, sizeof(bf) * Q * H));
            SRNN_TRY(ctx->weights.alloc((void**)&ctx->w_hid16, sizeof(bf) * H * H));
            SRNN_TRY(ctx->weights.alloc((void**)&ctx->w_out16, sizeof(bf) * Q * H));
            SRNN_TRY(ctx->weights.alloc((void**)&ctx->tbl16, sizeof(bf) * (size_t)FS0 * Q * H));
        }
    }
    // scratch: folded conv weights before re-layout
    int max_fs = 1, max_kin = 1;
    for (int i = 0; i < c.n_tiers; ++i) {
        if (ctx->tiers[i].fs > max_fs) max_fs = ctx->tiers[i].fs;
        if (ctx->tiers[i].kin > max_kin) max_kin = ctx->tiers[i].kin;
    }
    size_t need = 0;
    float *t_in = nullptr, *t_c = nullptr, *t_s = nullptr, *t_up = nullptr, *t_mi = nullptr, *t_mit = nullptr;
    for (int pass = 0; pass < 2; ++pass) {
        Bump b(pass ? ctx->ws : nullptr);
        t_in = b.take<float>((size_t)H * max_kin);
        t_c = b.take<float>((size_t)H * c.cond_dim);
        t_s = b.take<float>((size_t)H * c.spk_dim);
        t_up = b.take<float>((size_t)H * H * max_fs);
        t_mi = b.take<float>((size_t)H * Q * FS0);
        t_mit = b.take<float>((size_t)H * Q * FS0);
        need = b.off;
        if (!pass) SRNN_TRY(ensure_ws(ctx, need));
    }
    for (int i = 0; i < c.n_tiers; ++i) {
        TierPacked& t = ctx->tiers[i];
        const srnn_tier_params& tp = P->tiers[i];
        if (!tp.h0 || !tp.input_expand.bias || !tp.upsampling.bias) return fail(SRNN_ERR_ARG, "tier %d: missing h0/bias", i);
        if (t.top) {
            if (!tp.cond_expand.bias || !tp.spk_expand.bias || !tp.spk_embedding) return fail(SRNN_ERR_ARG, "top tier: missing cond/spk parameters");
            SRNN_TRY(wn_fold(tp.input_expand, t_in, H, t.n, st));
            SRNN_TRY(wn_fold(tp.cond_expand, t_c, H, c.cond_dim, st));
            SRNN_TRY(wn_fold(tp.spk_expand, t_s, H, c.spk_dim, st));
            SRNN_TRY(pack_top_in(t_in, t_c, t_s, tp.spk_embedding, tp.input_expand.bias, tp.cond_expand.bias,
                                 tp.spk_expand.bias, t.w_in, t.b_in, H, t.n, c.cond_dim, c.spk_dim, st));
        } else {
            SRNN_TRY(wn_fold(tp.input_expand, t.w_in, H, t.n, st));
            SRNN_TRY(copy_f32(tp.input_expand.bias, t.b_in, H, st));
        }
        SRNN_TRY(transpose_f32(t.w_in, t.w_in_t, H, t.kin, st));
        for (int l = 0; l < L; ++l) {
            if (!tp.weight_ih[l] || !tp.weight_hh[l] || !tp.bias_ih[l] || !tp.bias_hh[l]) return fail(SRNN_ERR_ARG, "tier %d: missing GRU layer %d", i, l);
            if (!ctx->has_bf16) {      // with the tcgen05 copies these two ride along in the pack launch below
                SRNN_TRY(copy_f32(tp.weight_ih[l], t.w_ih[l], (size_t)3 * H * H, st));
                SRNN_TRY(copy_f32(tp.weight_hh[l], t.w_hh[l], (size_t)3 * H * H, st));
            }
            SRNN_TRY(copy_f32(tp.bias_ih[l], t.b_ih[l], (size_t)3 * H, st));
            SRNN_TRY(copy_f32(tp.bias_hh[l], t.b_hh[l], (size_t)3 * H, st));
        }
        SRNN_TRY(wn_fold(tp.upsampling, t_up, H, H * t.fs, st));       // norm over (out,k) per INPUT channel
        SRNN_TRY(pack_up(t_up, tp.upsampling.bias, t.w_up, t.b_up, H, t.fs, st));
        SRNN_TRY(copy_f32(tp.h0, t.h0, (size_t)L * H, st));
    }
    if (!P->embedding || !P->mlp_hidden.bias || !P->mlp_output.bias) return fail(SRNN_ERR_ARG, "missing MLP parameters");
    SRNN_TRY(wn_fold(P->mlp_input, t_mi, H, Q * FS0, st));
    SRNN_TRY(transpose_mlp_in(t_mi, t_mit, H, Q, FS0, st));
    // Tbl[j] (Q,H) = E (Q,Q) . W_in[:,:,j]^T for all FS0 taps in one batched launch
    SRNN_TRY(gemm_f32_batched(FS0, Q, H, Q, P->embedding, Q, 0, t_mit, Q, (long long)H * Q, ctx->tbl, H, (long long)Q * H, st));
    SRNN_TRY(wn_fold(P->mlp_hidden, ctx->w_hid, H, H, st));
    SRNN_TRY(copy_f32(P->mlp_hidden.bias, ctx->b_hid, H, st));
    SRNN_TRY(wn_fold(P->mlp_output, ctx->w_out, Q, H, st));
    SRNN_TRY(copy_f32(P->mlp_output.bias, ctx->b_out, Q, st));
    SRNN_TRY(build_lut(ctx->lut, Q, c.ulaw, st));
    if (ctx->has_bf16) {   // bf16 operand copies (plain + transposed) for the tcgen05 path: one launch for all of them
        PackBf16Item it[PACK_BF16_MAX];
        int n = 0;
        auto add = [&](const float* src, __nv_bfloat16* d, __nv_bfloat16* dt, int rows, int cols, float* d32 = nullptr) {
            if (n < PACK_BF16_MAX) it[n] = PackBf16Item{src, d, dt, d32, rows, cols, 0};
            ++n;
        };
        for (int i = 0; i < c.n_tiers; ++i) {
            TierPacked& t = ctx->tiers[i];
            for (int l = 0; l < L; ++l) {
                add(P->tiers[i].weight_ih[l], t.w_ih16[l], t.w_ih16_t[l], 3 * H, H, t.w_ih[l]);
                add(P->tiers[i].weight_hh[l], t.w_hh16[l], t.w_hh16_t[l], 3 * H, H, t.w_hh[l]);
            }
            add(t.w_up, t.w_up16, t.w_up16_t, t.fs * H, H);
        }
        add(ctx->w_hid, ctx->w_hid16, ctx->w_hid16_t, H, H);
        add(ctx->w_out, ctx->w_out16, ctx->w_out16_t, Q, H);
        add(ctx->tbl, ctx->tbl16, nullptr, FS0 * Q, H);
        if (n > PACK_BF16_MAX) return fail(SRNN_ERR_UNSUPPORTED, "too many weight matrices for one pack launch (%d)", n);
        SRNN_TRY(pack_bf16_multi(it, n, st));
    }
    ctx->packed = true;
    ctx->x3_valid = false;
    ctx->gi_fold_valid = false;
    return SRNN_OK;
}

// Generation-time fold of every tier's input expansion into its first GRU layer (TierPacked::g_in_t, b_gi0), (re)built on the
// first bf16 generation call after srnn_pack_weights (training re-packs every step and never needs it).
static int ensure_gi_fold(srnn_ctx* ctx, cudaStream_t st) {
    if (ctx->gi_fold_valid) return SRNN_OK;
    const int H = ctx->H;
    for (int i = 0; i < ctx->cfg.n_tiers; ++i) {
        TierPacked& t = ctx->tiers[i];
        if (!t.g_in_t) {
            SRNN_TRY(ctx->weights.alloc((void**)&t.g_in_t, sizeof(float) * (size_t)t.kin * 3 * H));
            SRNN_TRY(ctx->weights.alloc((void**)&t.b_gi0, sizeof(float) * 3 * H));
        }
        // G^T (kin, 3H) = W_in^T (kin, H) . W_ih0^T;   b_gi0 (1, 3H) = b_in (1, H) . W_ih0^T + b_ih0
        SRNN_TRY(gemm_f32(t.kin, 3 * H, H, t.w_in_t, H, t.w_ih[0], H, nullptr, nullptr, 0, 0, t.g_in_t, 3 * H, st));
        SRNN_TRY(gemm_f32(1, 3 * H, H, t.b_in, H, t.w_ih[0], H, t.b_ih[0], nullptr, 0, 0, t.b_gi0, 3 * H, st));
        if (!t.top) {          // first frame through the upsampling of the tier above
            const TierPacked& u = ctx->tiers[i + 1];
            if (!t.gup0_16) {
                SRNN_TRY(ctx->weights.alloc((void**)&t.gup0_16, sizeof(__nv_bfloat16) * 3 * (size_t)H * H));
                SRNN_TRY(ctx->weights.alloc((void**)&t.b_gup0, sizeof(float) * 3 * H));
            }
            float *wt = nullptr, *g32 = nullptr;
            SRNN_CUDA(cudaMallocAsync((void**)&wt, sizeof(float) * (size_t)H * H, st));
            SRNN_CUDA(cudaMallocAsync((void**)&g32, sizeof(float) * 3 * (size_t)H * H, st));
            // W_up[0] is rows 0 .. H-1 of w_up, (o, k) row-major; the fp32 GEMM contracts over the second index of both operands
            int rc = transpose_f32(u.w_up, wt, H, H, st);                                                     // (k, o)
            if (rc == SRNN_OK) rc = gemm_f32(3 * H, H, H, t.w_ih[0], H, wt, H, nullptr, nullptr, 0, 0, g32, H, st);   // (g, k)
            if (rc == SRNN_OK) rc = f32_to_bf16_pad(g32, 3 * H, H, H, t.gup0_16, 3 * H, H, st);
            if (rc == SRNN_OK) rc = gemm_f32(1, 3 * H, H, u.b_up, H, t.w_ih[0], H, t.b_gi0, nullptr, 0, 0, t.b_gup0, 3 * H, st);
            cudaFreeAsync(wt, st);
            cudaFreeAsync(g32, st);
            SRNN_TRY(rc);
        }
    }
    ctx->gi_fold_valid = true;
    return SRNN_OK;
}

// SRNN_MODE_BF16X3: split-bf16 copies [hi | lo | hi] of the dense weights, (re)built on the first use after srnn_pack_weights
static int ensure_x3(srnn_ctx* ctx, cudaStream_t st) {
    if (ctx->x3_valid) return SRNN_OK;
    if (!ctx->has_bf16) return fail(SRNN_ERR_UNSUPPORTED, "split-bf16 tensor-core mode needs dim %% 64 == 0 (dim=%d)", ctx->H);
    typedef __nv_bfloat16 bf;
    const srnn_config& c = ctx->cfg;
    const int H = ctx->H, Q = ctx->Q, L = c.n_rnn;
    if (!ctx->w_hid3) {
        for (int i = 0; i < c.n_tiers; ++i) {
            TierPacked& t = ctx->tiers[i];
            for (int l = 0; l < L; ++l) {
                SRNN_TRY(ctx->weights.alloc((void**)&t.w_ih3[l], sizeof(bf) * 9 * H * H));
                SRNN_TRY(ctx->weights.alloc((void**)&t.w_hh3[l], sizeof(bf) * 9 * H * H));
            }
            SRNN_TRY(ctx->weights.alloc((void**)&t.w_up3, sizeof(bf) * 3 * (size_t)t.fs * H * H));
        }
        SRNN_TRY(ctx->weights.alloc((void**)&ctx->w_hid3, sizeof(bf) * 3 * H * H));
        SRNN_TRY(ctx->weights.alloc((void**)&ctx->w_out3, sizeof(bf) * 3 * Q * H));
    }
    for (int i = 0; i < c.n_tiers; ++i) {
        TierPacked& t = ctx->tiers[i];
        for (int l = 0; l < L; ++l) {
            SRNN_TRY(split3_bf16(t.w_ih[l], 3 * H, H, H, t.w_ih3[l], 1, st));
            SRNN_TRY(split3_bf16(t.w_hh[l], 3 * H, H, H, t.w_hh3[l], 1, st));
        }
        SRNN_TRY(split3_bf16(t.w_up, (long long)t.fs * H, H, H, t.w_up3, 1, st));
    }
    SRNN_TRY(split3_bf16(ctx->w_hid, H, H, H, ctx->w_hid3, 1, st));
    SRNN_TRY(split3_bf16(ctx->w_out, Q, H, H, ctx->w_out3, 1, st));
    ctx->x3_valid = true;
    return SRNN_OK;
}

// ------------------------------------------------------------------------------------------------
// Predictor.forward  (model.py:357-436)
// ------------------------------------------------------------------------------------------------
// lay the forward buffers out in the context scratch (two-pass bump allocation); returns the bytes used
static int plan_forward(srnn_ctx* ctx, int B, int T, int mode) {
    const srnn_config& c = ctx->cfg;
    const int H = ctx->H, NT = c.n_tiers, NL = c.n_rnn, lookback = ctx->lookback;
    const int Lseq = lookback + T - 1;
    const bool bf16 = mode == SRNN_MODE_BF16;
    typedef __nv_bfloat16 bf;
    FwdPlan& P = ctx->fwd;
    for (int pass = 0; pass < 2; ++pass) {
        Bump b(pass ? ctx->ws : nullptr);
        P.seq = b.take<uint8_t>((size_t)B * Lseq);
        for (int i = 0; i < NT; ++i) {
            const TierPacked& t = ctx->tiers[i];
            const size_t M = (size_t)B * (T / t.n);
            P.A[i] = b.take<float>(M * t.kin);
            P.X[i] = b.take<float>(M * H);
            P.X16[i] = b.take<bf>(bf16 ? M * H : 1);
            for (int l = 0; l < NL; ++l) {
                P.GI[i][l] = b.take<float>(M * 3 * H);
                P.GH[i][l] = b.take<float>(M * 3 * H);
                P.Y[i][l] = b.take<float>(M * H);
                P.Y16[i][l] = b.take<bf>(bf16 ? M * H : 1);
            }
            P.H0[i] = b.take<float>((size_t)NL * B * H);
            P.H016[i] = b.take<bf>(bf16 ? (size_t)NL * B * H : 1);
            const bool up16 = bf16 && i == 0 && H % 64 == 0 && H <= 2048 && 256 % (H / 8) == 0 && !getenv("SRNN_UP_F32");   // tier 0 feeds only the table gather
            P.UP[i] = b.take<float>(up16 ? 1 : M * t.fs * H);
            if (i == 0) P.UP16 = up16 ? b.take<bf>(M * t.fs * H) : nullptr;
        }
        P.X1 = b.take<float>(bf16 ? 1 : (size_t)B * T * H);
        P.X2 = b.take<float>(bf16 ? 1 : (size_t)B * T * H);
        P.X1h = b.take<bf>(bf16 ? (size_t)B * T * H : 1);
        P.X2h = b.take<bf>(bf16 ? (size_t)B * T * H : 1);
        P.S3 = b.take<bf>(mode == SRNN_MODE_BF16X3 ? (size_t)B * T * 3 * H : 1);
        P.bytes = b.off;
        // the backward pass works in the scratch right behind the saved activations: reserve it now, because growing the
        // scratch later would free them
        if (!pass) SRNN_TRY(ensure_ws(ctx, b.off + (bf16 ? backward_scratch_bytes_bf16(ctx, B, T) : backward_scratch_bytes(ctx, B, T))));
    }
    P.B = B;
    P.T = T;
    P.mode = mode;
    return SRNN_OK;
}

static inline int pick_bn(int rows) { return rows <= 32 ? 32 : (rows <= 64 ? 64 : (rows <= 128 ? 128 : 256)); }

// teacher-forced contraction out (rows, n_feat) = act (rows, K) . W (n_feat, K)^T + bias [relu]: ROWS orientation (vector
// epilogue) when there are enough rows to fill 128-row tiles, swap-AB otherwise
static int tf_gemm(const __nv_bfloat16* W, int n_feat, const __nv_bfloat16* act, int rows, int K, const float* bias,
                   float* out_f32, __nv_bfloat16* out_bf16, int ld_out, int relu, cudaStream_t st) {
    GemmOperands o{W, act, bias, nullptr, out_f32, out_bf16, n_feat, K, K, 0, ld_out, relu, nullptr};
    if (rows >= 256 && ld_out % 8 == 0 && n_feat % 16 == 0) return gemm_umma_rows(o, rows, K, 1, nullptr, st);
    return gemm_umma_multi(&o, 1, rows, K, 128, pick_bn(rows), st);
}

int srnn_predict_fwd(srnn_ctx* ctx, int32_t B, int32_t T, const int64_t* input_seq, const void* cond,
                     int32_t cond_is_f64, const int64_t* spk, float* const* hidden_io, int32_t reset_mask,
                     float* logp_out, int32_t mode, void* stream) {
    SRNN_TRY(check_ready(ctx));
    if (!input_seq || !cond || !spk || !hidden_io || !logp_out) return fail(SRNN_ERR_ARG, "null argument");
    if (B < 1 || T < 1 || T % ctx->lookback) return fail(SRNN_ERR_ARG, "T=%d must be a positive multiple of lookback=%d", T, ctx->lookback);
    if (mode != SRNN_MODE_FP32 && mode != SRNN_MODE_BF16 && mode != SRNN_MODE_BF16X3)
        return fail(SRNN_ERR_UNSUPPORTED, "predict_fwd: mode %d not available", mode);
    if (mode != SRNN_MODE_FP32 && !ctx->has_bf16)
        return fail(SRNN_ERR_UNSUPPORTED, "bf16 tensor-core mode needs dim %% 64 == 0 (dim=%d)", ctx->H);
    cudaStream_t st = (cudaStream_t)stream;
    const srnn_config& c = ctx->cfg;
    const int H = ctx->H, Q = ctx->Q, NT = c.n_tiers, NL = c.n_rnn, lookback = ctx->lookback;
    const int Lseq = lookback + T - 1;
    const bool bf16 = mode == SRNN_MODE_BF16, x3 = mode == SRNN_MODE_BF16X3;
    ctx->fwd.valid = false;
    if (x3) SRNN_TRY(ensure_x3(ctx, st));
    SRNN_TRY(plan_forward(ctx, B, T, mode));
    FwdPlan& P = ctx->fwd;

    SRNN_TRY(i64_to_u8(input_seq, P.seq, (size_t)B * Lseq, st));
    const float* upper = nullptr;
    for (int i = NT - 1; i >= 0; --i) {                                     // model.py:378 top tier first
        const TierPacked& t = ctx->tiers[i];
        const int F = T / t.n, M = B * F;
        if (!hidden_io[i]) return fail(SRNN_ERR_ARG, "hidden_io[%d] is null", i);
        SRNN_TRY(frame_input(P.seq, Lseq, lookback - t.n, nullptr, t.n, B, F, cond, cond_is_f64, B, T / lookback, spk,
                             c.cond_dim, c.spk_dim, ctx->lut, P.A[i], t.kin, t.top, st));
        SRNN_TRY(gemm_f32(M, H, t.kin, P.A[i], t.kin, t.w_in, t.kin, t.b_in, upper, H, 0, P.X[i], H, st,
                          bf16 ? P.X16[i] : nullptr));
        const float* in = P.X[i];
        const __nv_bfloat16* in16 = P.X16[i];
        for (int l = 0; l < NL; ++l) {
            float* Y = P.Y[i][l];
            __nv_bfloat16* Y16 = P.Y16[i][l];
            float* hid = hidden_io[i] + (size_t)l * B * H;
            float* h0 = P.H0[i] + (size_t)l * B * H;                        // the initial state this pass used (for BPTT)
            __nv_bfloat16* h016 = P.H016[i] + (size_t)l * B * H;
            if ((reset_mask >> i) & 1) SRNN_TRY(bcast_rows(t.h0 + (size_t)l * H, h0, B, H, st));   // model.py:222-228
            else SRNN_TRY(copy_f32(hid, h0, (size_t)B * H, st));
            float* GI = P.GI[i][l];
            float* GH = P.GH[i][l];
            if (bf16) {
                SRNN_TRY(f32_to_bf16_pad(h0, B, H, H, h016, B, H, st));
                SRNN_TRY(tf_gemm(t.w_ih16[l], 3 * H, in16, M, H, t.b_ih[l], GI, nullptr, 3 * H, 0, st));
            } else if (x3) {
                SRNN_TRY(gemm_x3(M, 3 * H, H, in, H, t.w_ih3[l], t.b_ih[l], 0, GI, 3 * H, P.S3, 0, 0, st));
            } else {
                SRNN_TRY(gemm_f32(M, 3 * H, H, in, H, t.w_ih[l], H, t.b_ih[l], nullptr, 0, 0, GI, 3 * H, st));
            }
            if (bf16 && gru_persist_supported(B, H, ctx->n_sms)) {      // all F frames of the layer in one persistent launch
                if (!ctx->gru_ctr) SRNN_TRY(ctx->weights.alloc((void**)&ctx->gru_ctr, 256));
                SRNN_TRY(gru_persist_fwd(B, F, H, GI, t.w_hh16[l], t.b_hh[l], h0, h016, GH, Y, Y16, hid, ctx->gru_ctr, st));
                in = Y;
                in16 = Y16;
                continue;
            }
            for (int f = 0; f < F; ++f) {
                const float* hp = f ? Y + (size_t)(f - 1) * H : h0;
                const int hp_ld = f ? F * H : H;
                float* gh = GH + (size_t)f * 3 * H;
                if (bf16) {
                    const __nv_bfloat16* hp16 = f ? Y16 + (size_t)(f - 1) * H : h016;
                    SRNN_TRY(gemm_umma(t.w_hh16[l], 3 * H, hp16, B, H, H, hp_ld, t.b_hh[l], nullptr, 0, gh, nullptr,
                                       F * 3 * H, 0, 128, pick_bn(B), st));
                } else if (x3) {
                    SRNN_TRY(gemm_x3(B, 3 * H, H, hp, hp_ld, t.w_hh3[l], t.b_hh[l], 0, gh, F * 3 * H, P.S3, 0, 0, st));
                } else {
                    SRNN_TRY(gemm_f32(B, 3 * H, H, hp, hp_ld, t.w_hh[l], H, t.b_hh[l], nullptr, 0, 0, gh, F * 3 * H, st));
                }
                SRNN_TRY(gru_gates(GI + (size_t)f * 3 * H, F * 3 * H, gh, F * 3 * H, hp, hp_ld, Y + (size_t)f * H, F * H,
                                   f == F - 1 ? hid : nullptr, B, H, st, bf16 ? Y16 + (size_t)f * H : nullptr,
                                   F * H));                                 // carry: model.py:348
            }
            in = Y;
            in16 = Y16;
        }
        if (bf16)
            SRNN_TRY(tf_gemm(t.w_up16, t.fs * H, in16, M, H, t.b_up, (i == 0 && P.UP16) ? nullptr : P.UP[i],
                             (i == 0 && P.UP16) ? P.UP16 : nullptr, t.fs * H, 0, st));
        else if (x3)
            SRNN_TRY(gemm_x3(M, t.fs * H, H, in, H, t.w_up3, t.b_up, 0, P.UP[i], t.fs * H, P.S3, 0, 0, st));
        else
            SRNN_TRY(gemm_f32(M, t.fs * H, H, in, H, t.w_up, H, t.b_up, nullptr, 0, 0, P.UP[i], t.fs * H, st));
        upper = P.UP[i];
    }
    // sample-level MLP  (model.py:422-436, 308-325)
    const int FS0 = ctx->FS0, R = B * T;
    if (bf16) {
        SRNN_TRY(mlp_gather_bf16(P.seq, Lseq, lookback - FS0, nullptr, ctx->tbl16, upper, (long long)T * H, H, P.X1h, B, T, H,
                                 FS0, st, P.UP16));
        SRNN_TRY(tf_gemm(ctx->w_hid16, H, P.X1h, R, H, ctx->b_hid, nullptr, P.X2h, H, 1, st));
        SRNN_TRY(tf_gemm(ctx->w_out16, Q, P.X2h, R, H, ctx->b_out, logp_out, nullptr, Q, 0, st));
    } else {
        SRNN_TRY(mlp_gather(P.seq, Lseq, lookback - FS0, nullptr, ctx->tbl, upper, (long long)T * H, H, P.X1, B, T, H, FS0, st));
        if (x3) {
            SRNN_TRY(gemm_x3(R, H, H, P.X1, H, ctx->w_hid3, ctx->b_hid, 1, P.X2, H, P.S3, 0, 0, st));
            SRNN_TRY(gemm_x3(R, Q, H, P.X2, H, ctx->w_out3, ctx->b_out, 0, logp_out, Q, P.S3, 0, 0, st));
        } else {
            SRNN_TRY(gemm_f32(R, H, H, P.X1, H, ctx->w_hid, H, ctx->b_hid, nullptr, 0, 1, P.X2, H, st));
            SRNN_TRY(gemm_f32(R, Q, H, P.X2, H, ctx->w_out, H, ctx->b_out, nullptr, 0, 0, logp_out, Q, st));
        }
    }
    SRNN_TRY(logsoftmax_rows(logp_out, R, st));
    P.reset_mask = reset_mask;
    P.cond = cond;
    P.cond_is_f64 = cond_is_f64;
    P.spk = spk;
    P.valid = true;
    return SRNN_OK;
}

// Backward of srnn_predict_fwd (the autograd the reference gets from torch, trainer/__init__.py:103).  Must follow the
// forward call it differentiates on the same context.  logp = that call's output, dlogp = dL/dlogp (B,T,Q);
// params = the same raw tensors given to srnn_pack_weights; grads = same struct of WRITABLE gradient buffers
// (entries may be null to skip).  Every non-null gradient is overwritten, not accumulated.
int srnn_predict_bwd(srnn_ctx* ctx, const float* logp, const float* dlogp, const srnn_params* params,
                     const srnn_params* grads, void* stream) {
    SRNN_TRY(check_ready(ctx));
    if (!logp || !dlogp || !params || !grads) return fail(SRNN_ERR_ARG, "null argument");
    if (!ctx->fwd.valid) return fail(SRNN_ERR_STATE, "srnn_predict_bwd needs a preceding srnn_predict_fwd on this context");
    const int rc = ctx->fwd.mode != SRNN_MODE_BF16 ? predict_bwd_f32(ctx, logp, dlogp, params, grads, (cudaStream_t)stream)
                                                   : predict_bwd_bf16(ctx, logp, dlogp, params, grads, (cudaStream_t)stream);
    ctx->fwd.valid = false;
    return rc;
}

// Backward of sequence_nll_loss_bits(srnn_predict_fwd(...), target) in one pass (nn.py:66-70 + trainer/__init__.py:102-103):
// the loss gradient is folded into the log-softmax backward, dL/dlogits = (exp(logp) - onehot(target)) * g * log2(e) / (B*T),
// so no dense dL/dlogp exists.  target (B, T) int64; gscale = device pointer to the upstream scalar gradient (null = 1).
int srnn_predict_bwd_nll(srnn_ctx* ctx, const float* logp, const int64_t* target, const float* gscale,
                         const srnn_params* params, const srnn_params* grads, void* stream) {
    SRNN_TRY(check_ready(ctx));
    if (!logp || !target || !params || !grads) return fail(SRNN_ERR_ARG, "null argument");
    if (!ctx->fwd.valid) return fail(SRNN_ERR_STATE, "srnn_predict_bwd_nll needs a preceding srnn_predict_fwd on this context");
    const int rc = ctx->fwd.mode != SRNN_MODE_BF16
                       ? predict_bwd_f32(ctx, logp, nullptr, params, grads, (cudaStream_t)stream, target, gscale)
                       : predict_bwd_bf16(ctx, logp, nullptr, params, grads, (cudaStream_t)stream, target, gscale);
    ctx->fwd.valid = false;
    return rc;
}

// Data-parallel overlap: make `stream` wait until the last srnn_predict_bwd on this context has finished every gradient that
// does not belong to the top tier (sample-level MLP, embedding, lower tiers); the caller can then all-reduce those on
// `stream` while the top tier's backward pass is still running on the compute stream.
int srnn_bwd_wait_early(srnn_ctx* ctx, void* stream) {
    if (!ctx) return fail(SRNN_ERR_ARG, "null context");
    return srnn_bwd_wait_stage(ctx, 2 * (ctx->cfg.n_tiers - 1), stream);
}

// Finer-grained form: the backward pass finalises gradients in the order  stage 0 = sample-level MLP + embedding,
// stage 1 + 2i = tier i's upsampling (conv_t weight_g / weight_v, bias), stage 2 + 2i = the rest of tier i (lowest tier first).
// `stream` waits for stage `stage` of the last srnn_predict_bwd[_nll] on this context.
int srnn_bwd_wait_stage(srnn_ctx* ctx, int32_t stage, void* stream) {
    if (!ctx) return fail(SRNN_ERR_ARG, "null context");
    if (stage < 0 || stage > 2 * ctx->cfg.n_tiers || !ctx->ev_stage[stage]) return fail(SRNN_ERR_ARG, "bad stage %d", stage);
    SRNN_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, ctx->ev_stage[stage], 0));
    return SRNN_OK;
}

// optim.py:10-13 (element-wise clamp of every gradient to [-clamp, clamp]) fused with torch.optim.Adam's update
// (train.py:238; no weight decay / amsgrad) over `count` tensors in one launch.  step = 1, 2, ... (bias correction).
int srnn_clamp_adam_step(int32_t count, float* const* params, const float* const* grads, float* const* exp_avg,
                         float* const* exp_avg_sq, const int64_t* sizes, float lr, float beta1, float beta2, float eps,
                         int32_t step, float clamp, void* stream) {
    if (count < 0 || (count && (!params || !grads || !exp_avg || !exp_avg_sq || !sizes))) return fail(SRNN_ERR_ARG, "bad argument");
    if (step < 1) return fail(SRNN_ERR_ARG, "step must be >= 1");
    return clamp_adam(count, params, grads, exp_avg, exp_avg_sq, (const long long*)sizes, lr, beta1, beta2, eps, step, clamp,
                      (cudaStream_t)stream);
}

// The same with every gradient multiplied by grad_scale before the clamp: data-parallel training passes 1 / world_size after
// the sum all-reduce, so "mean over ranks -> clamp -> Adam" (optim.py:10-13 on the full-batch gradient) is one pass.
int srnn_clamp_adam_step_scaled(int32_t count, float* const* params, const float* const* grads, float* const* exp_avg,
                                float* const* exp_avg_sq, const int64_t* sizes, float lr, float beta1, float beta2, float eps,
                                int32_t step, float clamp, float grad_scale, void* stream) {
    if (count < 0 || (count && (!params || !grads || !exp_avg || !exp_avg_sq || !sizes))) return fail(SRNN_ERR_ARG, "bad argument");
    if (step < 1) return fail(SRNN_ERR_ARG, "step must be >= 1");
    return clamp_adam(count, params, grads, exp_avg, exp_avg_sq, (const long long*)sizes, lr, beta1, beta2, eps, step, clamp,
                      (cudaStream_t)stream, grad_scale);
}

// Measurement hook: accumulated CUDA-event time (ms) and launch count of the persistent sample-level kernel over the last
// srnn_generate call that ran with the environment variable SRNN_TIME_KERNELS set (which issues the launches directly
// instead of through the CUDA graph so that events can bracket them on their stream).
int srnn_timed_kernel(const srnn_ctx* ctx, double* ms, int64_t* launches) {
    if (!ctx || !ms || !launches) return fail(SRNN_ERR_ARG, "null argument");
    *ms = ctx->timed_ms;
    *launches = ctx->timed_launches;
    return SRNN_OK;
}

// Bottle-neck conditioner chain (thesis-derived, parity unpinned; see include/srnn_b200.h): fold weight-norm, run the k = 1 chain.
int srnn_cond_chain_fwd(const srnn_cond_chain* chain, const float* cond, int32_t rows, float* out, float* scratch, void* stream) {
    if (!chain || !cond || !out || !scratch || rows < 0) return fail(SRNN_ERR_ARG, "null argument");
    if (chain->n_layers < 1 || chain->n_layers > SRNN_MAX_CHAIN) return fail(SRNN_ERR_ARG, "bad chain length");
    const float* w[SRNN_MAX_CHAIN];
    const float* b[SRNN_MAX_CHAIN];
    size_t off = 0;
    for (int l = 0; l < chain->n_layers; ++l) {
        const int din = chain->dims[l], dout = chain->dims[l + 1];
        if (din < 1 || dout < 1) return fail(SRNN_ERR_ARG, "bad chain width");
        SRNN_TRY(wn_fold(chain->layers[l], scratch + off, dout, din, (cudaStream_t)stream));
        w[l] = scratch + off;
        b[l] = chain->layers[l].bias;
        off += (size_t)din * dout;
    }
    return cond_chain_fwd(chain->n_layers, chain->dims, w, b, cond, rows, out, (cudaStream_t)stream);
}

// Name of the persistent sample-level kernel the last bf16 srnn_generate call of this process ran (bench.py's roofline line).
namespace srnn { const char* g_sample_kernel = "k_mlp_persist"; }
const char* srnn_sample_kernel_name(void) { return srnn::g_sample_kernel; }

// mean NLL in bits of log-probabilities against targets (nn.py:66-70); loss_out = one device float
int srnn_nll_loss_bits(srnn_ctx* ctx, const float* logp, const int64_t* target, int32_t rows, float* loss_out, void* stream) {
    if (!ctx || !logp || !target || !loss_out || rows < 1) return fail(SRNN_ERR_ARG, "bad argument");
    static const int NP = 256;
    if (!ctx->loss_partial) SRNN_TRY(ctx->weights.alloc((void**)&ctx->loss_partial, sizeof(float) * NP));
    return nll_bits(logp, target, rows, ctx->loss_partial, NP, loss_out, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// Generator.__call__  (model.py:445-520): one CUDA graph per top-tier period (lookback samples), replayed n_cond
// times.  fp32 mode: FFMA GEMMs.  bf16 mode: the same schedule with every H-wide contraction on tcgen05 (gemm_umma).
// ------------------------------------------------------------------------------------------------
// cluster_rows: 0 = k_mlp_persist (16-CTA row groups over L2), 16 / 24 = k_mlp_cluster (8-CTA clusters, rows per cluster)
static int generate_graph(srnn_ctx* ctx, bool bf16, bool persist, int B, int n_cond, const float* cond, int cond_rows,
                          const int64_t* spk, const float* uniforms, int u_ld, uint8_t* samples_out, float* audio_out,
                          float* logp_out, cudaStream_t user, int cluster_rows = 0, bool x3 = false) {
    const srnn_config& c = ctx->cfg;
    const int H = ctx->H, Q = ctx->Q, NT = c.n_tiers, NL = c.n_rnn, lookback = ctx->lookback, FS0 = ctx->FS0;
    const int T = n_cond * lookback, Lseq = lookback + T;
    typedef __nv_bfloat16 bf;
    uint8_t* seq = nullptr;
    int* step_base = nullptr;
    float *hid[SRNN_MAX_TIERS], *A[SRNN_MAX_TIERS], *X[SRNN_MAX_TIERS], *GI[SRNN_MAX_TIERS], *GH[SRNN_MAX_TIERS],
        *OUT[SRNN_MAX_TIERS], *X1 = nullptr, *X2 = nullptr, *LG = nullptr;
    float* GHL[SRNN_MAX_TIERS][SRNN_MAX_RNN];      // per-layer recurrent projections of the fused-cell schedule
    bf *hid16[SRNN_MAX_TIERS], *X16[SRNN_MAX_TIERS], *X1h = nullptr, *X2h = nullptr, *S3 = nullptr;
    float *part = nullptr, *pcarry = nullptr, *GIP = nullptr;
    bf* UP16 = nullptr;
    const int fs_top = ctx->tiers[NT - 1].fs;
    unsigned* gctr = nullptr;
    const int RG = (B + 31) / 32, NS = H / 64;
    const int n_clusters = cluster_rows ? (B + cluster_rows - 1) / cluster_rows : 0;
    const size_t x1_rows = (size_t)RG * 32 > (size_t)n_clusters * cluster_rows ? (size_t)RG * 32 : (size_t)n_clusters * cluster_rows;
    const int sample_ctas = cluster_rows ? n_clusters * 8 : RG * NS;
    for (int pass = 0; pass < 2; ++pass) {
        Bump b(pass ? ctx->ws : nullptr);
        seq = b.take<uint8_t>((size_t)B * Lseq);
        step_base = b.take<int>(1);
        for (int i = 0; i < NT; ++i) {
            const TierPacked& t = ctx->tiers[i];
            hid[i] = b.take<float>((size_t)NL * B * H);
            A[i] = b.take<float>((size_t)B * t.kin);
            X[i] = b.take<float>((size_t)B * H);
            GI[i] = b.take<float>((size_t)B * 3 * H);
            GH[i] = b.take<float>((size_t)B * 3 * H);
            for (int l = 0; l < NL; ++l) GHL[i][l] = b.take<float>(bf16 ? (size_t)B * 3 * H : 1);
            OUT[i] = b.take<float>((size_t)B * t.fs * H);
            hid16[i] = b.take<bf>((size_t)NL * B * H);
            X16[i] = b.take<bf>((size_t)B * H);
        }
        X1 = b.take<float>((size_t)B * H);
        X2 = b.take<float>((size_t)B * H);
        LG = b.take<float>((size_t)B * Q);
        X1h = b.take<bf>(x1_rows * H);
        X2h = b.take<bf>((size_t)B * H);
        S3 = b.take<bf>(x3 ? (size_t)B * 3 * H : 1);
        part = b.take<float>(persist && !cluster_rows ? (size_t)RG * NS * 32 * Q : 1);
        gctr = b.take<unsigned>(2 * RG);
        pcarry = b.take<float>(cluster_rows ? 2 * x1_rows * H : 1);
        GIP = b.take<float>(bf16 && NT == 2 ? (size_t)B * fs_top * 3 * H : 1);    // input-fold schedule: W_ih0 upper + b_gi0 of tier 0
        UP16 = b.take<bf>(bf16 && NT == 2 ? (size_t)B * fs_top * H : 1);          // ... and the bf16 top-tier output it contracts
        if (!pass) SRNN_TRY(ensure_ws(ctx, b.off));
    }
    // private capture stream (the caller's stream may be the legacy default stream, which cannot be captured)
    cudaStream_t st;
    cudaEvent_t ev_in, ev_out;
    SRNN_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    SRNN_CUDA(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
    SRNN_CUDA(cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming));
    SRNN_CUDA(cudaEventRecord(ev_in, user));
    SRNN_CUDA(cudaStreamWaitEvent(st, ev_in, 0));

    SRNN_TRY(fill_u8(seq, (uint8_t)(Q / 2), (size_t)B * Lseq, st));                  // q_zero  model.py:459
    SRNN_CUDA(cudaMemsetAsync(step_base, 0, sizeof(int), st));
    SRNN_TRY(add_int(step_base, lookback, st));                                      // i starts at lookback  model.py:462
    for (int i = 0; i < NT; ++i) {
        for (int l = 0; l < NL; ++l)                                                 // reset_hidden_states  model.py:451
            SRNN_TRY(bcast_rows(ctx->tiers[i].h0 + (size_t)l * H, hid[i] + (size_t)l * B * H, B, H, st));
        if (bf16) SRNN_TRY(f32_to_bf16_pad(hid[i], NL * B, H, H, hid16[i], NL * B, H, st));
    }
    // batch-row tile of the tier GEMMs (UMMA M = 128 features): one thread issues a tcgen05.mma every ~45+ cycles, so
    // only N = 256 tiles (128-cycle tensor floor) keep the tensor pipe rather than the issue slot busy
    const int bn_tier = B <= 32 ? 32 : (B <= 64 ? 64 : (B <= 128 ? 128 : 256));
    // With more than 128 utterances a 128-row tile doubles the CTA count (measured: 19.8 -> 16.3 us for the gi+gh launch at
    // B = 256) as long as the grid still fits one wave of SMs; the tier-2 upsampling (160 feature tiles) keeps 256-row tiles.
    const bool fused_cell = bf16 && gru_cell_gen_supported(H);
    const bool skip_tiers = getenv("SRNN_SKIP_TIERS") != nullptr;   // timing experiment only (results are wrong)
    const bool time_tiers_early = getenv("SRNN_TIME_TIERS") != nullptr;
    cudaStream_t st2 = nullptr;                 // side branch of the tier steps (captured into the same graph)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    if (fused_cell) {
        SRNN_CUDA(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
        SRNN_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        SRNN_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    // tile of the per-sample MLP GEMMs in the one-launch-per-contraction schedule: the instantiated shape with the most tiles
    // that still fit one wave of SMs (64 x 32 up to 288 utterances at dim 1024; 128 x 64 at 1024, 128 x 256 at 4096 -- with
    // 64 x 32 throughout, 4096 utterances meant 2048 tiny tiles per launch)
    auto mlp_tile = [&](int n_feat, int& bm, int& bn) {
        static const int cand[6][2] = {{64, 32}, {64, 64}, {128, 32}, {128, 64}, {128, 128}, {128, 256}};
        int best = -1, best_tiles = 0;
        for (int k = 0; k < 6; ++k) {
            const int tiles = cdiv(n_feat, cand[k][0]) * cdiv(B, cand[k][1]);
            if (tiles <= ctx->n_sms && tiles > best_tiles) { best = k; best_tiles = tiles; }
        }
        if (best < 0) best = 5;
        bm = cand[best][0];
        bn = cand[best][1];
    };
    int bm_hid = 64, bn_hid = 32, bm_out = 64, bn_out = 32;
    mlp_tile(H, bm_hid, bn_hid);
    mlp_tile(Q, bm_out, bn_out);
    auto bn_for = [&](int n_feat, int nprob) {
        if (bn_tier < 256) return bn_tier;
        return cdiv(n_feat, 128) * cdiv(B, 128) * nprob <= ctx->n_sms ? 128 : 256;
    };
    // Recurrent projections gh_l = W_hh_l h_l + b_hh_l of a tier's NEXT step depend only on the state its current step leaves
    // behind.  With the persistent sample kernel on RG x NS CTAs (128 at C2) the remaining SMs (20) are idle for the ~180 us
    // of a launch, so these GEMMs (96 tiles at C2) run there, capped to that many CTAs, beside the sample kernel instead of in
    // front of the next tier step ("shadow" schedule; SRNN_NO_SHADOW_GH=1 restores the fork beside the input expansion).
    int spare_sms = persist ? ctx->n_sms - sample_ctas : 0;
    if (cluster_rows && spare_sms > 24) spare_sms = 24;      // leave whole GPCs free: a cluster needs 8 free SMs of ONE GPC
    if (getenv("SRNN_SHADOW_CTAS") && atoi(getenv("SRNN_SHADOW_CTAS")) > 0 && atoi(getenv("SRNN_SHADOW_CTAS")) < spare_sms)
        spare_sms = atoi(getenv("SRNN_SHADOW_CTAS"));          // experiment: fewer CTAs for the shadow GEMMs
    const bool shadow_gh = bf16 && fused_cell && persist && spare_sms >= 8 && !skip_tiers && !getenv("SRNN_NO_SHADOW_GH");
    // consecutive kernels of a tier step (input expansion -> cells -> upsampling) are launched as programmatic dependents
    const bool pdl = bf16 && fused_cell && !time_tiers_early && !getenv("SRNN_NO_PDL");
    bool prev_tier_kernel = false;                 // the previous launch on st is one of those kernels
    // ... and so is the cluster sample kernel (no cooperative launch): as a dependent of the upsampling it sets up barriers and
    // TMEM and requests its first weight tiles while that kernel drains; the tier step after it is pre-launched from its last
    // sample step.  SRNN_NO_PDL_SAMPLE=1 restores plain launches on both sides.
    const bool pdl_sample = pdl && cluster_rows && !getenv("SRNN_NO_PDL_SAMPLE");
    bool prev_sample_kernel = false;
    cudaEvent_t ev_gh[SRNN_MAX_TIERS] = {};
    bool gh_todo[SRNN_MAX_TIERS] = {}, gh_pending[SRNN_MAX_TIERS] = {};
    auto launch_gh = [&](int i, cudaStream_t s, int cap) -> int {
        const TierPacked& t = ctx->tiers[i];
        gemm_umma_set_cta_cap(cap);
        int rc = SRNN_OK;
        for (int l = 0; l < NL && rc == SRNN_OK; l += 2) {
            const int np = NL - l >= 2 ? 2 : 1;
            GemmOperands ops[2];
            for (int q = 0; q < np; ++q)
                ops[q] = GemmOperands{t.w_hh16[l + q], hid16[i] + (size_t)(l + q) * B * H, t.b_hh[l + q], nullptr,
                                      GHL[i][l + q], nullptr, 3 * H, H, H, 0, 3 * H, 0, nullptr};
            rc = gemm_umma_multi(ops, np, B, H, 128, cap ? (B <= 128 ? bn_tier : 128) : bn_for(3 * H, np), s);
        }
        gemm_umma_set_cta_cap(0);
        return rc;
    };
    if (shadow_gh) {
        for (int i = 0; i < NT; ++i) {
            SRNN_CUDA(cudaEventCreateWithFlags(&ev_gh[i], cudaEventDisableTiming));
            SRNN_TRY(launch_gh(i, st, 0));                                           // first step: from h0, outside the graph
        }
    }
    // Same idea for the top tier's input expansion (k_tier_input_split): everything but the last tier-0 frame of the previous
    // period (and the conditioner / speaker columns) is accumulated beside the last sample launch of that period; the serial
    // path keeps K = FS0 columns.  XP (the partial sums) lives in the top tier's GI buffer, unused by the fused-cell schedule.
    const bool shadow_in = shadow_gh && !getenv("SRNN_NO_SHADOW_IN");
    // CTAs of the L2 weight prefetch beside each sample launch (0 = off)
    const int l2_prefetch = shadow_gh && getenv("SRNN_L2_PREFETCH") ? atoi(getenv("SRNN_L2_PREFETCH")) : 0;
    // Input fold (TierPacked::g_in_t): the input expansion of a tier is folded into its first GRU layer, gi_0 = G a + W_ih0 upper +
    // b_gi0.  Top tier: the whole of gi_0 is a K = kin contraction whose known columns are summed in the shadow (GP, width 3H
    // instead of XP, width H); k_gru_cell_lite adds the last FS0 sample columns and does the gate math -- the input kernel
    // and the first cell GEMM leave the serial path.  Tier 0 of a two-tier model: W_ih0 upper + b_gi0 for all fs_top frames of
    // the period in ONE tcgen05 GEMM behind the top tier's upsampling (which writes its output in bf16), then per frame the
    // same lite kernel over the frame's n sample columns.  SRNN_NO_GI_FOLD=1: separate input kernels + cell GEMMs.
    bool fold_ok = NT <= 2 && gru_cell_lite_supported(ctx->tiers[NT - 1].n - (ctx->tiers[NT - 1].n - FS0 > 0 ? ctx->tiers[NT - 1].n - FS0 : 0));
    if (NT == 2) fold_ok = fold_ok && gru_cell_lite_supported(ctx->tiers[0].n);
    const bool gi_fold = shadow_in && fold_ok && !getenv("SRNN_NO_GI_FOLD");
    if (gi_fold) SRNN_TRY(ensure_gi_fold(ctx, st));
    const TierPacked& ttop = ctx->tiers[NT - 1];
    float* XP = GI[NT - 1];                        // gi_fold: GP (B, 3H)
    const int xp_lo = ttop.n - FS0 > 0 ? ttop.n - FS0 : 0, xp_hi = ttop.n;
    cudaEvent_t ev_xp = nullptr, ev_gip = nullptr;
    bool xp_pending = false, gip_todo = false, gip_pending = false, up_top_todo = false;
    auto launch_xp = [&](int off, cudaStream_t s) -> int {
        if (gi_fold)
            return tier_input_split(false, seq, Lseq, off, step_base, ttop.n, B, cond, cond_rows, n_cond, spk, c.cond_dim, ctx->lut,
                                    ttop.g_in_t, ttop.b_gi0, XP, nullptr, nullptr, 3 * H, ttop.kin, xp_lo, xp_hi, s);
        return tier_input_split(false, seq, Lseq, off, step_base, ttop.n, B, cond, cond_rows, n_cond, spk, c.cond_dim, ctx->lut,
                                ttop.w_in_t, ttop.b_in, XP, nullptr, nullptr, H, ttop.kin, xp_lo, xp_hi, s);
    };
    if (shadow_in) {
        SRNN_CUDA(cudaEventCreateWithFlags(&ev_xp, cudaEventDisableTiming));
        SRNN_CUDA(cudaEventCreateWithFlags(&ev_gip, cudaEventDisableTiming));
        SRNN_TRY(launch_xp(-ttop.n, st));                                            // first period: the q_zero prefix
    }

    if (persist) SRNN_CUDA(cudaMemsetAsync(X1h, 0, sizeof(bf) * x1_rows * H, st));
    if (persist) SRNN_CUDA(cudaMemsetAsync(gctr, 0, sizeof(unsigned) * 2 * RG, st));     // group barrier counters: once per call
    const bool carry = cluster_rows && !getenv("SRNN_NO_PCARRY");
    if (carry) SRNN_TRY(mlp_cluster_carry_init(ctx->tbl16, FS0, H, (int)x1_rows, Q / 2, pcarry, st));
    long long* trace = nullptr;
    const size_t trace_n = (size_t)(cluster_rows ? 8 : 1) * FS0 * 64;
    if (persist && getenv("SRNN_TRACE")) {
        SRNN_CUDA(cudaMalloc((void**)&trace, sizeof(long long) * trace_n));
        SRNN_CUDA(cudaMemset(trace, 0, sizeof(long long) * trace_n));
    }
    const bool time_kernels = persist && getenv("SRNN_TIME_KERNELS");
    const bool time_tiers = getenv("SRNN_TIME_TIERS") != nullptr;   // development aid: per-kernel durations of the tier steps
    bool use_graph = !getenv("SRNN_NO_GRAPH") && !time_kernels && !time_tiers;
    std::vector<cudaEvent_t> tev;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    auto mark = [&](const char* name) {
        if (!time_tiers) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        marks.push_back({name, e});
    };
    const long long before = g_launches.load();
    // everything the captured kernels bake in: shapes, schedule switches, caller pointers and the scratch base
    srnn_ctx::GenGraph& gc = ctx->gen_graph;
    const unsigned long long gkey[] = {(unsigned long long)B, (unsigned long long)n_cond, (unsigned long long)cond_rows,
                                       (unsigned long long)u_ld, (unsigned long long)cluster_rows,
                                       (unsigned long long)(bf16 | persist << 1 | x3 << 2 | shadow_gh << 3 | shadow_in << 4 | pdl << 5 |
                                                            skip_tiers << 6 | fused_cell << 7 | pdl_sample << 8 | carry << 9 | gi_fold << 10 | (unsigned long long)l2_prefetch << 16),
                                       (unsigned long long)spare_sms, (unsigned long long)(uintptr_t)cond,
                                       (unsigned long long)(uintptr_t)spk, (unsigned long long)(uintptr_t)uniforms,
                                       (unsigned long long)(uintptr_t)samples_out, (unsigned long long)(uintptr_t)audio_out,
                                       (unsigned long long)(uintptr_t)logp_out, (unsigned long long)(uintptr_t)ctx->ws,
                                       (unsigned long long)(getenv("SRNN_MC_DBG") ? atoi(getenv("SRNN_MC_DBG")) : 0),
                                       (unsigned long long)((getenv("SRNN_GEMM_DEEP_RING") != nullptr) | (getenv("SRNN_GEMM_SINGLE_BUF") != nullptr) << 1 |
                                                            (getenv("SRNN_SWAP_PAIR") != nullptr) << 2)};
    const int ngkey = (int)(sizeof(gkey) / sizeof(gkey[0]));
    const bool cache_ok = use_graph && !trace && !getenv("SRNN_NO_GRAPH_CACHE");
    const bool cache_hit = cache_ok && gc.exec && gc.nkey == ngkey && !memcmp(gc.key, gkey, sizeof(gkey));
    if (use_graph && !cache_hit) SRNN_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    auto body = [&]() -> int {
        prev_tier_kernel = false;
        prev_sample_kernel = false;
        for (int pos = 0; pos < lookback; ++pos) {                                   // i = *step_base + pos
            for (int i = NT - 1; i >= 0; --i) {
                const TierPacked& t = ctx->tiers[i];
                if (pos % t.n) continue;                                             // model.py:465
                if (skip_tiers) continue;
                const float* upper = nullptr;
                int up_ld = 0;
                if (!t.top) {                                                        // model.py:491-495
                    const TierPacked& u = ctx->tiers[i + 1];
                    upper = OUT[i + 1] + (size_t)((pos / t.n) % u.fs) * H;
                    up_ld = u.fs * H;
                }
                mark(t.top ? "(start top)" : "(start)");
                if (bf16 && fused_cell && !shadow_gh) {      // fork: gh of every layer runs beside the input expansion (independent of it)
                    SRNN_CUDA(cudaEventRecord(ev_fork, st));
                    SRNN_CUDA(cudaStreamWaitEvent(st2, ev_fork, 0));
                    for (int l = 0; l < NL; l += 2) {
                        const int np = NL - l >= 2 ? 2 : 1;
                        GemmOperands ops[2];
                        for (int q = 0; q < np; ++q)
                            ops[q] = GemmOperands{t.w_hh16[l + q], hid16[i] + (size_t)(l + q) * B * H, t.b_hh[l + q], nullptr,
                                                  GHL[i][l + q], nullptr, 3 * H, H, H, 0, 3 * H, 0, nullptr};
                        SRNN_TRY(gemm_umma_multi(ops, np, B, H, 128, bn_for(3 * H, np), st2));
                    }
                    SRNN_CUDA(cudaEventRecord(ev_join, st2));
                }
                if (gi_fold) {
                    // folded first layer: [top: nothing | tier 0, first frame of the period: one GEMM for all its frames] -> lite cell
                    if (!t.top && (pos / t.n) % fs_top == 0) {
                        // W_ih0 upper + b_gi0 of the period's FIRST frame on the serial path; the other fs_top - 1 frames follow
                        // in the shadow of the next sample launch
                        g_pdl = pdl && prev_tier_kernel;
                        // (through the folded upsampling of the tier above: straight from its new state; that tier's upsampling
                        // itself runs in the shadow with the other frames)
                        const int rc_g = gemm_umma(t.gup0_16, 3 * H, hid16[i + 1] + (size_t)(NL - 1) * B * H, B, H, H, H, t.b_gup0,
                                                   nullptr, 0, GIP, nullptr, fs_top * 3 * H, 0, 128, bn_for(3 * H, 1), st);
                        g_pdl = 0;
                        SRNN_TRY(rc_g);
                        gip_todo = fs_top > 1;
                        mark("gi pre");
                    } else if (!t.top && gip_pending) {
                        SRNN_CUDA(cudaStreamWaitEvent(st, ev_gip, 0));
                        gip_pending = false;
                    }
                    if (gh_pending[i]) {                         // gh of this step ran beside the previous sample launch
                        SRNN_CUDA(cudaStreamWaitEvent(st, ev_gh[i], 0));
                        gh_pending[i] = false;
                    }
                    g_pdl = pdl && (prev_tier_kernel || (pdl_sample && prev_sample_kernel));
                    prev_sample_kernel = false;
                    const int rc_l = t.top
                        ? gru_cell_lite(B, H, XP, 3 * H, t.g_in_t, xp_lo, xp_hi, seq, Lseq, pos - t.n, step_base, ctx->lut, GHL[i][0],
                                        hid[i], hid16[i], st)
                        : gru_cell_lite(B, H, GIP + (size_t)((pos / t.n) % fs_top) * 3 * H, (long long)fs_top * 3 * H, t.g_in_t, 0, t.n,
                                        seq, Lseq, pos - t.n, step_base, ctx->lut, GHL[i][0], hid[i], hid16[i], st);
                    g_pdl = 0;
                    SRNN_TRY(rc_l);
                    prev_tier_kernel = true;
                    mark("cell lite");
                }
                const float* in = X[i];
                const bf* in16 = gi_fold ? hid16[i] : X16[i];
                if (!gi_fold) {
                    g_pdl = pdl && (prev_tier_kernel || (pdl_sample && prev_sample_kernel));
                    prev_sample_kernel = false;
                    int rc_in = (t.top && shadow_in)
                        ? tier_input_split(true, seq, Lseq, pos - t.n, step_base, t.n, B, cond, cond_rows, n_cond, spk, c.cond_dim,
                                           ctx->lut, t.w_in_t, t.b_in, XP, X[i], X16[i], H, t.kin, xp_lo, xp_hi, st)
                        : tier_input_gen(seq, Lseq, pos - t.n, step_base, t.n, B, cond, cond_rows, n_cond, spk, c.cond_dim,
                                               c.spk_dim, ctx->lut, t.w_in_t, t.b_in, upper, up_ld, X[i], bf16 ? X16[i] : nullptr,
                                               H, t.kin, t.top, st);
                    g_pdl = 0;
                    SRNN_TRY(rc_in);
                    prev_tier_kernel = true;
                    mark(t.top ? "input top" : "input");
                }
                if (bf16 && fused_cell) {
                    // Fused-cell schedule: the recurrent projections gh_l = W_hh_l h_l + b_hh_l of ALL layers depend only on
                    // the previous step's state, so they go first (two layers per launch); then one launch per layer does
                    // gi_l = W_ih_l x + b_ih_l on tcgen05 and the gate math (k_gru_cell_gen): 1 + NL launches instead of 2 NL.
                    if (!shadow_gh) SRNN_CUDA(cudaStreamWaitEvent(st, ev_join, 0));      // join: the forked gh launch(es) above
                    else if (gh_pending[i]) {                            // gh of this step ran beside the previous sample launch
                        SRNN_CUDA(cudaStreamWaitEvent(st, ev_gh[i], 0));
                        gh_pending[i] = false;
                    }
                    for (int l = gi_fold ? 1 : 0; l < NL; ++l) {
                        g_pdl = pdl;
                        const int rc_c = gru_cell_gen(B, H, in16, t.w_ih16[l], t.b_ih[l], GHL[i][l], hid[i] + (size_t)l * B * H,
                                                      hid16[i] + (size_t)l * B * H, st);
                        g_pdl = 0;
                        SRNN_TRY(rc_c);
                        mark("cell");
                        in16 = hid16[i] + (size_t)l * B * H;
                    }
                }
                for (int l = 0; l < NL && !(bf16 && fused_cell); ++l) {
                    float* h = hid[i] + (size_t)l * B * H;
                    bf* h16 = hid16[i] + (size_t)l * B * H;
                    if (bf16) {      // gi = W_ih x + b_ih and gh = W_hh h + b_hh side by side in one launch
                        GemmOperands ops[2] = {
                            {t.w_ih16[l], in16, t.b_ih[l], nullptr, GI[i], nullptr, 3 * H, H, H, 0, 3 * H, 0},
                            {t.w_hh16[l], h16, t.b_hh[l], nullptr, GH[i], nullptr, 3 * H, H, H, 0, 3 * H, 0}};
                        SRNN_TRY(gemm_umma_multi(ops, 2, B, H, 128, bn_for(3 * H, 2), st));
                    } else if (x3) {
                        SRNN_TRY(gemm_x3(B, 3 * H, H, in, H, t.w_ih3[l], t.b_ih[l], 0, GI[i], 3 * H, S3, 128, bn_for(3 * H, 1), st));
                        SRNN_TRY(gemm_x3(B, 3 * H, H, h, H, t.w_hh3[l], t.b_hh[l], 0, GH[i], 3 * H, S3, 128, bn_for(3 * H, 1), st));
                    } else {
                        SRNN_TRY(gemm_f32(B, 3 * H, H, in, H, t.w_ih[l], H, t.b_ih[l], nullptr, 0, 0, GI[i], 3 * H, st));
                        SRNN_TRY(gemm_f32(B, 3 * H, H, h, H, t.w_hh[l], H, t.b_hh[l], nullptr, 0, 0, GH[i], 3 * H, st));
                    }
                    mark("gi+gh gemm");
                    SRNN_TRY(gru_gates(GI[i], 3 * H, GH[i], 3 * H, h, H, h, H, nullptr, B, H, st, bf16 ? h16 : nullptr));
                    mark("gates");
                    in = h;
                    in16 = h16;
                }
                if (gi_fold && t.top && NT == 2) {
                    // consumed only by tier 0's folded first layer, whose first frame goes through gup0_16: off the serial path,
                    // launched with the shadow work of the next sample launch
                    up_top_todo = true;
                } else if (bf16 && gemm_umma_pair_wide_ok(t.fs * H, B)) {      // big upsampling (tier 2 at C2): one wave of CTA pairs
                    GemmOperands o{t.w_up16, in16, t.b_up, nullptr, OUT[i], nullptr, t.fs * H, H, H, 0, t.fs * H, 0, nullptr};
                    g_pdl = pdl;
                    const int rc_u = gemm_umma_pair_wide(o, B, H, st);
                    g_pdl = 0;
                    SRNN_TRY(rc_u);
                    if (time_tiers && getenv("SRNN_UP_TWICE")) {     // timing experiment: the same launch with its weights in L2
                        mark("upsample (first)");
                        SRNN_TRY(gemm_umma_pair_wide(o, B, H, st));
                    }
                } else if (bf16 && gemm_umma_swap_pair_ok(t.fs * H, B)) {
                    GemmOperands o{t.w_up16, in16, t.b_up, nullptr, OUT[i], nullptr, t.fs * H, H, H, 0, t.fs * H, 0, nullptr};
                    SRNN_TRY(gemm_umma_swap_pair(o, B, H, st));
                } else if (bf16) {
                    g_pdl = pdl;
                    const int rc_u = gemm_umma(t.w_up16, t.fs * H, in16, B, H, H, H, t.b_up, nullptr, 0, OUT[i], nullptr,
                                               t.fs * H, 0, 128, bn_for(t.fs * H, 1), st);
                    g_pdl = 0;
                    SRNN_TRY(rc_u);
                }
                else if (x3)
                    SRNN_TRY(gemm_x3(B, t.fs * H, H, in, H, t.w_up3, t.b_up, 0, OUT[i], t.fs * H, S3, 128, bn_for(t.fs * H, 1), st));
                else
                    SRNN_TRY(gemm_f32(B, t.fs * H, H, in, H, t.w_up, H, t.b_up, nullptr, 0, 0, OUT[i], t.fs * H, st));
                mark(t.top ? "upsample top" : "upsample");
                if (shadow_gh) gh_todo[i] = true;
            }
            const float* up0 = OUT[0] + (size_t)(pos % FS0) * H;                     // model.py:504-513
            if (persist) {
                if (pos % FS0) continue;             // one persistent launch covers the FS0 samples of a tier-0 frame
                const bool sample_dep = pdl_sample && prev_tier_kernel;
                prev_tier_kernel = false;
                if (shadow_gh) {                     // next step's recurrent projections: on the spare SMs, beside this launch
                    const bool xp_now = shadow_in && pos == lookback - FS0;      // last sample launch of the period
                    bool any = xp_now || gip_todo;
                    for (int i = 0; i < NT; ++i) any = any || gh_todo[i];
                    if (any) {
                        SRNN_CUDA(cudaEventRecord(ev_fork, st));
                        SRNN_CUDA(cudaStreamWaitEvent(st2, ev_fork, 0));
                        for (int i = 0; i < NT; ++i) {           // lowest tier first: its next step comes first
                            if (gh_todo[i]) {
                                SRNN_TRY(launch_gh(i, st2, spare_sms));
                                SRNN_CUDA(cudaEventRecord(ev_gh[i], st2));
                                gh_todo[i] = false;
                                gh_pending[i] = true;
                            }
                            if (i == 0 && gip_todo) {            // folded first layer of tier 0: frames 1 .. fs_top-1 of the period
                                const TierPacked& t0 = ctx->tiers[0];
                                gemm_umma_set_cta_cap(spare_sms);
                                int rc_g = SRNN_OK;
                                if (up_top_todo) {               // the top tier's upsampling (bf16, (B * fs, H) row-major) they contract
                                    const TierPacked& tt = ctx->tiers[NT - 1];
                                    rc_g = gemm_umma(tt.w_up16, tt.fs * H, hid16[NT - 1] + (size_t)(NL - 1) * B * H, B, H, H, H, tt.b_up,
                                                     nullptr, 0, nullptr, UP16, tt.fs * H, 0, 128, B <= 128 ? bn_tier : 128, st2);
                                    up_top_todo = false;
                                }
                                for (int f = 1; f < fs_top && rc_g == SRNN_OK; f += 2) {
                                    const int np = fs_top - f >= 2 ? 2 : 1;
                                    GemmOperands ops[2];
                                    for (int q = 0; q < np; ++q)
                                        ops[q] = GemmOperands{t0.w_ih16[0], UP16 + (size_t)(f + q) * H, t0.b_gi0, nullptr,
                                                              GIP + (size_t)(f + q) * 3 * H, nullptr, 3 * H, H, fs_top * H, 0,
                                                              fs_top * 3 * H, 0, nullptr};
                                    rc_g = gemm_umma_multi(ops, np, B, H, 128, B <= 128 ? bn_tier : 128, st2);
                                }
                                gemm_umma_set_cta_cap(0);
                                SRNN_TRY(rc_g);
                                SRNN_CUDA(cudaEventRecord(ev_gip, st2));
                                gip_todo = false;
                                gip_pending = true;
                            }
                        }
                        if (l2_prefetch) {               // bf16 weights of the tier step(s) that follow this sample launch -> L2
                            L2PrefetchArgs pa{};
                            auto want = [&](const void* ptr, size_t bytes) {
                                if (pa.n < 8) { pa.ptr[pa.n] = ptr; pa.bytes[pa.n] = bytes; ++pa.n; }
                            };
                            const int next_pos = (pos + FS0) % lookback;
                            for (int i = NT - 1; i >= 0; --i) {
                                const TierPacked& t = ctx->tiers[i];
                                if (next_pos % t.n) continue;
                                for (int l = 0; l < NL && l < 2; ++l) want(t.w_ih16[l], sizeof(bf) * 3 * (size_t)H * H);
                                want(t.w_up16, sizeof(bf) * (size_t)t.fs * H * H);
                            }
                            if (pa.n) SRNN_TRY(prefetch_l2(pa, l2_prefetch, st2));
                        }
                        if (xp_now) {                    // next period's top-tier input, all but its last FS0 sample columns
                            SRNN_TRY(launch_xp(lookback - ttop.n, st2));
                            SRNN_CUDA(cudaEventRecord(ev_xp, st2));
                            xp_pending = true;
                        }
                    }
                }
                srnn::MlpPersistParams mp;
                mp.B = B; mp.H = H; mp.FS = FS0; mp.nsteps = FS0; mp.pos0 = pos; mp.lookback = lookback;
                mp.Lseq = Lseq; mp.T = T; mp.step_base = step_base; mp.seq = seq; mp.c0 = OUT[0];
                mp.tbl = ctx->tbl16; mp.b_hid = ctx->b_hid; mp.b_out = ctx->b_out; mp.x1 = X1h; mp.part = part;
                mp.dbg = getenv("SRNN_MC_DBG") ? atoi(getenv("SRNN_MC_DBG")) : 0;
                mp.pcarry = carry ? pcarry : nullptr;
                mp.carry_plane = (long long)x1_rows * H;
                mp.ctr = gctr; mp.uniforms = uniforms; mp.u_ld = u_ld; mp.logp_out = logp_out; mp.trace = trace;
                auto launch_sample = [&]() -> int {
                    g_pdl = sample_dep;
                    const int rc_s = cluster_rows ? mlp_cluster_launch(ctx->w_hid16, ctx->w_out16, mp, cluster_rows, st)
                                                  : mlp_persist_launch(ctx->w_hid16, ctx->w_out16, mp, st);
                    g_pdl = 0;
                    prev_sample_kernel = cluster_rows != 0;
                    return rc_s;
                };
                if (time_kernels) {
                    cudaEvent_t e0, e1;
                    SRNN_CUDA(cudaEventCreate(&e0));
                    SRNN_CUDA(cudaEventCreate(&e1));
                    SRNN_CUDA(cudaEventRecord(e0, st));
                    SRNN_TRY(launch_sample());
                    SRNN_CUDA(cudaEventRecord(e1, st));
                    tev.push_back(e0);
                    tev.push_back(e1);
                } else {
                    SRNN_TRY(launch_sample());
                }
                continue;
            }
            if (bf16) {
                SRNN_TRY(mlp_gather_bf16(seq, Lseq, pos - FS0, step_base, ctx->tbl16, up0, (long long)FS0 * H, 0, X1h, B,
                                         1, H, FS0, st));
                SRNN_TRY(gemm_umma(ctx->w_hid16, H, X1h, B, H, H, H, ctx->b_hid, nullptr, 0, nullptr, X2h, H, 1, bm_hid, bn_hid, st));
                SRNN_TRY(gemm_umma(ctx->w_out16, Q, X2h, B, H, H, H, ctx->b_out, nullptr, 0, LG, nullptr, Q, 0, bm_out, bn_out, st));
            } else {
                SRNN_TRY(mlp_gather(seq, Lseq, pos - FS0, step_base, ctx->tbl, up0, (long long)FS0 * H, 0, X1, B, 1, H,
                                    FS0, st));
                if (x3) {
                    SRNN_TRY(gemm_x3(B, H, H, X1, H, ctx->w_hid3, ctx->b_hid, 1, X2, H, S3, bm_hid, bn_hid, st));
                    SRNN_TRY(gemm_x3(B, Q, H, X2, H, ctx->w_out3, ctx->b_out, 0, LG, Q, S3, bm_out, bn_out, st));
                } else {
                    SRNN_TRY(gemm_f32(B, H, H, X1, H, ctx->w_hid, H, ctx->b_hid, nullptr, 0, 1, X2, H, st));
                    SRNN_TRY(gemm_f32(B, Q, H, X2, H, ctx->w_out, H, ctx->b_out, nullptr, 0, 0, LG, Q, st));
                }
            }
            SRNN_TRY(softmax_sample(LG, uniforms, u_ld, seq, Lseq, pos, lookback, step_base, logp_out,
                                    (long long)T * Q, B, st));                       // model.py:514-517
        }
        for (int i = 0; i < NT; ++i) {            // join the side stream: the next period starts with every gh in place
            if (!gh_pending[i]) continue;
            SRNN_CUDA(cudaStreamWaitEvent(st, ev_gh[i], 0));
            gh_pending[i] = false;
        }
        if (xp_pending) {
            SRNN_CUDA(cudaStreamWaitEvent(st, ev_xp, 0));
            xp_pending = false;
        }
        if (gip_pending) {
            SRNN_CUDA(cudaStreamWaitEvent(st, ev_gip, 0));
            gip_pending = false;
        }
        SRNN_TRY(add_int(step_base, lookback, st));
        return SRNN_OK;
    };
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    if (cache_hit) {
        for (int p = 0; p < n_cond; ++p) SRNN_CUDA(cudaGraphLaunch(gc.exec, st));
        g_launches.fetch_add(gc.nodes * (long long)n_cond);
        ++gc.reuses;
    } else if (use_graph) {
        const int rc = body();
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        const long long nodes = g_launches.load() - before;
        if (rc != SRNN_OK && !persist) {
            if (graph) cudaGraphDestroy(graph);
            cudaStreamDestroy(st);
            return rc;
        }
        if (rc != SRNN_OK) ce = cudaErrorStreamCaptureUnsupported;   // the cooperative launch refused to be captured
        if (ce == cudaSuccess) ce = cudaGraphInstantiate(&exec, graph, 0);
        if (ce == cudaSuccess) {
            for (int p = 0; p < n_cond; ++p) SRNN_CUDA(cudaGraphLaunch(exec, st));
            g_launches.fetch_add(nodes * (long long)(n_cond - 1));
            if (cache_ok) {           // keep it for the next call of the same shape (the previous one is released once idle)
                if (gc.exec) cudaGraphExecDestroy(gc.exec);
                if (gc.graph) cudaGraphDestroy(gc.graph);
                gc.exec = exec;
                gc.graph = graph;
                gc.nodes = nodes;
                gc.nkey = ngkey;
                memcpy(gc.key, gkey, sizeof(gkey));
                exec = nullptr;
                graph = nullptr;
            }
        } else {
            // e.g. a driver that cannot put the cooperative persistent kernel into a graph: issue the launches directly
            cudaGetLastError();
            g_launches.fetch_sub(nodes);
            if (graph) cudaGraphDestroy(graph);
            graph = nullptr;
            exec = nullptr;
            use_graph = false;
            if (getenv("SRNN_TRACE")) fprintf(stderr, "[srnn] graph path unavailable (%s); direct launches\n", cudaGetErrorString(ce));
        }
    }
    if (!use_graph) {
        for (int p = 0; p < n_cond; ++p) SRNN_TRY(body());
    }
    SRNN_TRY(dequant_audio(seq, Lseq, lookback, ctx->lut, samples_out, audio_out, B, T, st));   // model.py:520
    if (time_tiers && !marks.empty()) {
        SRNN_CUDA(cudaStreamSynchronize(st));
        std::vector<std::pair<const char*, std::pair<double, int>>> agg;
        for (size_t k = 1; k < marks.size(); ++k) {
            if (marks[k].first[0] == '(') continue;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks[k - 1].second, marks[k].second);
            bool found = false;
            for (auto& a : agg)
                if (!strcmp(a.first, marks[k].first)) { a.second.first += ms; a.second.second++; found = true; }
            if (!found) agg.push_back({marks[k].first, {ms, 1}});
        }
        fprintf(stderr, "[srnn tiers] avg us per kernel:");
        for (auto& a : agg) fprintf(stderr, " %s=%.1f (n=%d)", a.first, 1e3 * a.second.first / a.second.second, a.second.second);
        fprintf(stderr, "\n");
        for (auto& m2 : marks) cudaEventDestroy(m2.second);
    }
    if (time_kernels) {   // CUDA-event duration of every persistent launch, on the stream it ran on
        SRNN_CUDA(cudaStreamSynchronize(st));
        double tot = 0;
        for (size_t k = 0; k + 1 < tev.size(); k += 2) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, tev[k], tev[k + 1]);
            tot += ms;
            cudaEventDestroy(tev[k]);
            cudaEventDestroy(tev[k + 1]);
        }
        ctx->timed_ms = tot;
        ctx->timed_launches = (long long)(tev.size() / 2);
    }
    if (trace && cluster_rows) {   // debugging aid: phase durations (SM cycles) of the 8 CTAs of cluster 0, last launch
        SRNN_CUDA(cudaStreamSynchronize(st));
        std::vector<long long> ht(trace_n);
        SRNN_CUDA(cudaMemcpy(ht.data(), trace, sizeof(long long) * trace_n, cudaMemcpyDeviceToHost));
        static const char* nm[9] = {"tbl+waitP", "x1 rows", "signal", "flags+TMA+MMA1", "epi1", "MMA2", "epi2+push", "land wait",
                                    "reduce+sample"};
        for (int cta = 0; cta < 8; ++cta) {
            const long long* tr = ht.data() + (size_t)cta * FS0 * 64;
            double acc[9] = {0}, tot = 0, woke = 0, lastmma = 0;
            for (int k = 2; k < FS0; ++k) {
                for (int j = 0; j < 9; ++j) acc[j] += (double)(tr[k * 64 + j + 1] - tr[k * 64 + j]);
                woke += (double)(tr[k * 64 + 10] - tr[k * 64 + 3]);
                lastmma += (double)(tr[k * 64 + 14] - tr[k * 64 + 3]);
            }
            for (int k = 3; k < FS0; ++k) tot += (double)(tr[k * 64] - tr[(k - 1) * 64]);
            fprintf(stderr, "[srnn trace] k_mlp_cluster CTA%d:", cta);
            for (int j = 0; j < 9; ++j) fprintf(stderr, " %s=%.0f", nm[j], acc[j] / (FS0 - 2));
            double red = 0, smx = 0;
            for (int k = 2; k < FS0; ++k) {
                red += (double)(tr[k * 64 + 11] - tr[k * 64 + 8]);
                smx += (double)(tr[k * 64 + 12] - tr[k * 64 + 11]);
            }
            {
                double ga[4] = {0}, d1 = 0, m2 = 0;
                for (int k = 2; k < FS0; ++k) {
                    for (int g2 = 0; g2 < 4; ++g2) ga[g2] += (double)(tr[k * 64 + 16 + g2] - tr[k * 64 + 10]);
                    d1 += (double)(tr[k * 64 + 4] - tr[k * 64 + 10]);
                    m2 += (double)(tr[k * 64 + 15] - tr[k * 64 + 5]);
                }
                fprintf(stderr, " [since TMA issue: group landed %.0f %.0f %.0f %.0f, D1 done %.0f; MMA2 issued %.0f after x2]",
                        ga[0] / (FS0 - 2), ga[1] / (FS0 - 2), ga[2] / (FS0 - 2), ga[3] / (FS0 - 2), d1 / (FS0 - 2), m2 / (FS0 - 2));
            }
            fprintf(stderr, " (reduce %.0f softmax+sample %.0f) | flags@%.0f lastMMA@%.0f (since signal) | step=%.0f | entry->step0=%lld step0=%lld\n",
                    red / (FS0 - 2), smx / (FS0 - 2), woke / (FS0 - 2), lastmma / (FS0 - 2), tot / (FS0 - 3), tr[0] - tr[63], tr[64] - tr[0]);
            if (cta == 0)
                for (int k = 0; k < 2; ++k) {
                    fprintf(stderr, "[srnn trace]    CTA0 step %d:", k);
                    for (int j = 0; j < 9; ++j) fprintf(stderr, " %s=%lld", nm[j], tr[k * 64 + j + 1] - tr[k * 64 + j]);
                    fprintf(stderr, " | x1 flags since signal %lld, D1 since TMA issue %lld\n", tr[k * 64 + 10] - tr[k * 64 + 3],
                            tr[k * 64 + 4] - tr[k * 64 + 10]);
                }
            fprintf(stderr, "[srnn trace]    prologue CTA%d (cycles since entry): init+alloc+sync %lld, weights in TMEM %lld, P(0) gathered %lld, cluster sync done %lld, step 0 starts %lld\n",
                    cta, tr[20] - tr[63], tr[21] - tr[63], tr[22] - tr[63], tr[23] - tr[63], tr[0] - tr[63]);
        }
        cudaFree(trace);
        trace = nullptr;
    }
    if (trace) {   // debugging aid: average phase durations (SM cycles) of CTA 0 over the last persistent launch
        SRNN_CUDA(cudaStreamSynchronize(st));
        {
            long long* ht = (long long*)malloc(sizeof(long long) * trace_n);
            cudaMemcpy(ht, trace, sizeof(long long) * trace_n, cudaMemcpyDeviceToHost);
            cudaFree(trace);
            trace = ht;
        }
        static const char* names_v1[9] = {"wait P", "x1 slice", "barrier A", "TMA+MMA1", "epilogue1", "MMA2", "epilogue2",
                                          "barrier B", "reduce+sample"};
        static const char* names_v2[9] = {"tbl loads+wait P", "x1 rows", "signal", "flags+TMA+MMA1", "epilogue1", "MMA2",
                                          "epilogue2+push", "land wait", "reduce+sample"};
        const char* const* names = cluster_rows ? names_v2 : names_v1;
        double acc[9] = {0}, tot = 0;
        double w[5] = {0};
        for (int k = 2; k < FS0; ++k) {
            for (int j = 0; j < 9; ++j) acc[j] += (double)(trace[k * 64 + j + 1] - trace[k * 64 + j]);
            for (int j = 0; j < 5; ++j) w[j] += (double)(trace[k * 64 + 10 + j] - trace[k * 64 + 3]);
        }
        for (int k = 3; k < FS0; ++k) tot += (double)(trace[k * 64] - trace[(k - 1) * 64]);
        fprintf(stderr, "[srnn trace] since barrier A: W0 woke=%.0f W0 issued all=%.0f W1 full[0]=%.0f W1 full[7]=%.0f W1 last MMA issued=%.0f\n",
                w[0] / (FS0 - 2), w[1] / (FS0 - 2), w[2] / (FS0 - 2), w[3] / (FS0 - 2), w[4] / (FS0 - 2));
        {
            const int k = FS0 - 1;
            fprintf(stderr, "[srnn trace] W1 per k-block (wait start, wait end) since barrier A, last step:");
            for (int kb = 0; kb < 16 && kb < H / 64; ++kb)
                fprintf(stderr, " [%lld,%lld]", trace[k * 64 + 16 + 2 * kb] - trace[k * 64 + 3], trace[k * 64 + 17 + 2 * kb] - trace[k * 64 + 3]);
            fprintf(stderr, "\n");
        }
        fprintf(stderr, "[srnn trace] last launch: entry -> step 0 = %lld, step 0 = %lld, step 1 = %lld, step 2 = %lld cycles\n",
                trace[0] - trace[63], trace[64] - trace[0], trace[128] - trace[64], trace[192] - trace[128]);
        for (int k = 0; k < 2 && k < FS0; ++k) {
            fprintf(stderr, "[srnn trace] step %d phases:", k);
            for (int j = 0; j < 9; ++j) fprintf(stderr, " %s=%lld", names[j], trace[k * 64 + j + 1] - trace[k * 64 + j]);
            fprintf(stderr, "\n");
        }
        fprintf(stderr, "[srnn trace] %s CTA0 cycles/step:", cluster_rows ? "k_mlp_cluster" : "k_mlp_persist");
        for (int j = 0; j < 9; ++j) fprintf(stderr, " %s=%.0f", names[j], acc[j] / (FS0 - 2));
        fprintf(stderr, " | step=%.0f\n", tot / (FS0 - 3));
        free(trace);
    }
    SRNN_CUDA(cudaEventRecord(ev_out, st));
    SRNN_CUDA(cudaStreamWaitEvent(user, ev_out, 0));
    // the graph/stream objects can be released once the work is enqueued; CUDA defers destruction until completion
    if (exec) SRNN_CUDA(cudaGraphExecDestroy(exec));
    if (graph) SRNN_CUDA(cudaGraphDestroy(graph));
    SRNN_CUDA(cudaEventDestroy(ev_in));
    SRNN_CUDA(cudaEventDestroy(ev_out));
    SRNN_CUDA(cudaStreamDestroy(st));
    for (int i = 0; i < SRNN_MAX_TIERS; ++i)
        if (ev_gh[i]) SRNN_CUDA(cudaEventDestroy(ev_gh[i]));
    if (ev_xp) SRNN_CUDA(cudaEventDestroy(ev_xp));
    if (ev_gip) SRNN_CUDA(cudaEventDestroy(ev_gip));
    if (st2) {
        SRNN_CUDA(cudaStreamDestroy(st2));
        SRNN_CUDA(cudaEventDestroy(ev_fork));
        SRNN_CUDA(cudaEventDestroy(ev_join));
    }
    return SRNN_OK;
}

int srnn_generate(srnn_ctx* ctx, int32_t B, int32_t n_cond, const float* cond, int32_t cond_rows,
                  const int64_t* spk, const float* uniforms, uint8_t* samples_out, float* audio_out,
                  float* logp_out, int32_t mode, void* stream) {
    SRNN_TRY(check_ready(ctx));
    if (!cond || !spk || !uniforms) return fail(SRNN_ERR_ARG, "null argument");
    if (!samples_out && !audio_out) return fail(SRNN_ERR_ARG, "need samples_out or audio_out");
    if (B < 1 || n_cond < 1) return fail(SRNN_ERR_ARG, "B and n_cond must be positive");
    if (cond_rows != 1 && cond_rows != B) return fail(SRNN_ERR_ARG, "cond_rows must be 1 or B");
    if (mode != SRNN_MODE_FP32 && !ctx->has_bf16)
        return fail(SRNN_ERR_UNSUPPORTED, "bf16 tensor-core mode needs dim %% 64 == 0 (dim=%d)", ctx->H);
    if (mode == SRNN_MODE_BF16X3) {     // fp32 control flow, dense contractions as split-bf16 tcgen05 GEMMs
        SRNN_TRY(ensure_x3(ctx, (cudaStream_t)stream));
        return generate_graph(ctx, false, false, B, n_cond, cond, cond_rows, spk, uniforms, B, samples_out, audio_out, logp_out,
                              (cudaStream_t)stream, 0, true);
    }
    if (mode == SRNN_MODE_FP32 || mode == SRNN_MODE_BF16 || mode == SRNN_MODE_BF16_GRAPH) {
        const bool bf16 = mode != SRNN_MODE_FP32;
        int n_sms = 0;
        SRNN_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, ctx->device));
        // sample-level kernel: the 8-CTA cluster form (H = 1024) when the batch fits the co-resident clusters, else the 16-CTA
        // row-group form over L2 (any H % 64 == 0); SRNN_MLP_V1=1 forces the latter
        const int maxc = (mode == SRNN_MODE_BF16 && ctx->H == 1024 && !getenv("SRNN_MLP_V1")) ? mlp_cluster_max_clusters() : 0;
        auto cluster_rows_for = [&](int b) {
            if (!mlp_cluster_supported(ctx->H, ctx->FS0, b, maxc)) return 0;
            if (getenv("SRNN_CLUSTER_ROWS")) return atoi(getenv("SRNN_CLUSTER_ROWS")) == 16 && (b + 15) / 16 <= maxc ? 16 : 24;
            return mlp_cluster_rows(b, maxc);
        };
        const int crows = cluster_rows_for(B);
        const bool persist = mode == SRNN_MODE_BF16 && (crows || mlp_persist_supported(ctx->H, ctx->FS0, B, n_sms));
        if (mode == SRNN_MODE_BF16) srnn::g_sample_kernel = crows ? "k_mlp_cluster" : "k_mlp_persist";
        // Batches beyond what the persistent sample-level kernel can keep co-resident (RG * NS CTAs <= SMs: 288 utterances at
        // dim 1024): balanced utterance chunks run back to back through the persistent path (utterances are independent), which
        // beats the one-GEMM-launch-per-contraction schedule up to ~800 utterances (measured chunked / unchunked, x real-time:
        // 512 utterances 1430 / 1009, 768: ~1430 / 1399, 1024: 1427 / 1684; profiles/README.md, C4 sweep) -- at most 3 chunks.
        if (mode == SRNN_MODE_BF16 && !persist && mlp_persist_supported(ctx->H, ctx->FS0, 32, n_sms)) {
            const int NS = ctx->H / 64, max_chunk = (n_sms / NS) * 32;
            const int chunks = (B + max_chunk - 1) / max_chunk;
            static const int max_chunks = getenv("SRNN_MAX_CHUNKS") ? atoi(getenv("SRNN_MAX_CHUNKS")) : 3;
            if (chunks <= max_chunks) {
                const int per = ((B + chunks - 1) / chunks + 31) / 32 * 32;
                const size_t T = (size_t)n_cond * ctx->lookback;
                for (int b0 = 0; b0 < B; b0 += per) {
                    const int Bc = B - b0 < per ? B - b0 : per;
                    const bool own = cond_rows == B;
                    SRNN_TRY(generate_graph(ctx, true, true, Bc, n_cond, cond + (own ? (size_t)b0 * n_cond * ctx->cfg.cond_dim : 0),
                                            own ? Bc : 1, spk + (own ? b0 : 0), uniforms + b0, B,
                                            samples_out ? samples_out + (size_t)b0 * T : nullptr,
                                            audio_out ? audio_out + (size_t)b0 * T : nullptr,
                                            logp_out ? logp_out + (size_t)b0 * T * SRNN_Q : nullptr, (cudaStream_t)stream,
                                            cluster_rows_for(Bc)));
                }
                return SRNN_OK;
            }
        }
        return generate_graph(ctx, bf16, persist, B, n_cond, cond, cond_rows, spk, uniforms, B, samples_out, audio_out,
                              logp_out, (cudaStream_t)stream, crows);
    }
    return fail(SRNN_ERR_UNSUPPORTED, "generate: mode %d not available", mode);
}

// ------------------------------------------------------------------------------------------------
// test hooks
// ------------------------------------------------------------------------------------------------
int srnn_sample_rows(const float* p, const float* u, int32_t rows, int32_t* idx, void* stream) {
    if (!p || !u || !idx) return fail(SRNN_ERR_ARG, "null argument");
    return sample_rows(p, u, rows, idx, (cudaStream_t)stream);
}

// Training-data quantiser (dataset.py:249-253): x (rows, cols; row stride ld) fp32 audio in [-1, 1] -> q (rows, cols) int64.
int srnn_quantize(const float* x, int32_t rows, int32_t cols, int64_t ld, int32_t q_levels, int32_t ulaw, int64_t* q, void* stream) {
    if (!x || !q) return fail(SRNN_ERR_ARG, "null argument");
    if (q_levels < 2 || q_levels > 65536) return fail(SRNN_ERR_ARG, "bad q_levels");
    return quantize_samples(x, rows, cols, ld, q_levels, ulaw, q, (cudaStream_t)stream);
}

int srnn_dequant_lut(const srnn_ctx* ctx, float* out, void* stream) {
    SRNN_TRY(check_ready(ctx));
    if (!out) return fail(SRNN_ERR_ARG, "null argument");
    return copy_f32(ctx->lut, out, ctx->Q, (cudaStream_t)stream);
}

// SampleLevelMLP.forward (model.py:308-325) on the context's packed weights: prev_samples (B, T+FS0-1) int64,
// upper (B, T, H) fp32 -> logp (B, T, 256).  Same kernels as the tail of srnn_predict_fwd.
int srnn_mlp_fwd(srnn_ctx* ctx, int32_t B, int32_t T, const int64_t* prev_samples, const float* upper, float* logp_out,
                 int32_t mode, void* stream) {
    SRNN_TRY(check_ready(ctx));
    if (!prev_samples || !upper || !logp_out || B < 1 || T < 1) return fail(SRNN_ERR_ARG, "bad argument");
    if (mode != SRNN_MODE_FP32 && mode != SRNN_MODE_BF16 && mode != SRNN_MODE_BF16X3)
        return fail(SRNN_ERR_UNSUPPORTED, "mlp_fwd: mode %d not available", mode);
    if (mode != SRNN_MODE_FP32 && !ctx->has_bf16) return fail(SRNN_ERR_UNSUPPORTED, "tensor-core modes need dim %% 64 == 0");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == SRNN_MODE_BF16X3) SRNN_TRY(ensure_x3(ctx, st));
    const int H = ctx->H, Q = ctx->Q, FS0 = ctx->FS0, R = B * T, W = T + FS0 - 1;
    uint8_t* seq = nullptr;
    SRNN_CUDA(cudaMallocAsync((void**)&seq, (size_t)B * W, st));
    SRNN_TRY(i64_to_u8(prev_samples, seq, (size_t)B * W, st));
    int rc = SRNN_OK;
    if (mode == SRNN_MODE_BF16) {
        __nv_bfloat16 *x1 = nullptr, *x2 = nullptr;
        SRNN_CUDA(cudaMallocAsync((void**)&x1, sizeof(__nv_bfloat16) * (size_t)R * H, st));
        SRNN_CUDA(cudaMallocAsync((void**)&x2, sizeof(__nv_bfloat16) * (size_t)R * H, st));
        rc = mlp_gather_bf16(seq, W, 0, nullptr, ctx->tbl16, upper, (long long)T * H, H, x1, B, T, H, FS0, st);
        if (rc == SRNN_OK) rc = tf_gemm(ctx->w_hid16, H, x1, R, H, ctx->b_hid, nullptr, x2, H, 1, st);
        if (rc == SRNN_OK) rc = tf_gemm(ctx->w_out16, Q, x2, R, H, ctx->b_out, logp_out, nullptr, Q, 0, st);
        cudaFreeAsync(x1, st);
        cudaFreeAsync(x2, st);
    } else {
        float *x1 = nullptr, *x2 = nullptr;
        SRNN_CUDA(cudaMallocAsync((void**)&x1, sizeof(float) * (size_t)R * H, st));
        SRNN_CUDA(cudaMallocAsync((void**)&x2, sizeof(float) * (size_t)R * H, st));
        rc = mlp_gather(seq, W, 0, nullptr, ctx->tbl, upper, (long long)T * H, H, x1, B, T, H, FS0, st);
        if (mode == SRNN_MODE_BF16X3) {
            __nv_bfloat16* s3 = nullptr;
            SRNN_CUDA(cudaMallocAsync((void**)&s3, sizeof(__nv_bfloat16) * (size_t)R * 3 * H, st));
            if (rc == SRNN_OK) rc = gemm_x3(R, H, H, x1, H, ctx->w_hid3, ctx->b_hid, 1, x2, H, s3, 0, 0, st);
            if (rc == SRNN_OK) rc = gemm_x3(R, Q, H, x2, H, ctx->w_out3, ctx->b_out, 0, logp_out, Q, s3, 0, 0, st);
            cudaFreeAsync(s3, st);
        } else {
            if (rc == SRNN_OK) rc = gemm_f32(R, H, H, x1, H, ctx->w_hid, H, ctx->b_hid, nullptr, 0, 1, x2, H, st);
            if (rc == SRNN_OK) rc = gemm_f32(R, Q, H, x2, H, ctx->w_out, H, ctx->b_out, nullptr, 0, 0, logp_out, Q, st);
        }
        cudaFreeAsync(x1, st);
        cudaFreeAsync(x2, st);
    }
    if (rc == SRNN_OK) rc = logsoftmax_rows(logp_out, R, st);
    cudaFreeAsync(seq, st);
    return rc;
}

// FrameLevelRNN.forward (model.py:180-263) for one tier: the per-module form of the tier loop of srnn_predict_fwd (same kernels:
// input expansion GEMM with the upper conditioning as addend, frame-by-frame GRU, learned upsampling as a GEMM).
int srnn_tier_fwd(srnn_ctx* ctx, int32_t tier, int32_t B, int32_t F, const float* prev_samples, const float* upper,
                  const float* cond, const int64_t* spk, float* hidden_io, int32_t reset, float* out, int32_t mode, void* stream) {
    SRNN_TRY(check_ready(ctx));
    const srnn_config& c = ctx->cfg;
    if (tier < 0 || tier >= c.n_tiers || B < 1 || F < 1 || !prev_samples || !hidden_io || !out) return fail(SRNN_ERR_ARG, "tier_fwd: bad argument");
    if (mode != SRNN_MODE_FP32 && mode != SRNN_MODE_BF16X3) return fail(SRNN_ERR_UNSUPPORTED, "tier_fwd: mode %d not available", mode);
    const TierPacked& t = ctx->tiers[tier];
    if (t.top ? (!cond || !spk || upper) : (!upper || cond || spk))
        return fail(SRNN_ERR_ARG, "tier_fwd: the top tier takes cond + spk, the others the upper tier's conditioning (model.py:199-217)");
    cudaStream_t st = (cudaStream_t)stream;
    const bool x3 = mode == SRNN_MODE_BF16X3;
    if (x3) SRNN_TRY(ensure_x3(ctx, st));
    const int H = ctx->H, NL = c.n_rnn, M = B * F;
    ctx->fwd.valid = false;                      // the scratch below overlays the saved activations of a teacher-forced pass
    float *A = nullptr, *X = nullptr, *GI = nullptr, *GH = nullptr, *Y[2] = {nullptr, nullptr}, *h0 = nullptr;
    __nv_bfloat16* S3 = nullptr;
    for (int pass = 0; pass < 2; ++pass) {
        Bump b(pass ? ctx->ws : nullptr);
        A = b.take<float>((size_t)M * t.kin);
        X = b.take<float>((size_t)M * H);
        GI = b.take<float>((size_t)M * 3 * H);
        GH = b.take<float>((size_t)M * 3 * H);
        Y[0] = b.take<float>((size_t)M * H);
        Y[1] = b.take<float>((size_t)M * H);
        h0 = b.take<float>((size_t)B * H);
        S3 = b.take<__nv_bfloat16>(x3 ? (size_t)M * 3 * H : 1);
        if (!pass) SRNN_TRY(ensure_ws(ctx, b.off));
    }
    SRNN_TRY(tier_assemble_f32(prev_samples, t.n, cond, c.cond_dim, spk, c.spk_dim, M, F, A, t.kin, st));
    SRNN_TRY(gemm_f32(M, H, t.kin, A, t.kin, t.w_in, t.kin, t.b_in, upper, H, 0, X, H, st));
    const float* in = X;
    for (int l = 0; l < NL; ++l) {
        float* y = Y[l & 1];
        float* hid = hidden_io + (size_t)l * B * H;
        if (reset) SRNN_TRY(bcast_rows(t.h0 + (size_t)l * H, h0, B, H, st));
        else SRNN_TRY(copy_f32(hid, h0, (size_t)B * H, st));
        if (x3) SRNN_TRY(gemm_x3(M, 3 * H, H, in, H, t.w_ih3[l], t.b_ih[l], 0, GI, 3 * H, S3, 0, 0, st));
        else SRNN_TRY(gemm_f32(M, 3 * H, H, in, H, t.w_ih[l], H, t.b_ih[l], nullptr, 0, 0, GI, 3 * H, st));
        for (int f = 0; f < F; ++f) {
            const float* hp = f ? y + (size_t)(f - 1) * H : h0;
            const int hp_ld = f ? F * H : H;
            float* gh = GH + (size_t)f * 3 * H;
            if (x3) SRNN_TRY(gemm_x3(B, 3 * H, H, hp, hp_ld, t.w_hh3[l], t.b_hh[l], 0, gh, F * 3 * H, S3, 0, 0, st));
            else SRNN_TRY(gemm_f32(B, 3 * H, H, hp, hp_ld, t.w_hh[l], H, t.b_hh[l], nullptr, 0, 0, gh, F * 3 * H, st));
            SRNN_TRY(gru_gates(GI + (size_t)f * 3 * H, F * 3 * H, gh, F * 3 * H, hp, hp_ld, y + (size_t)f * H, F * H,
                               f == F - 1 ? hid : nullptr, B, H, st, nullptr, F * H));
        }
        in = y;
    }
    if (x3) return gemm_x3(M, t.fs * H, H, in, H, t.w_up3, t.b_up, 0, out, t.fs * H, S3, 0, 0, st);
    return gemm_f32(M, t.fs * H, H, in, H, t.w_up, H, t.b_up, nullptr, 0, 0, out, t.fs * H, st);
}

// One GRU layer over F frames (torch nn.GRU as used at model.py:133-159,244).  FP32: the frame-by-frame fp32 schedule;
// BF16: the persistent tcgen05 kernels of gru_persist.cu (B <= 128, H % 64 == 0).
int srnn_gru_seq_fwd(int32_t B, int32_t F, int32_t H, const float* gi, const float* w_hh, const float* b_hh, const float* h0,
                     float* y, float* gh, float* h_last, int32_t mode, void* stream) {
    if (!gi || !w_hh || !b_hh || !h0 || !y || !gh) return fail(SRNN_ERR_ARG, "null argument");
    if (B < 1 || F < 1 || H < 1) return fail(SRNN_ERR_ARG, "bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == SRNN_MODE_FP32) {
        for (int f = 0; f < F; ++f) {
            const float* hp = f ? y + (size_t)(f - 1) * H : h0;
            const int hp_ld = f ? F * H : H;
            SRNN_TRY(gemm_f32(B, 3 * H, H, hp, hp_ld, w_hh, H, b_hh, nullptr, 0, 0, gh + (size_t)f * 3 * H, F * 3 * H, st));
            SRNN_TRY(gru_gates(gi + (size_t)f * 3 * H, F * 3 * H, gh + (size_t)f * 3 * H, F * 3 * H, hp, hp_ld, y + (size_t)f * H,
                               F * H, f == F - 1 ? h_last : nullptr, B, H, st));
        }
        return SRNN_OK;
    }
    if (mode != SRNN_MODE_BF16) return fail(SRNN_ERR_UNSUPPORTED, "gru_seq_fwd: mode %d not available", mode);
    int dev = 0, n_sms = 0;
    SRNN_CUDA(cudaGetDevice(&dev));
    SRNN_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
    if (!gru_persist_supported(B, H, n_sms)) return fail(SRNN_ERR_UNSUPPORTED, "gru_seq_fwd bf16: needs B <= 128 and H %% 64 == 0");
    __nv_bfloat16 *w16 = nullptr, *h16 = nullptr, *y16 = nullptr;
    unsigned* ctr = nullptr;
    SRNN_CUDA(cudaMallocAsync((void**)&w16, sizeof(__nv_bfloat16) * 3 * (size_t)H * H, st));
    SRNN_CUDA(cudaMallocAsync((void**)&h16, sizeof(__nv_bfloat16) * (size_t)B * H, st));
    SRNN_CUDA(cudaMallocAsync((void**)&y16, sizeof(__nv_bfloat16) * (size_t)B * F * H, st));
    SRNN_CUDA(cudaMallocAsync((void**)&ctr, 256, st));
    SRNN_TRY(f32_to_bf16_pad(w_hh, 3 * H, H, H, w16, 3 * H, H, st));
    SRNN_TRY(f32_to_bf16_pad(h0, B, H, H, h16, B, H, st));
    int rc = gru_persist_fwd(B, F, H, gi, w16, b_hh, h0, h16, gh, y, y16, h_last, ctr, st);
    cudaFreeAsync(w16, st);
    cudaFreeAsync(h16, st);
    cudaFreeAsync(y16, st);
    cudaFreeAsync(ctr, st);
    return rc;
}

// BPTT of srnn_gru_seq_fwd: dy (B*F, H) -> dgi, dgh (B*F, 3H), dh0 (B, H).
int srnn_gru_seq_bwd(int32_t B, int32_t F, int32_t H, const float* gi, const float* gh, const float* y, const float* h0,
                     const float* w_hh, const float* dy, float* dgi, float* dgh, float* dh0, int32_t mode, void* stream) {
    if (!gi || !gh || !y || !h0 || !w_hh || !dy || !dgi || !dgh) return fail(SRNN_ERR_ARG, "null argument");
    if (B < 1 || F < 1 || H < 1) return fail(SRNN_ERR_ARG, "bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == SRNN_MODE_FP32) {
        float* scratch = nullptr;
        SRNN_CUDA(cudaMallocAsync((void**)&scratch, sizeof(float) * 2 * (size_t)B * H, st));
        int rc = gru_seq_bwd_f32(B, F, H, gi, gh, y, h0, dy, w_hh, dgi, dgh, dh0, scratch, st);
        cudaFreeAsync(scratch, st);
        return rc;
    }
    if (mode != SRNN_MODE_BF16) return fail(SRNN_ERR_UNSUPPORTED, "gru_seq_bwd: mode %d not available", mode);
    int dev = 0, n_sms = 0;
    SRNN_CUDA(cudaGetDevice(&dev));
    SRNN_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
    if (!gru_persist_supported(B, H, n_sms)) return fail(SRNN_ERR_UNSUPPORTED, "gru_seq_bwd bf16: needs B <= 128 and H %% 64 == 0");
    __nv_bfloat16 *wt16 = nullptr, *dgi16 = nullptr, *dgh16 = nullptr;
    unsigned* ctr = nullptr;
    SRNN_CUDA(cudaMallocAsync((void**)&wt16, sizeof(__nv_bfloat16) * 3 * (size_t)H * H, st));
    SRNN_CUDA(cudaMallocAsync((void**)&dgi16, sizeof(__nv_bfloat16) * 3 * (size_t)B * F * H, st));
    SRNN_CUDA(cudaMallocAsync((void**)&dgh16, sizeof(__nv_bfloat16) * 3 * (size_t)B * F * H, st));
    SRNN_CUDA(cudaMallocAsync((void**)&ctr, 256, st));
    SRNN_TRY(transpose_to_bf16(w_hh, 3 * H, H, H, wt16, 3 * H, st));
    int rc = gru_persist_bwd(B, F, H, gi, gh, y, h0, dy, wt16, dgi, dgh, dgi16, dgh16, dh0, ctr, st);
    cudaFreeAsync(wt16, st);
    cudaFreeAsync(dgi16, st);
    cudaFreeAsync(dgh16, st);
    cudaFreeAsync(ctr, st);
    return rc;
}

int srnn_gemm(int32_t M, int32_t N, int32_t K, const float* A, const float* B, const float* bias,
              const float* addend, int32_t relu, float* C, int32_t mode, void* stream) {
    if (!A || !B || !C) return fail(SRNN_ERR_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == SRNN_MODE_FP32) return gemm_f32(M, N, K, A, K, B, K, bias, addend, N, relu, C, N, st);
    if (mode == SRNN_MODE_BF16X3) {   // split-bf16 product on tcgen05 (the contraction of the tensor-core parity mode)
        if (K % 64 || addend) return fail(SRNN_ERR_ARG, "gemm hook (split bf16): K %% 64 == 0, no addend");
        const int Np = (N + 127) / 128 * 128;
        __nv_bfloat16 *a3 = nullptr, *w3 = nullptr;
        SRNN_CUDA(cudaMallocAsync((void**)&a3, sizeof(__nv_bfloat16) * (size_t)M * 3 * K, st));
        SRNN_CUDA(cudaMallocAsync((void**)&w3, sizeof(__nv_bfloat16) * (size_t)Np * 3 * K, st));
        SRNN_CUDA(cudaMemsetAsync(w3, 0, sizeof(__nv_bfloat16) * (size_t)Np * 3 * K, st));
        int rc = split3_bf16(B, N, K, K, w3, 1, st);
        if (rc == SRNN_OK) rc = gemm_x3(M, N, K, A, K, w3, bias, relu, C, N, a3, 0, 0, st);
        cudaFreeAsync(a3, st);
        cudaFreeAsync(w3, st);
        return rc;
    }
    if ((mode & 0xff) == SRNN_MODE_BF16) {
        // tile selector for tests: bits 8..15 = UMMA M (0 -> 128), bits 16..27 = batch-row tile (0 -> 64),
        // bit 28 = ROWS orientation (activation rows on the TMEM lanes; bits 16..27 then give the feature tile),
        // bit 29 = split the K loop three ways (plain fp32 output only: bias/addend/relu must be absent)
        const int bm = ((mode >> 8) & 0xff) ? ((mode >> 8) & 0xff) : 128;
        const int bn = ((mode >> 16) & 0xfff) ? ((mode >> 16) & 0xfff) : 64;
        const bool rows = (mode >> 28) & 1, split = (mode >> 29) & 1, mnmaj = (mode >> 30) & 1;
        if (mnmaj) {   // bit 30: C = A . B^T through the MN-major kernel: operands handed over as A^T (K, M) and B^T (K, N)
            if (bias || addend || relu || N % 8) return fail(SRNN_ERR_ARG, "gemm hook (MN-major): plain output, N %% 8 == 0");
            const int Mp = (M + 7) / 8 * 8, Np8 = (N + 7) / 8 * 8;
            __nv_bfloat16 *at = nullptr, *bt = nullptr;
            float* scratch = nullptr;
            SRNN_CUDA(cudaMallocAsync((void**)&at, sizeof(__nv_bfloat16) * (size_t)K * Mp, st));
            SRNN_CUDA(cudaMallocAsync((void**)&bt, sizeof(__nv_bfloat16) * (size_t)K * Np8, st));
            if (split) SRNN_CUDA(cudaMallocAsync((void**)&scratch, sizeof(float) * 3 * (size_t)M * N, st));
            SRNN_TRY(transpose_to_bf16(A, M, K, K, at, Mp, st));
            SRNN_TRY(transpose_to_bf16(B, N, K, K, bt, Np8, st));
            int rc = gemm_umma_tn(at, Mp, bt, Np8, M, N, K, C, N, split ? 3 : 1, scratch, st);
            cudaFreeAsync(at, st);
            cudaFreeAsync(bt, st);
            if (scratch) cudaFreeAsync(scratch, st);
            return rc;
        }
        const int Kp = (K + 63) / 64 * 64, Np = (N + bm - 1) / bm * bm;
        __nv_bfloat16 *a16 = nullptr, *w16 = nullptr;
        SRNN_CUDA(cudaMallocAsync((void**)&a16, sizeof(__nv_bfloat16) * (size_t)M * Kp, st));
        SRNN_CUDA(cudaMallocAsync((void**)&w16, sizeof(__nv_bfloat16) * (size_t)Np * Kp, st));
        SRNN_TRY(f32_to_bf16_pad(A, M, K, K, a16, M, Kp, st));
        SRNN_TRY(f32_to_bf16_pad(B, N, K, K, w16, Np, Kp, st));
        float* scratch = nullptr;
        if (split) SRNN_CUDA(cudaMallocAsync((void**)&scratch, sizeof(float) * 3 * (size_t)M * N, st));
        GemmOperands o{w16, a16, bias, addend, C, nullptr, N, Kp, Kp, N, N, relu, nullptr};
        // ROWS without split-K goes through gemm_umma_rows, i.e. also through the CTA-pair kernel when SRNN_GEMM_PAIR selects it;
        // SRNN_GEMM_HOOK_SWAP_PAIR routes the swap-AB form through the CTA-pair kernel (tests)
        int rc = (rows && !split && N % 8 == 0 && getenv("SRNN_GEMM_HOOK_WIDE_PAIR") && gemm_umma_pair_wide_ok(N, M, true))
                     ? gemm_umma_pair_wide(o, M, Kp, st)
                 : (!rows && !split && getenv("SRNN_GEMM_HOOK_SWAP_PAIR")) ? gemm_umma_swap_pair(o, M, Kp, st)
                 : (rows && !split && N % 8 == 0) ? gemm_umma_rows(o, M, Kp, 1, nullptr, st)
                                                : gemm_umma_ex(&o, 1, M, Kp, bm, bn, rows, split ? 3 : 1, scratch, st);
        cudaFreeAsync(a16, st);
        cudaFreeAsync(w16, st);
        if (scratch) cudaFreeAsync(scratch, st);
        return rc;
    }
    return fail(SRNN_ERR_UNSUPPORTED, "gemm: mode %d not available", mode);
}

}  // extern "C"
