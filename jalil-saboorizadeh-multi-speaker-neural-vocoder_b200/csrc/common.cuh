// Shared declarations of the SampleRNN B200 library (internal; the public C-ABI is include/srnn_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <vector>
#include "srnn_b200.h"

namespace srnn {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

int fail(int code, const char* fmt, ...);

#define SRNN_CUDA(expr)                                                                          \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return ::srnn::fail(SRNN_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,      \
                                cudaGetErrorString(_e));                                         \
    } while (0)

#define SRNN_TRY(expr)                                                                           \
    do {                                                                                         \
        int _r = (expr);                                                                         \
        if (_r != SRNN_OK) return _r;                                                            \
    } while (0)

// every kernel launch goes through here so srnn_launch_count() is the library's own claim
#define SRNN_LAUNCH(kernel, grid, block, smem, stream, ...)                                      \
    do {                                                                                         \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                              \
        ::srnn::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
        cudaError_t _e = cudaPeekAtLastError();                                                  \
        if (_e != cudaSuccess)                                                                   \
            return ::srnn::fail(SRNN_ERR_CUDA, "%s:%d launch %s -> %s", __FILE__, __LINE__,      \
                                #kernel, cudaGetErrorString(_e));                                \
    } while (0)

// Programmatic dependent launch: while g_pdl is set (generation schedule, consecutive tier kernels), the launch carries
// cudaLaunchAttributeProgrammaticStreamSerialization, so the grid may start -- barrier init, TMEM allocation, tensor-map
// prefetch -- while its predecessor drains; every kernel launched through this macro executes griddepcontrol.wait before it
// touches global memory and griddepcontrol.launch_dependents at its start (both are no-ops in a plain launch).
extern thread_local int g_pdl;
#define SRNN_LAUNCH_PDL(kernel, grid_, block_, smem_, stream_, ...)                                  \
    do {                                                                                         \
        if (::srnn::g_pdl) {                                                                     \
            cudaLaunchConfig_t _cfg = {};                                                        \
            _cfg.gridDim = (grid_);                                                              \
            _cfg.blockDim = (block_);                                                            \
            _cfg.dynamicSmemBytes = (smem_);                                                     \
            _cfg.stream = (stream_);                                                             \
            cudaLaunchAttribute _at[1];                                                          \
            _at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                      \
            _at[0].val.programmaticStreamSerializationAllowed = 1;                               \
            _cfg.attrs = _at;                                                                    \
            _cfg.numAttrs = 1;                                                                   \
            cudaError_t _e = cudaLaunchKernelEx(&_cfg, kernel, __VA_ARGS__);                     \
            ::srnn::g_launches.fetch_add(1, std::memory_order_relaxed);                          \
            if (_e != cudaSuccess)                                                               \
                return ::srnn::fail(SRNN_ERR_CUDA, "%s:%d launch %s -> %s", __FILE__, __LINE__,  \
                                    #kernel, cudaGetErrorString(_e));                            \
        } else {                                                                                 \
            SRNN_LAUNCH(kernel, grid_, block_, smem_, stream_, __VA_ARGS__);                        \
        }                                                                                        \
    } while (0)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
#endif

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- packed weights of one FrameLevelRNN tier (model.py:67-178) --------------------------------
struct TierPacked {
    int fs = 0;        // frame_size (upsampling ratio)
    int n = 0;         // n_frame_samples
    int kin = 0;       // K of the input GEMM: n (+ cond_dim + spk_dim on the top tier)
    bool top = false;
    float* w_in = nullptr;    // (H, kin)   [input_expand | cond_expand | spk_expand . E^T]
    float* w_in_t = nullptr;  // (kin, H)   transpose of w_in for the fused generation-time input kernel
    float* b_in = nullptr;    // (H)        summed biases
    float* w_ih[SRNN_MAX_RNN] = {};   // (3H, H)
    float* w_hh[SRNN_MAX_RNN] = {};
    float* b_ih[SRNN_MAX_RNN] = {};
    float* b_hh[SRNN_MAX_RNN] = {};
    float* w_up = nullptr;    // (fs*H, H)  row j*H+o = conv_t.weight[:, o, j]   (nn.py:33-43)
    float* b_up = nullptr;    // (fs*H)     j*H+o -> upsampling.bias[o, j]
    float* h0 = nullptr;      // (n_rnn, H)
    // bf16 copies for the tcgen05 path (present when H % 64 == 0)
    __nv_bfloat16* w_ih16[SRNN_MAX_RNN] = {};
    __nv_bfloat16* w_hh16[SRNN_MAX_RNN] = {};
    __nv_bfloat16* w_up16 = nullptr;
    // transposed bf16 copies (H, 3H) / (H, fs*H): the A operands of the dIn = dOut . W GEMMs of the backward pass
    __nv_bfloat16* w_ih16_t[SRNN_MAX_RNN] = {};
    __nv_bfloat16* w_hh16_t[SRNN_MAX_RNN] = {};
    __nv_bfloat16* w_up16_t = nullptr;
    // generation-time fold of the input expansion into the first GRU layer (ensure_gi_fold):
    //   gi_0 = W_ih0 (W_in a + b_in + upper) + b_ih0 = G a + W_ih0 upper + b_gi0,   G = W_ih0 W_in,  b_gi0 = b_ih0 + W_ih0 b_in
    float* g_in_t = nullptr;  // (kin, 3H) = G^T: consecutive threads read consecutive gate rows
    float* b_gi0 = nullptr;   // (3H)
    // ... and, for a tier fed by the tier above, through that tier's upsampling for its FIRST frame (j = 0):
    //   W_ih0 upper_0 + b_gi0 = (W_ih0 W_up[0]) h_above + (W_ih0 b_up[0] + b_gi0): one GEMM straight from the upper tier's state,
    //   so the upper tier's upsampling itself leaves the serial path
    __nv_bfloat16* gup0_16 = nullptr;   // (3H, H) bf16
    float* b_gup0 = nullptr;            // (3H)
    // split-bf16 copies (n_feat, 3K) of SRNN_MODE_BF16X3, packed lazily on the first use of that mode (ensure_x3)
    __nv_bfloat16* w_ih3[SRNN_MAX_RNN] = {};
    __nv_bfloat16* w_hh3[SRNN_MAX_RNN] = {};
    __nv_bfloat16* w_up3 = nullptr;
};

struct Arena {
    std::vector<void*> ptrs;
    int alloc(void** p, size_t bytes);
    void release();
};

}  // namespace srnn

namespace srnn {
// Buffers of the last teacher-forced forward pass (inside the context's scratch): the saved activations of the
// backward pass.  Row index of frame-level tensors = b*F + f; of sample-level tensors = b*T + t.
struct FwdPlan {
    bool valid = false;
    int B = 0, T = 0, mode = 0, reset_mask = 0;
    const void* cond = nullptr;   // caller's conditioner / speaker tensors of that pass (must stay alive until backward)
    int cond_is_f64 = 0;
    const int64_t* spk = nullptr;
    uint8_t* seq = nullptr;                                   // (B, lookback+T-1)
    float* A[SRNN_MAX_TIERS] = {};                            // (M, kin) assembled tier inputs
    float* X[SRNN_MAX_TIERS] = {};                            // (M, H) GRU layer-0 input
    __nv_bfloat16* X16[SRNN_MAX_TIERS] = {};
    float* GI[SRNN_MAX_TIERS][SRNN_MAX_RNN] = {};             // (M, 3H) W_ih x + b_ih
    float* GH[SRNN_MAX_TIERS][SRNN_MAX_RNN] = {};             // (M, 3H) W_hh h_{t-1} + b_hh
    float* Y[SRNN_MAX_TIERS][SRNN_MAX_RNN] = {};              // (M, H) layer outputs h_t
    __nv_bfloat16* Y16[SRNN_MAX_TIERS][SRNN_MAX_RNN] = {};
    float* H0[SRNN_MAX_TIERS] = {};                           // (n_rnn, B, H) initial hidden state used
    __nv_bfloat16* H016[SRNN_MAX_TIERS] = {};
    float* UP[SRNN_MAX_TIERS] = {};                           // (M, fs*H) upsampled conditioning for the tier below
    __nv_bfloat16* UP16 = nullptr;                            // bf16 mode: tier 0's conditioning of the MLP, stored in bf16
    float* X1 = nullptr;                                      // (B*T, H) relu(gather + c0)        [fp32 mode]
    float* X2 = nullptr;                                      // (B*T, H) relu(hidden)             [fp32 mode]
    __nv_bfloat16* X1h = nullptr;                             // bf16 mode
    __nv_bfloat16* X2h = nullptr;
    __nv_bfloat16* S3 = nullptr;                              // SRNN_MODE_BF16X3: split copy (rows, 3H) of the current GEMM input
    size_t bytes = 0;
};
}  // namespace srnn

struct srnn_ctx {
    srnn_config cfg{};
    srnn::FwdPlan fwd;
    int lookback = 0;
    int H = 0, Q = SRNN_Q, FS0 = 0;
    bool packed = false;
    srnn::TierPacked tiers[SRNN_MAX_TIERS];
    float* tbl = nullptr;     // (FS0, Q, H) folded embedding o conv table
    float* w_hid = nullptr;   // (H, H)
    float* b_hid = nullptr;
    float* w_out = nullptr;   // (Q, H)
    float* b_out = nullptr;
    float* lut = nullptr;     // (Q) 2*dequantize(q)
    float* loss_partial = nullptr;
    double timed_ms = 0;          // see srnn_timed_kernel
    long long timed_launches = 0;
    bool has_bf16 = false;    // tcgen05 path available (H % 64 == 0)
    __nv_bfloat16* tbl16 = nullptr;
    __nv_bfloat16* w_hid16 = nullptr;
    __nv_bfloat16* w_out16 = nullptr;
    __nv_bfloat16* w_hid16_t = nullptr;   // (H, H) transposed
    __nv_bfloat16* w_out16_t = nullptr;   // (H, Q) transposed
    __nv_bfloat16* w_hid3 = nullptr;      // (H, 3H) / (Q, 3H) split-bf16 copies of SRNN_MODE_BF16X3
    __nv_bfloat16* w_out3 = nullptr;
    bool x3_valid = false;                // the *3 copies match the packed fp32 weights
    bool gi_fold_valid = false;           // tiers[].g_in_t / b_gi0 match the packed weights
    // instantiated generation graph of the last srnn_generate shape: reused while every baked-in value (batch, mode, schedule
    // switches, caller and scratch pointers) is unchanged -- capture + instantiation cost ~1 ms, which short utterances and
    // small batches feel (the graph itself is position independent: the sample index lives in a device counter)
    struct GenGraph {
        unsigned long long key[20] = {};
        int nkey = 0;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        long long nodes = 0, reuses = 0;
    } gen_graph;
    unsigned* gru_ctr = nullptr;  // frame-barrier counter of the persistent GRU kernels
    // recorded by srnn_predict_bwd as the gradients become final: [0] sample-level MLP + embedding, [1 + 2i] tier i's
    // upsampling, [2 + 2i] the rest of tier i (srnn_bwd_wait_stage: data-parallel all-reduce overlapping the backward pass)
    cudaEvent_t ev_stage[2 * SRNN_MAX_TIERS + 1] = {};
    int n_sms = 0;
    srnn::Arena weights;      // freed on destroy
    // grow-only scratch for predict / generate
    void* ws = nullptr;
    size_t ws_bytes = 0;
    int device = 0;
};

namespace srnn {

int ensure_ws(srnn_ctx* ctx, size_t bytes);

// ---- fp32 kernels (kernels_f32.cu) --------------------------------------------------------------
int tier_assemble_f32(const float* prev, int n, const float* cond, int cond_dim, const int64_t* spk, int spk_dim, int rows, int F,
                      float* A, int kin, cudaStream_t st);
struct L2PrefetchArgs {
    const void* ptr[8];
    size_t bytes[8];
    int n;
};
int prefetch_l2(const L2PrefetchArgs& a, int ctas, cudaStream_t st);
int gemm_f32(int M, int N, int K, const float* A, int lda, const float* B, int ldb, const float* bias,
             const float* add, int ldadd, int relu, float* C, int ldc, cudaStream_t st,
             __nv_bfloat16* C16 = nullptr);
int gemm_f32_batched(int batch, int M, int N, int K, const float* A, int lda, long long sA, const float* B, int ldb, long long sB,
                     float* C, int ldc, long long sC, cudaStream_t st);
int wn_fold(const srnn_conv_params& p, float* out, int rows, int cols, cudaStream_t st);
int copy_f32(const float* src, float* dst, size_t n, cudaStream_t st);
int fill_u8(uint8_t* dst, uint8_t v, size_t n, cudaStream_t st);
int build_lut(float* lut, int q_levels, int ulaw, cudaStream_t st);
int i64_to_u8(const int64_t* src, uint8_t* dst, size_t n, cudaStream_t st);
int quantize_samples(const float* x, int rows, int cols, long long ld, int q_levels, int ulaw, int64_t* q, cudaStream_t st);
int pack_top_in(const float* w_in, const float* w_c, const float* w_s, const float* emb, const float* b_in,
                const float* b_c, const float* b_s, float* w_out, float* b_out, int H, int n, int cond_dim,
                int spk_dim, cudaStream_t st);
int pack_up(const float* wf, const float* bias, float* w_up, float* b_up, int H, int k, cudaStream_t st);
int unpack_up_grad(const float* dwp, const float* dbp, float* dwf, float* dbias, int H, int k, cudaStream_t st);
int transpose_mlp_in(const float* w, float* wt, int H, int Q, int FS, cudaStream_t st);
// A[b*F+f, :] = [lut[seq[b, off + f*n + i]] (i<n) | cond[crow(b), f0+f, :] | onehot(spk[crow(b)])]
int frame_input(const uint8_t* seq, int seq_ld, int off, const int* step_base, int n, int B, int F,
                const void* cond, int cond_is_f64, int cond_rows, int cond_frames,
                const int64_t* spk, int cond_dim, int spk_dim, const float* lut, float* A, int kin, bool top,
                cudaStream_t st);
int tier_input_gen(const uint8_t* seq, int seq_ld, int off, const int* step_base, int n, int B, const float* cond,
                   int cond_rows, int cond_frames, const int64_t* spk, int cond_dim, int spk_dim, const float* lut,
                   const float* w_in_t, const float* b_in, const float* upper, int up_ld, float* X,
                   __nv_bfloat16* X16, int H, int kin, bool top, cudaStream_t st);
int tier_input_split(bool tail, const uint8_t* seq, int seq_ld, int off, const int* step_base, int n, int B, const float* cond,
                     int cond_rows, int cond_frames, const int64_t* spk, int cond_dim, const float* lut, const float* w_in_t,
                     const float* b_in, float* partial, float* X, __nv_bfloat16* X16, int H, int kin, int k_lo, int k_hi,
                     cudaStream_t st);
int transpose_f32(const float* src, float* dst, int rows, int cols, cudaStream_t st);
// GRU cell tail (model.py:244): h' from gi (+bias already in), gh (+bias already in), h
int gru_gates(const float* gi, int gi_ld, const float* gh, int gh_ld, const float* h_prev, int hp_ld,
              float* h_out, int ho_ld, float* h_out2, int B, int H, cudaStream_t st,
              __nv_bfloat16* h16 = nullptr, int h16_ld = 0);
int nll_bits(const float* logp, const int64_t* target, int rows, float* partial, int n_partial, float* out,
             cudaStream_t st);
int bcast_rows(const float* src, float* dst, int B, int H, cudaStream_t st);
// x1[r,:] = relu(sum_j Tbl[j][seq[b, off + t + j]] + upper[r,:])
int mlp_gather(const uint8_t* seq, int seq_ld, int off, const int* step_base, const float* tbl,
               const float* upper, long long up_bstride, long long up_tstride, float* x1, int B, int T, int H,
               int FS, cudaStream_t st);
int mlp_gather_bf16(const uint8_t* seq, int seq_ld, int off, const int* step_base, const __nv_bfloat16* tbl,
                    const float* upper, long long up_bstride, long long up_tstride, __nv_bfloat16* x1, int B, int T,
                    int H, int FS, cudaStream_t st, const __nv_bfloat16* upper16 = nullptr);
int logsoftmax_rows(float* x, int rows, cudaStream_t st);
// generation tail: logits (B, 256) -> [logp] -> defined sampler -> seq[b, pos]
int softmax_sample(const float* logits, const float* uniforms, int u_ld, uint8_t* seq, int seq_ld, int pos_off,
                   int lookback, const int* step_base, float* logp_out, long long logp_bstride, int B,
                   cudaStream_t st);
int sample_rows(const float* p, const float* u, int rows, int* idx, cudaStream_t st);
int dequant_audio(const uint8_t* seq, int seq_ld, int off, const float* lut, uint8_t* samples, float* audio,
                  int B, int T, cudaStream_t st);
int add_int(int* p, int v, cudaStream_t st);
int cond_chain_fwd(int n_layers, const int* dims, const float* const* w, const float* const* b, const float* cond, int rows,
                   float* out, cudaStream_t st);

// ---- backward pass + optimizer (backward.cu) --------------------------------------------------------
size_t backward_scratch_bytes(const srnn_ctx* ctx, int B, int T);
// nll_target != null: the upstream gradient is sequence_nll_loss_bits' (fused, dlogp ignored); nll_gscale = device scalar or null
int predict_bwd_f32(srnn_ctx* ctx, const float* logp, const float* dlogp, const srnn_params* P, const srnn_params* G,
                    cudaStream_t st, const int64_t* nll_target = nullptr, const float* nll_gscale = nullptr);
size_t backward_scratch_bytes_bf16(const srnn_ctx* ctx, int B, int T);
int predict_bwd_bf16(srnn_ctx* ctx, const float* logp, const float* dlogp, const srnn_params* P, const srnn_params* G,
                     cudaStream_t st, const int64_t* nll_target = nullptr, const float* nll_gscale = nullptr);
int gru_seq_bwd_f32(int B, int Fr, int H, const float* GI, const float* GH, const float* Y, const float* h0, const float* dY,
                    const float* w_hh, float* dGI, float* dGH, float* dh0, float* scratch, cudaStream_t st);
int clamp_adam(int count, float* const* params, const float* const* grads, float* const* m, float* const* v,
               const long long* sizes, float lr, float beta1, float beta2, float eps, int step, float clamp, cudaStream_t st,
               float grad_scale = 1.f);

// ---- tcgen05 / TMA kernels (gemm_umma.cu) ------------------------------------------------------------
int gemm_umma(const __nv_bfloat16* W, int n_feat, const __nv_bfloat16* act, int n_rows, int K, int ld_w, int ld_act,
              const float* bias, const float* addend, int ld_add, float* out_f32, __nv_bfloat16* out_bf16,
              int ld_out, int relu, int bm, int bn, cudaStream_t st);
// persistent fused sample-level kernel (mlp_persist.cu)
struct MlpPersistParams {
    int B, H, FS, nsteps, pos0, lookback, Lseq, T;
    const int* step_base;
    uint8_t* seq;                 // (B, Lseq) quantised samples (read + written)
    const float* c0;              // tier-0 output (B, FS*H): conditioning of sample phase p at [b][p*H + f]
    const __nv_bfloat16* tbl;     // (FS, 256, H) folded embedding-o-conv table
    const float* b_hid;
    const float* b_out;
    __nv_bfloat16* x1;            // (RG*32, H) exchange buffer
    float* part;                  // (RG, NS, 32, 256) split-K partial logits
    unsigned* ctr;                // (RG) group-barrier counters, zeroed by the launcher
    const float* uniforms;        // (T, u_ld): row t holds the uniforms of step t, utterance b at column b
    int u_ld;
    float* logp_out;              // (B, T, 256) or null
    long long* trace;             // optional (nsteps, 10) clock64 stamps of CTA 0 (SRNN_TRACE=1), else null
    int dbg = 0;                  // development switches of k_mlp_cluster (SRNN_MC_DBG), 0 in production
    // k_mlp_cluster: (2, rows, H) fp32 table parts sum_{j < FS-2} Tbl[j][.] of the FIRST and SECOND sample of the next launch,
    // written by the gather warps of this one while its last samples are drawn (the samples they need exist three / two steps
    // before the end) and consumed by the next launch, whose first two steps then need c0 + one row per utterance instead of
    // FS-2 table rows on the serial path.  Seeded by mlp_cluster_carry_init for the first frame of a call.  Null: every
    // launch gathers its first samples itself.
    float* pcarry = nullptr;
    long long carry_plane = 0;    // elements between the two planes (rows * H)
};
int mlp_cluster_carry_init(const __nv_bfloat16* tbl, int FS, int H, int rows, int q_zero, float* pcarry, cudaStream_t st);
int mlp_persist_launch(const __nv_bfloat16* w_hid16, const __nv_bfloat16* w_out16, const MlpPersistParams& p,
                       cudaStream_t st);
// cluster form (mlp_cluster.cu, H = 1024): rows_per_cluster = 16 or 24; p.x1 holds ceil(B / rows) * rows rows; part / ctr unused
bool mlp_cluster_supported(int H, int FS, int B, int max_clusters);
int mlp_cluster_rows(int B, int max_clusters);
int mlp_cluster_max_clusters();
int mlp_cluster_launch(const __nv_bfloat16* w_hid16, const __nv_bfloat16* w_out16, const MlpPersistParams& p, int rows_per_cluster,
                       cudaStream_t st);
extern const char* g_sample_kernel;
size_t mlp_persist_smem(int H);
bool mlp_persist_supported(int H, int FS, int B, int n_sms);
struct GemmOperands {
    const __nv_bfloat16* W;       // (n_feat, K) weights, the UMMA A operand
    const __nv_bfloat16* act;     // (n_rows, K) activations, the UMMA B operand
    const float* bias;            // (n_feat) or null
    const float* addend;          // (n_rows, ld_add) or null
    float* out_f32;               // (n_rows, ld_out) or null
    __nv_bfloat16* out_bf16;      // (n_rows, ld_out) or null
    int n_feat, ld_w, ld_act, ld_add, ld_out, relu;
    const __nv_bfloat16* mask;    // optional (n_rows, ld_out): keep the value where mask > 0 (ReLU backward), bf16 output
};
int transpose_to_bf16(const float* src, int rows, int cols, long long ld_src, __nv_bfloat16* dst, int ld_dst, cudaStream_t st);
int transpose_to_bf16(const __nv_bfloat16* src, int rows, int cols, long long ld_src, __nv_bfloat16* dst, int ld_dst, cudaStream_t st);
int gemm_umma_multi(const GemmOperands* ops, int nprob, int n_rows, int K, int bm, int bn, cudaStream_t st);
void gemm_umma_set_cta_cap(int cap);   // > 0: following single-CTA-kernel launches of this thread use at most cap CTAs; 0 = off
// persistent GRU recurrence over all frames of one layer (gru_persist.cu)
bool gru_persist_supported(int B, int H, int n_sms);
// generation-time fused cell: gi GEMM + gate math in one launch (gru_persist.cu)
bool gru_cell_gen_supported(int H);
int gru_cell_gen(int B, int H, const __nv_bfloat16* x16, const __nv_bfloat16* w_ih16, const float* b_ih, const float* GH, float* h,
                 __nv_bfloat16* h16, cudaStream_t st);
// generation-time first GRU layer with the input expansion folded in (TierPacked::g_in_t): gi = gipre + G[:, k_lo:k_hi) . samples,
// then the gate math; gipre (B rows, leading dimension gipre_ld) already holds W_ih0 upper + b_gi0 (or the shadow part of the
// top tier); samples = lut[seq[b][start + k]] for the sample columns k in [k_lo, k_hi)
bool gru_cell_lite_supported(int n_sample_columns);
int gru_cell_lite(int B, int H, const float* gipre, long long gipre_ld, const float* g_in_t, int k_lo, int k_hi, const uint8_t* seq,
                  int seq_ld, int start_static, const int* step_base, const float* lut, const float* GH, float* h,
                  __nv_bfloat16* h16, cudaStream_t st);
int gru_persist_fwd(int B, int F, int H, const float* GI, const __nv_bfloat16* w_hh16, const float* b_hh, const float* h0,
                    const __nv_bfloat16* h0_16, float* GH, float* Y, __nv_bfloat16* Y16, float* h_last, unsigned* ctr,
                    cudaStream_t st);
int gru_persist_bwd(int B, int F, int H, const float* GI, const float* GH, const float* Y, const float* h0, const float* dY,
                    const __nv_bfloat16* w_hh16_t, float* dGI, float* dGH, __nv_bfloat16* dGI16, __nv_bfloat16* dGH16,
                    float* dh0, unsigned* ctr, cudaStream_t st, float* bias_part = nullptr, float* db_ih = nullptr,
                    float* db_hh = nullptr);
int gemm_umma_ex(const GemmOperands* ops, int nprob, int n_rows, int K, int bm, int bn, bool rows, int ksplit,
                 float* split_scratch, cudaStream_t st);
int gemm_umma_rows(const GemmOperands& o, int n_rows, int K, int ksplit, float* split_scratch, cudaStream_t st);
int gemm_umma_tn(const __nv_bfloat16* X, int ld_x, const __nv_bfloat16* Y, int ld_y, int M, int N, int Ktot, float* out,
                 int ld_out, int ksplit, float* split_scratch, cudaStream_t st);
bool gemm_umma_swap_pair_ok(int n_feat, int n_rows);
int gemm_umma_swap_pair(const GemmOperands& o, int n_rows, int K, cudaStream_t st);
// one-wave CTA-pair form with (256 + 32)-feature tiles: the generation-time upsampling of the lower tier
bool gemm_umma_pair_wide_ok(int n_feat, int n_rows, bool force = false);
int gemm_umma_pair_wide(const GemmOperands& o, int n_rows, int K, cudaStream_t st);
int sum_splits(const float* part, int splits, size_t n, size_t stride, float* out, cudaStream_t st);
int f32_to_bf16_pad(const float* src, int rows, int cols, int ld_src, __nv_bfloat16* dst, int rows_p, int cols_p,
                    cudaStream_t st);
// All bf16 operand copies of the packed fp32 weights in ONE launch (was 19 conversion + 12 transpose launches after every
// optimizer step): for each listed matrix src (rows, cols) fp32 -> dst (rows, cols) bf16 and / or dst_t (cols, rows) bf16 (+ an fp32 copy).
struct PackBf16Item {
    const float* src;
    __nv_bfloat16* dst;       // may be null
    __nv_bfloat16* dst_t;     // may be null
    float* dst32;             // optional fp32 copy (the GRU matrices are taken straight from the caller's tensors)
    int rows, cols;
    int tile0;                // first tile of this matrix in the launch's tile list (filled by pack_bf16_multi)
};
constexpr int PACK_BF16_MAX = 40;
int pack_bf16_multi(PackBf16Item* items, int n, cudaStream_t st);
// SRNN_MODE_BF16X3 operand form: fp32 (rows, K) -> bf16 (rows, 3K), every value split into hi = bf16(x), lo = bf16(x - hi) and
// laid out along K as [hi | hi | lo] (activations, weight_order = 0) or [hi | lo | hi] (weights, weight_order = 1), so that ONE
// tcgen05 GEMM over K' = 3K accumulates Wh.xh + Wl.xh + Wh.xl in fp32 (the dropped Wl.xl term is ~2^-16 relative)
int split3_bf16(const float* src, long long rows, int K, long long ld_src, __nv_bfloat16* dst, int weight_order, cudaStream_t st);
// the same split for MN-major operands (K is the slow dimension): three planes of n elements, [hi; hi; lo] or [hi; lo; hi]
int split3_planes_bf16(const float* src, size_t n, __nv_bfloat16* dst, int weight_order, cudaStream_t st);
// SRNN_MODE_BF16X3 contraction C (rows, n_feat) = A (rows, K) . W^T + bias [relu] with fp32 in / out: the activations are split
// into s3 (rows, 3K) and meet the pre-split weights W3 (n_feat, 3K) in one tcgen05 GEMM.  bm / bn = 0: chosen from the shape.
int gemm_x3(int rows, int n_feat, int K, const float* A, long long lda, const __nv_bfloat16* W3, const float* bias, int relu,
            float* C, int ldc, __nv_bfloat16* s3, int bm, int bn, cudaStream_t st);

}  // namespace srnn
