/*
 * srnn_b200.h -- C-ABI of the B200-native SampleRNN hot path.
 *
 * The reference (mahdeslami11/jalil-saboorizadeh-Multi-speaker-Neural-Vocoder) has no FFI layer: its
 * boundary is the Python class API of model.py (SampleRNN / Predictor / Generator) plus the
 * checkpoint layout.  This header is the seam a maintainer binds underneath those classes
 * (ctypes stub in INTEGRATION.md); every entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain C types only; all pointers are DEVICE pointers unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - every call returns 0 on success, <0 on error (text via srnn_last_error()); nothing throws;
 *   - no call synchronises the device; work is enqueued on `stream`;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns SRNN_ERR_CUDA.
 */
#ifndef SRNN_B200_H
#define SRNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SRNN_API __attribute__((visibility("default")))
#else
#define SRNN_API
#endif

#define SRNN_MAX_TIERS 4
#define SRNN_MAX_RNN   4
#define SRNN_Q         256   /* q_levels supported by the kernels (train.py:36 default) */
#define SRNN_MAX_CHAIN 8     /* layers of the bottle-neck conditioner chain */

#define SRNN_OK            0
#define SRNN_ERR_ARG      -1
#define SRNN_ERR_CUDA     -2
#define SRNN_ERR_STATE    -3
#define SRNN_ERR_UNSUPPORTED -4

/* arithmetic mode of the GEMM-shaped work */
#define SRNN_MODE_FP32   0   /* fp32 FFMA everywhere: the 1e-3 "fp32 parity" gate                        */
#define SRNN_MODE_BF16   1   /* bf16 operands on tcgen05 tensor cores, fp32 accumulate: the speed mode */
#define SRNN_MODE_BF16_GRAPH 2 /* generation only: bf16 arithmetic, one tcgen05 GEMM launch per contraction
                                  (no persistent sample kernel); kept as an A/B reference for the fused kernel */
#define SRNN_MODE_BF16X3 3   /* tensor-core parity mode: the control flow of SRNN_MODE_FP32 with every dense contraction
                                (GRU projections, upsampling, MLP hidden / output) on tcgen05 as a split-bf16 product
                                W.x ~= Wh.xh + Wl.xh + Wh.xl (hi = bf16(v), lo = bf16(v - hi); one GEMM over K' = 3K, fp32
                                accumulation): holds the 1e-3 logit gate of the fp32 mode at tensor-core speed.  Forward
                                paths (srnn_predict_fwd, srnn_generate); a backward pass after it runs the fp32 kernels. */

typedef struct srnn_ctx srnn_ctx;

/* Constructor arguments of SampleRNN (model.py:20). */
typedef struct {
    int32_t n_tiers;                        /* len(frame_sizes) = number of FrameLevelRNN tiers            */
    int32_t frame_sizes[SRNN_MAX_TIERS];    /* frame_sizes[0] = lowest tier (model.py:46-55)               */
    int32_t n_rnn;                          /* GRU layers per tier                                         */
    int32_t dim;                            /* H                                                           */
    int32_t q_levels;                       /* must equal SRNN_Q                                           */
    int32_t cond_dim;                       /* conditioner width (43, or 86 with look-ahead)               */
    int32_t spk_dim;                        /* speaker count = embedding width (model.py:103-106)          */
    int32_t ulaw;                           /* 1: utils.udequantize, 0: utils.linear_dequantize            */
} srnn_config;

/* One Conv1d / ConvTranspose1d as stored in Predictor.state_dict():
 * either `weight`, or the weight_norm pair `weight_g` + `weight_v`; `bias` may be NULL. fp32. */
typedef struct {
    const float* weight;
    const float* weight_g;
    const float* weight_v;
    const float* bias;
} srnn_conv_params;

/* FrameLevelRNN parameters (model.py:67-178), reference layouts, fp32. */
typedef struct {
    const float* h0;                        /* (n_rnn, H)                                                  */
    srnn_conv_params input_expand;          /* (H, n_frame_samples, 1)                                     */
    srnn_conv_params cond_expand;           /* (H, cond_dim, 1)   top tier only                            */
    const float* spk_embedding;             /* (spk_dim, spk_dim) top tier only                            */
    srnn_conv_params spk_expand;            /* (H, spk_dim, 1)    top tier only                            */
    const float* weight_ih[SRNN_MAX_RNN];   /* (3H, H) gate row blocks r, z, n                             */
    const float* weight_hh[SRNN_MAX_RNN];
    const float* bias_ih[SRNN_MAX_RNN];     /* (3H)                                                        */
    const float* bias_hh[SRNN_MAX_RNN];
    srnn_conv_params upsampling;            /* conv_t weight (H_in, H_out, k); .bias = upsampling.bias (H, k) */
} srnn_tier_params;

typedef struct {
    srnn_tier_params tiers[SRNN_MAX_TIERS];
    const float* embedding;                 /* (Q, Q)        model.py:274-277                              */
    srnn_conv_params mlp_input;             /* (H, Q, FS0), no bias  model.py:279-284                      */
    srnn_conv_params mlp_hidden;            /* (H, H, 1)     model.py:287-293                              */
    srnn_conv_params mlp_output;            /* (Q, H, 1)     model.py:295-301                              */
} srnn_params;

/* ---- lifecycle -------------------------------------------------------------------------------- */
SRNN_API const char* srnn_last_error(void);
SRNN_API int srnn_version(void);
/* Replaces SampleRNN.__init__ (model.py:20-62) as far as device state goes. */
SRNN_API int srnn_create(const srnn_config* cfg, srnn_ctx** out);
SRNN_API int srnn_destroy(srnn_ctx* ctx);
SRNN_API int srnn_lookback(const srnn_ctx* ctx);                 /* SampleRNN.lookback  model.py:60-62            */

/* Snapshot the parameters into the kernels' layouts: weight-norm fold (model.py:119-121,130-131,
 * 177-178,303-306 recompute it every forward), K-concatenated top-tier input matrix
 * [input_expand | cond_expand | spk_expand . spk_embedding^T], phase-major upsampling matrix
 * (nn.py:33-43), the exact embedding-o-conv fold Tbl[j][q][:] = W_in[:,:,j] . E[q,:]
 * (model.py:311-317), the dequantiser LUT (utils.py:18-19,39-63) and bf16 copies. */
SRNN_API int srnn_pack_weights(srnn_ctx* ctx, const srnn_params* params, void* stream);

/* ---- Predictor.forward (model.py:357-436) ----------------------------------------------------- */
/* input_seq (B, lookback+T-1) int64 in [0,Q); cond (B, T/lookback, cond_dim) fp32 or fp64 (cond_is_f64);
 * spk (B) int64; hidden_io[t] (n_rnn, B, H) fp32 per tier, always overwritten with the carry
 * (Runner.run_rnn model.py:348); reset_mask bit t set = tier t starts from its h0 (hidden_states[rnn] is
 * None, model.py:222-228), clear = starts from hidden_io[t]; logp_out (B, T, Q) fp32 log-probabilities. */
SRNN_API int srnn_predict_fwd(srnn_ctx* ctx, int32_t B, int32_t T, const int64_t* input_seq, const void* cond,
                     int32_t cond_is_f64, const int64_t* spk, float* const* hidden_io, int32_t reset_mask,
                     float* logp_out, int32_t mode, void* stream);

/* Backward of the last srnn_predict_fwd on this context (what torch autograd does for the reference at
 * trainer/__init__.py:103): logp = that call's output (B,T,Q), dlogp = dL/dlogp, params = the raw tensors given to
 * srnn_pack_weights, grads = the same struct pointing at WRITABLE gradient buffers of identical shapes (null entries
 * are skipped; non-null ones are overwritten).  Hidden-state carries are detached (model.py:348): h0 receives a
 * gradient only for tiers that started from it in that forward pass, zeros otherwise (torch-0.4 zero_grad semantics,
 * SURVEY App. C #12).  Runs in the arithmetic mode of that forward pass (SRNN_MODE_FP32 or SRNN_MODE_BF16). */
SRNN_API int srnn_predict_bwd(srnn_ctx* ctx, const float* logp, const float* dlogp, const srnn_params* params,
                              const srnn_params* grads, void* stream);

/* srnn_predict_bwd with the loss folded in: backward of sequence_nll_loss_bits(logp, target) (nn.py:66-70,
 * trainer/__init__.py:102-103).  target (B, T) int64 in [0,Q); gscale = device pointer to the upstream scalar gradient of
 * the loss (NULL = 1).  dL/dlogits = (exp(logp) - onehot(target)) * g * log2(e) / (B*T) is produced by one kernel straight
 * into the buffers the output-layer GEMMs read: no dense (B,T,Q) dL/dlogp is materialised. */
SRNN_API int srnn_predict_bwd_nll(srnn_ctx* ctx, const float* logp, const int64_t* target, const float* gscale,
                                  const srnn_params* params, const srnn_params* grads, void* stream);

/* Data-parallel training (no counterpart in the single-GPU reference): `stream` waits until the last srnn_predict_bwd on
 * this context has finalised every gradient below the top tier (MLP, embedding, lower tiers), so that their NCCL all-reduce
 * can overlap the top tier's backward pass. */
SRNN_API int srnn_bwd_wait_early(srnn_ctx* ctx, void* stream);
/* Finer-grained form.  A backward pass finalises gradients in this order: stage 0 = sample-level MLP + embedding,
 * stage 1 + 2i = tier i's upsampling (conv_t weight_g / weight_v and bias), stage 2 + 2i = the rest of tier i, lowest tier
 * first; `stream` waits for that stage, so a bucketed NCCL all-reduce can follow the backward pass stage by stage. */
SRNN_API int srnn_bwd_wait_stage(srnn_ctx* ctx, int32_t stage, void* stream);

/* optim.py:10-13 element-wise clamp of every gradient to [-clamp, clamp] fused with torch.optim.Adam's update
 * (train.py:238: betas (0.9, 0.999), eps 1e-8, no weight decay) over `count` tensors in one launch; step = 1, 2, ... */
SRNN_API int srnn_clamp_adam_step(int32_t count, float* const* params, const float* const* grads, float* const* exp_avg,
                                  float* const* exp_avg_sq, const int64_t* sizes, float lr, float beta1, float beta2,
                                  float eps, int32_t step, float clamp, void* stream);

/* The same with every gradient multiplied by grad_scale before the clamp (1 / world_size after a sum all-reduce: the mean
 * over ranks, the clamp and Adam in one pass; SURVEY 8e "average -> clamp -> Adam"). */
SRNN_API int srnn_clamp_adam_step_scaled(int32_t count, float* const* params, const float* const* grads,
                                         float* const* exp_avg, float* const* exp_avg_sq, const int64_t* sizes, float lr,
                                         float beta1, float beta2, float eps, int32_t step, float clamp, float grad_scale,
                                         void* stream);

/* sequence_nll_loss_bits (nn.py:66-70): loss_out (1 device float) = -mean_r logp[r, target[r]] * log2(e), rows = B*T;
 * deterministic two-stage reduction. */
SRNN_API int srnn_nll_loss_bits(srnn_ctx* ctx, const float* logp, const int64_t* target, int32_t rows, float* loss_out,
                                void* stream);

/* ---- Generator.__call__ (model.py:445-520) ---------------------------------------------------- */
/* cond (cond_rows, n_cond, cond_dim) fp32 with cond_rows == 1 (reference form: one conditioner for all
 * sequences, model.py:484-487) or == B (per-utterance extension); spk (cond_rows) int64;
 * uniforms (n_cond*lookback, B) fp32 in [0,1), indexed [t][b] (defined sampler, see srnn_sample_rows);
 * samples_out (B, n_cond*lookback) uint8 quantised samples; audio_out (same shape) fp32 dequantised
 * audio = what the reference returns (model.py:520), may be NULL; logp_out (B, T, Q) fp32 per-step
 * log-probs for parity tests, may be NULL.  The whole autoregressive loop runs on the device. */
SRNN_API int srnn_generate(srnn_ctx* ctx, int32_t B, int32_t n_cond, const float* cond, int32_t cond_rows,
                  const int64_t* spk, const float* uniforms, uint8_t* samples_out, float* audio_out,
                  float* logp_out, int32_t mode, void* stream);

/* ---- bottle-neck conditioner chain (BASELINE.json configs[4], run_sampleneck.sh:18-19 `--ind_cond_dim 30`) ------------- */
/* The voice-conversion variant replaces the single cond_expand Conv1d(cond_dim -> H) by a chain of k = 1 Conv1d layers
 * cond_dim -> 40 -> 30 -> 20 -> ind_cond_dim -> H.  Its source lives on a branch that is NOT in the reference tree
 * (run_sampleneck.sh:2), so shapes and the ReLU after every chain layer follow the thesis (doc/Barbany_report.pdf 3.2.1):
 * PARITY UNPINNED.  This entry point runs the chain up to ind_cond_dim; its output is the `cond` of a SampleRNN built with
 * cond_dim = ind_cond_dim, whose cond_expand is the last layer.  layers[l]: weight (dims[l+1], dims[l], 1) or the weight_norm
 * pair, bias (dims[l+1]) or NULL; cond (rows, dims[0]) fp32 -> out (rows, dims[n_layers]) fp32;
 * scratch: sum_l dims[l+1]*dims[l] floats (the folded weights). */
typedef struct {
    int32_t n_layers;
    int32_t dims[SRNN_MAX_CHAIN + 1];
    srnn_conv_params layers[SRNN_MAX_CHAIN];
} srnn_cond_chain;
SRNN_API int srnn_cond_chain_fwd(const srnn_cond_chain* chain, const float* cond, int32_t rows, float* out, float* scratch,
                                 void* stream);

/* ---- per-kernel test hooks -------------------------------------------------------------------- */
/* The defined sampler replacing Tensor.multinomial (model.py:517): p (rows, 256) fp32 unnormalised,
 * u (rows) fp32 -> idx (rows) int32.  Bit-exact with oracle/srnn_oracle.py:sample_rows. */
SRNN_API int srnn_sample_rows(const float* p, const float* u, int32_t rows, int32_t* idx, void* stream);
/* Training-data quantiser used by FolderDataset.__getitem__ (dataset.py:249-253): ulaw != 0 -> utils.uquantize
 * (utils.py:33-36,48-51,58-59; the reference's out-of-range index 256 at x == 1.0 is clamped to q_levels-1),
 * else utils.linear_quantize (utils.py:9-15, per-row min/max).  x (rows, cols) fp32 with row stride ld -> q int64. */
SRNN_API int srnn_quantize(const float* x, int32_t rows, int32_t cols, int64_t ld, int32_t q_levels, int32_t ulaw,
                           int64_t* q, void* stream);
/* out (256) fp32 = 2*dequantize(q) (model.py:385,471). */
SRNN_API int srnn_dequant_lut(const srnn_ctx* ctx, float* out, void* stream);
/* SampleLevelMLP.forward (model.py:308-325) with the context's packed weights: prev_samples (B, T+FS0-1) int64 in [0,Q),
 * upper (B, T, H) fp32 conditioning from tier 0 -> logp_out (B, T, Q) fp32 log-probabilities. */
SRNN_API int srnn_mlp_fwd(srnn_ctx* ctx, int32_t B, int32_t T, const int64_t* prev_samples, const float* upper,
                          float* logp_out, int32_t mode, void* stream);
/* FrameLevelRNN.forward (model.py:180-263) of tier `tier` (0 = lowest) on the context's packed weights, F frames:
 * prev_samples (B, F, n_frame_samples) fp32 = 2*dequantize(window) as the reference passes it; upper (B, F, H) fp32 conditioning
 * from the tier above, or null on the top tier, which instead takes cond (B, F, cond_dim) fp32 and spk (B) int64 (one speaker
 * per utterance, model.py:206-217); hidden_io (n_rnn, B, H) fp32: read unless `reset` (then the tier starts from its h0,
 * model.py:222-228) and always overwritten with the new state; out (B, F*frame_size, H) fp32 = the upsampled output.
 * SRNN_MODE_FP32 or SRNN_MODE_BF16X3.  The per-module form of what srnn_predict_fwd runs fused; inference only. */
SRNN_API int srnn_tier_fwd(srnn_ctx* ctx, int32_t tier, int32_t B, int32_t F, const float* prev_samples, const float* upper,
                           const float* cond, const int64_t* spk, float* hidden_io, int32_t reset, float* out, int32_t mode,
                           void* stream);
/* One GRU layer over F frames = `self.rnn(input, hidden)` (model.py:244; torch nn.GRU, gate row blocks r, z, n) given the
 * input projections: gi (B*F, 3H) = W_ih x + b_ih with row b*F+f, w_hh (3H, H), b_hh (3H), h0 (B, H) ->
 * y (B*F, H) = h_f, gh (B*F, 3H) = W_hh h_{f-1} + b_hh (kept for the backward pass), h_last (B, H) or NULL.
 * SRNN_MODE_FP32: frame-by-frame fp32; SRNN_MODE_BF16: the persistent tcgen05 recurrence (B <= 128, H % 64 == 0). */
SRNN_API int srnn_gru_seq_fwd(int32_t B, int32_t F, int32_t H, const float* gi, const float* w_hh, const float* b_hh,
                              const float* h0, float* y, float* gh, float* h_last, int32_t mode, void* stream);
/* Back-propagation through time of srnn_gru_seq_fwd (what autograd does for nn.GRU at trainer/__init__.py:103):
 * dy (B*F, H) -> dgi, dgh (B*F, 3H) gradients of the two projections, dh0 (B, H) or NULL. */
SRNN_API int srnn_gru_seq_bwd(int32_t B, int32_t F, int32_t H, const float* gi, const float* gh, const float* y,
                              const float* h0, const float* w_hh, const float* dy, float* dgi, float* dgh, float* dh0,
                              int32_t mode, void* stream);
/* C (M,N) = A (M,K) . B (N,K)^T + bias (N) [+ addend (M,N)] [relu]; row-major fp32; mode as above. */
SRNN_API int srnn_gemm(int32_t M, int32_t N, int32_t K, const float* A, const float* B, const float* bias,
              const float* addend, int32_t relu, float* C, int32_t mode, void* stream);
/* Measurement hook: CUDA-event time (ms) and launch count of the persistent sample-level kernel over the last
 * srnn_generate call made with the environment variable SRNN_TIME_KERNELS set (direct launches instead of the graph). */
SRNN_API int srnn_timed_kernel(const srnn_ctx* ctx, double* ms, int64_t* launches);
/* Name of the persistent sample-level kernel the last SRNN_MODE_BF16 srnn_generate call ran ("k_mlp_persist", ...). */
SRNN_API const char* srnn_sample_kernel_name(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches). */
SRNN_API int64_t srnn_launch_count(void);
/* Number of srnn_generate calls on this context that replayed the instantiated CUDA graph of the previous call instead of
 * capturing a new one (same batch, length, mode, schedule switches, caller pointers and scratch; SRNN_NO_GRAPH_CACHE=1 disables). */
SRNN_API int64_t srnn_graph_reuse_count(const srnn_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SRNN_B200_H */
