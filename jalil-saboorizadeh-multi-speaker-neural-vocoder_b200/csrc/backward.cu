// Teacher-forced backward pass (the autograd of Predictor.forward, trainer/__init__.py:99-103) and the fused
// clamp + Adam update (optim.py:4-21 + torch.optim.Adam, train.py:238-241).
//
// This is the fp32 (SRNN_MODE_FP32) implementation: every contraction is an FFMA tile GEMM with arbitrary operand
// strides, so the three GEMM forms of back-propagation (dIn = dOut.W, dW = dOut^T.In, and the forward form) share one
// kernel.  The embedding-o-conv fold of the forward pass is differentiated exactly: dTbl is a bucketed row sum of
// dpre1 by sample value (deterministic, no atomics), then folded back onto W_in and E.
#include "common.cuh"

namespace srnn {

// ------------------------------------------------------------------------------------------------
// C[m,n] = sum_k A(m,k) * B(n,k) (+ add[m,n]);  A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk]
// ------------------------------------------------------------------------------------------------
constexpr int SBM = 64, SBN = 64, SBK = 16;

__global__ void __launch_bounds__(256)
k_gemm_f32_strided(int M, int N, int K, const float* __restrict__ A, long long sam, long long sak,
                   const float* __restrict__ B, long long sbn, long long sbk, const float* add, int ldadd, float* C,
                   int ldc, int kper) {
    __shared__ float As[SBK][SBM + 4];
    __shared__ float Bs[SBK][SBN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
    if (kper > 0) {                                // split-K: slice blockIdx.z of the K range -> partial matrix z
        const int kbeg = blockIdx.z * kper;
        A += (long long)kbeg * sak;
        B += (long long)kbeg * sbk;
        K = K - kbeg < kper ? K - kbeg : kper;
        C += (size_t)blockIdx.z * M * ldc;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += SBK) {
        for (int i = tid; i < SBM * SBK; i += 256) {
            int r, c;
            if (sak == 1) { r = i / SBK; c = i % SBK; } else { c = i / SBM; r = i % SBM; }   // walk the unit-stride index fastest
            const int gm = m0 + r, gk = k0 + c;
            As[c][r] = (gm < M && gk < K) ? A[gm * sam + gk * sak] : 0.f;
        }
        for (int i = tid; i < SBN * SBK; i += 256) {
            int r, c;
            if (sbk == 1) { r = i / SBK; c = i % SBK; } else { c = i / SBN; r = i % SBN; }
            const int gn = n0 + r, gk = k0 + c;
            Bs[c][r] = (gn < N && gk < K) ? B[gn * sbn + gk * sbk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SBK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            if (add) v += add[(size_t)gm * ldadd + gn];
            C[(size_t)gm * ldc + gn] = v;
        }
    }
}

static int gemm_s(int M, int N, int K, const float* A, long long sam, long long sak, const float* B, long long sbn,
                  long long sbk, const float* add, int ldadd, float* C, int ldc, cudaStream_t st) {
    if (M <= 0 || N <= 0) return SRNN_OK;
    dim3 grid(cdiv(N, SBN), cdiv(M, SBM));
    SRNN_LAUNCH(k_gemm_f32_strided, grid, 256, 0, st, M, N, K, A, sam, sak, B, sbn, sbk, add, ldadd, C, ldc, 0);
    return SRNN_OK;
}
// Same contraction with the K range cut into slices (one grid.z layer each) when the M x N tile grid alone cannot fill
// the GPU; the partial matrices land in `scratch` (splits x M x ldc floats) and are summed in fixed order.
static int gemm_s_splitk(int M, int N, int K, const float* A, long long sam, long long sak, const float* B, long long sbn,
                         long long sbk, float* C, int ldc, float* scratch, size_t scratch_floats, cudaStream_t st) {
    if (M <= 0 || N <= 0) return SRNN_OK;
    const int tiles = cdiv(N, SBN) * cdiv(M, SBM);
    int splits = cdiv(592, tiles);                                    // ~4 CTAs per SM
    const size_t cap = scratch ? scratch_floats / ((size_t)M * ldc) : 0;
    if ((size_t)splits > cap) splits = (int)cap;
    if (splits > K / (4 * SBK)) splits = K / (4 * SBK);
    if (splits < 2) return gemm_s(M, N, K, A, sam, sak, B, sbn, sbk, nullptr, 0, C, ldc, st);
    int kper = cdiv(cdiv(K, splits), SBK) * SBK;
    splits = cdiv(K, kper);
    dim3 grid(cdiv(N, SBN), cdiv(M, SBM), splits);
    SRNN_LAUNCH(k_gemm_f32_strided, grid, 256, 0, st, M, N, K, A, sam, sak, B, sbn, sbk, nullptr, 0, scratch, ldc, kper);
    return sum_splits(scratch, splits, (size_t)M * ldc, (size_t)M * ldc, C, st);
}
// dIn (rows, K) = dOut (rows, N) . W (N, K)  [+ add]
static int gemm_dx(int rows, int Kdim, int Ndim, const float* dOut, int ld_do, const float* W, int ld_w, const float* add,
                   int ldadd, float* dIn, int ld_di, cudaStream_t st) {
    return gemm_s(rows, Kdim, Ndim, dOut, ld_do, 1, W, 1, ld_w, add, ldadd, dIn, ld_di, st);
}
// dW (N, K) = dOut (rows, N)^T . In (rows, K)
static int gemm_dw(int Ndim, int Kdim, int rows, const float* dOut, int ld_do, const float* In, int ld_in, float* dW,
                   int ld_dw, cudaStream_t st) {
    return gemm_s(Ndim, Kdim, rows, dOut, 1, ld_do, In, 1, ld_in, nullptr, 0, dW, ld_dw, st);
}

static inline int gsz(size_t n) { return (int)((n + 255) / 256 > 8192 ? 8192 : (n + 255) / 256); }

// ------------------------------------------------------------------------------------------------
// element-wise / reduction kernels
// ------------------------------------------------------------------------------------------------
// log_softmax backward, one warp per 256-wide row: dlogits = dlogp - exp(logp) * sum(dlogp)
__global__ void k_logsoftmax_bwd(const float* __restrict__ dlogp, const float* __restrict__ logp, float* __restrict__ dlogits,
                                 int rows) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const size_t o = (size_t)row * SRNN_Q + lane * 8;
    float g[8], s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = dlogp[o + i]; s += g[i]; }
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) dlogits[o + i] = g[i] - expf(logp[o + i]) * s;
}

// Fused backward of sequence_nll_loss_bits (nn.py:66-70) through log_softmax: with L = -mean_r logp[r, target[r]] * log2(e)
// and an upstream scalar gradient *gscale (1 if null), dL/dlogits[r, :] = (exp(logp[r, :]) - onehot(target[r])) * g * log2(e) / rows.
// One warp per row; writes the fp32 gradient (bias column sums) and, when d16 is given, the bf16 copy the tensor-core
// GEMMs read -- no dense dL/dlogp is ever materialised.
__global__ void k_nll_logsoftmax_bwd(const float* __restrict__ logp, const int64_t* __restrict__ target,
                                     const float* __restrict__ gscale, float* __restrict__ dlogits,
                                     __nv_bfloat16* __restrict__ d16, int rows) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float s = (gscale ? __ldg(gscale) : 1.f) * 1.4426950408889634f / (float)rows;
    long long tq = target[row];
    const int tgt = tq < 0 ? 0 : (tq > SRNN_Q - 1 ? SRNN_Q - 1 : (int)tq);
    const size_t o = (size_t)row * SRNN_Q + lane * 8;
    const float4 a = *reinterpret_cast<const float4*>(logp + o), b = *reinterpret_cast<const float4*>(logp + o + 4);
    float g[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = (expf(g[i]) - (lane * 8 + i == tgt ? 1.f : 0.f)) * s;
    *reinterpret_cast<float4*>(dlogits + o) = make_float4(g[0], g[1], g[2], g[3]);
    *reinterpret_cast<float4*>(dlogits + o + 4) = make_float4(g[4], g[5], g[6], g[7]);
    if (d16) {
        __nv_bfloat162 h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(g[2 * i], g[2 * i + 1]);
        *reinterpret_cast<uint4*>(d16 + o) = *reinterpret_cast<const uint4*>(h);
    }
}

__global__ void k_relu_mask(const float* __restrict__ dx, const float* __restrict__ x, float* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = x[i] > 0.f ? dx[i] : 0.f;
}

// column sums of X (rows, cols, ld): stage 1 -> partial[chunk][col], stage 2 -> out[col] (fixed order)
constexpr int CS_CHUNKS = 64;
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T>
__global__ void k_colsum_partial(const T* __restrict__ X, int rows, int cols, int ld, float* __restrict__ partial, int per) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;      // four independent chains: the loads of a row block overlap
    int r = r0;
    for (; r + 3 < r1; r += 4) {
        s0 += to_f(X[(size_t)r * ld + col]);
        s1 += to_f(X[(size_t)(r + 1) * ld + col]);
        s2 += to_f(X[(size_t)(r + 2) * ld + col]);
        s3 += to_f(X[(size_t)(r + 3) * ld + col]);
    }
    for (; r < r1; ++r) s0 += to_f(X[(size_t)r * ld + col]);
    partial[(size_t)blockIdx.y * cols + col] = (s0 + s1) + (s2 + s3);
}
__global__ void k_colsum_final(const float* __restrict__ partial, int cols, int chunks, float* __restrict__ out) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += partial[(size_t)c * cols + col];
    out[col] = s;
}
// `partial` holds CS_CHUNKS * max_cols floats (max_cols = the widest matrix of the pass): narrow matrices use the spare
// room for more row chunks so that the grid still fills the GPU (fixed summation order either way).
struct ColsumScratch {
    float* partial;
    size_t cap;                       // floats available
    operator float*() const { return partial; }
};
template <typename T>
static int colsum(const T* X, int rows, int cols, int ld, const ColsumScratch& cs, float* out, cudaStream_t st) {
    float* partial = cs.partial;
    const size_t g_colsum_cap = cs.cap;
    int chunks = CS_CHUNKS;
    const int want = cdiv(1184, cdiv(cols, 128));                         // ~8 CTAs per SM
    if (want > chunks) chunks = want;
    if (g_colsum_cap && (size_t)chunks * cols > g_colsum_cap) chunks = (int)(g_colsum_cap / cols);
    if (!g_colsum_cap) chunks = CS_CHUNKS;
    if (chunks > rows) chunks = rows > 0 ? rows : 1;
    const int per = cdiv(rows, chunks);
    chunks = cdiv(rows, per);
    SRNN_LAUNCH((k_colsum_partial<T>), dim3(cdiv(cols, 128), chunks), 128, 0, st, X, rows, cols, ld, partial, per);
    SRNN_LAUNCH(k_colsum_final, cdiv(cols, 128), 128, 0, st, partial, cols, chunks, out);
    return SRNN_OK;
}

// out = a + b (+ c)
__global__ void k_add3(const float* a, const float* b, const float* c, float* out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = a[i] + b[i] + (c ? c[i] : 0.f);
}

// ------------------------------------------------------------------------------------------------
// dTbl[j][q][h] = sum over rows r=(b,t) with seq[b, off + t + j] == q of dpre1[r][h]
// The bucket of a window position (b, pi) does not depend on the tap: position pi with value q feeds tap j with row
// t = pi - j.  So the window positions are counting-sorted by value ONCE (ascending position order, deterministic), and a
// block per (value, feature chunk, segment) walks its positions, reading for each the FS consecutive rows
// dpre1[b, pi-FS+1 .. pi] into FS register accumulators: no atomics, no shared-memory table, coalesced row reads.
// ------------------------------------------------------------------------------------------------
constexpr int DT_SEG = 4;      // segments per value bucket (partials summed in fixed order)
constexpr int DT_MAXFS = 32;

// Counting sort of the window positions by sample value, ascending position inside every bucket (deterministic):
// each WARP owns a chunk of QS_CHUNK consecutive positions.  Pass 1: per-chunk histograms; pass 2: scans (per value over the
// chunks, then over the values); pass 3: every warp re-walks its chunk and writes each position at its final rank.
constexpr int QS_CHUNK = 512;
__device__ __forceinline__ int qs_value(const uint8_t* __restrict__ seq, int seq_ld, int off, int W, int i) {
    return seq[(size_t)(i / W) * seq_ld + off + (i % W)];
}
__global__ void __launch_bounds__(128)
k_qs_hist(const uint8_t* __restrict__ seq, int seq_ld, int off, int N, int W, int nchunk, int* __restrict__ hist /* (Q, nchunk) */) {
    __shared__ int cnt[4][SRNN_Q];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 4 + w;
    for (int q = lane; q < SRNN_Q; q += 32) cnt[w][q] = 0;
    __syncwarp();
    if (c < nchunk) {
        const int i0 = c * QS_CHUNK;
        for (int r = 0; r < QS_CHUNK; r += 32) {
            const int i = i0 + r + lane;
            const int v = i < N ? qs_value(seq, seq_ld, off, W, i) : -1;
            const unsigned m = __match_any_sync(0xffffffffu, v);
            if (v >= 0 && lane == __ffs(m) - 1) cnt[w][v] += __popc(m);
            __syncwarp();
        }
        for (int q = lane; q < SRNN_Q; q += 32) hist[(size_t)q * nchunk + c] = cnt[w][q];
    }
}
// block q: exclusive scan of hist[q][:] in place, total -> counts[q]
__global__ void k_qs_scan_chunks(int* __restrict__ hist, int nchunk, int* __restrict__ counts) {
    const int q = blockIdx.x;
    __shared__ int part[256];
    int* row = hist + (size_t)q * nchunk;
    const int per = (nchunk + 255) / 256;
    const int c0 = threadIdx.x * per, c1 = min(nchunk, c0 + per);
    int s = 0;
    for (int c = c0; c < c1; ++c) s += row[c];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = 0; t < 256; ++t) { const int v = part[t]; part[t] = run; run += v; }
        counts[q] = run;
    }
    __syncthreads();
    int run = part[threadIdx.x];
    for (int c = c0; c < c1; ++c) { const int v = row[c]; row[c] = run; run += v; }
}
__global__ void k_q_prefix(const int* __restrict__ counts, int* __restrict__ starts) {
    if (threadIdx.x == 0) {
        int s = 0;
        for (int q = 0; q < SRNN_Q; ++q) { starts[q] = s; s += counts[q]; }
        starts[SRNN_Q] = s;
    }
}
__global__ void __launch_bounds__(128)
k_qs_scatter(const uint8_t* __restrict__ seq, int seq_ld, int off, int N, int W, int nchunk, const int* __restrict__ hist,
             const int* __restrict__ starts, int* __restrict__ pos) {
    __shared__ int run[4][SRNN_Q];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 4 + w;
    if (c >= nchunk) return;
    for (int q = lane; q < SRNN_Q; q += 32) run[w][q] = starts[q] + hist[(size_t)q * nchunk + c];
    __syncwarp();
    const int i0 = c * QS_CHUNK;
    for (int r = 0; r < QS_CHUNK; r += 32) {
        const int i = i0 + r + lane;
        const int v = i < N ? qs_value(seq, seq_ld, off, W, i) : -1;
        const unsigned m = __match_any_sync(0xffffffffu, v);
        if (v >= 0) pos[run[w][v] + __popc(m & ((1u << lane) - 1))] = i;
        __syncwarp();
        if (v >= 0 && lane == __ffs(m) - 1) run[w][v] += __popc(m);
        __syncwarp();
    }
}
// grid (Q, ceil(H/1024), DT_SEG), 256 threads x 4 consecutive features (one 8-byte bf16 / 16-byte fp32 load per row)
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T1, int FSMAX>
__global__ void __launch_bounds__(256, 2)
k_dtbl_accum(const int* __restrict__ pos, const int* __restrict__ starts, const T1* __restrict__ dpre1, int T, int W, int H,
             int FS, int nb, float* __restrict__ partial /* (DT_SEG, FS, Q, H) */) {
    const int q = blockIdx.x, h = blockIdx.y * 1024 + threadIdx.x * 4, seg = blockIdx.z;
    const int s0 = starts[q], n = starts[q + 1] - s0;
    if (h >= H) return;
    // Segment = a RANGE of window positions (not a share of the bucket): blocks are dispatched z-major, so the blocks in
    // flight at any time work on the same quarter of dpre1 (68 MB at C3), which then stays in L2 across the FS re-reads.
    const long long ntot = (long long)W * nb;
    const int lo_pos = (int)(ntot * seg / DT_SEG), hi_pos = (int)(ntot * (seg + 1) / DT_SEG);
    int k0, k1;
    {
        int a = s0, b2 = s0 + n;                              // first index with pos >= lo_pos (positions ascend in a bucket)
        while (a < b2) { const int m = (a + b2) >> 1; if (pos[m] < lo_pos) a = m + 1; else b2 = m; }
        k0 = a;
        b2 = s0 + n;
        while (a < b2) { const int m = (a + b2) >> 1; if (pos[m] < hi_pos) a = m + 1; else b2 = m; }
        k1 = a;
    }
    float acc[FSMAX][4];
#pragma unroll
    for (int j = 0; j < FSMAX; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    int i_next = k0 < k1 ? pos[k0] : 0;
    for (int k = k0; k < k1; ++k) {
        const int i = i_next, b = i / W, pi = i % W;
        if (k + 1 < k1) i_next = pos[k + 1];                               // next position's index while this one's rows load
        const T1* rowbase = dpre1 + ((size_t)b * T + pi) * H + h;          // row t = pi - j  ->  rowbase - j*H
        if constexpr (sizeof(T1) == 2 && FSMAX % 5 == 0) {
            if (FS == FSMAX && pi >= FS - 1 && pi < T) {                    // interior position: every tap's row exists
#pragma unroll
                for (int j0 = 0; j0 < FSMAX; j0 += 5) {                     // five unconditional 8-byte loads in flight, then the adds
                    uint2 raw[5];
#pragma unroll
                    for (int u = 0; u < 5; ++u) raw[u] = *reinterpret_cast<const uint2*>(rowbase - (long long)(j0 + u) * H);
#pragma unroll
                    for (int u = 0; u < 5; ++u) {
                        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[u].x));
                        const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[u].y));
                        acc[j0 + u][0] += a.x; acc[j0 + u][1] += a.y; acc[j0 + u][2] += c.x; acc[j0 + u][3] += c.y;
                    }
                }
                continue;
            }
        }
#pragma unroll
        for (int j = 0; j < FSMAX; ++j) {
            const int t = pi - j;
            if (j < FS && t >= 0 && t < T) {
                float v[4];
                ld4(rowbase - (long long)j * H, v);
                acc[j][0] += v[0]; acc[j][1] += v[1]; acc[j][2] += v[2]; acc[j][3] += v[3];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < FSMAX; ++j)
        if (j < FS)
            *reinterpret_cast<float4*>(partial + (((size_t)seg * FS + j) * SRNN_Q + q) * H + h) =
                make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
}
// sum the segment partials (fixed order) -> dTbl (FS,Q,H) and its per-tap transpose dTblT (FS,H,Q)
__global__ void k_dtbl_final(const float* __restrict__ partial, int FS, int H, float* __restrict__ out, float* __restrict__ outT) {
    const size_t n = (size_t)FS * SRNN_Q * H;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < DT_SEG; ++c) s += partial[(size_t)c * n + i];
        out[i] = s;
        const int h = (int)(i % H), q = (int)((i / H) % SRNN_Q), j = (int)(i / ((size_t)H * SRNN_Q));
        outT[((size_t)j * H + h) * SRNN_Q + q] = s;
    }
}

template <typename T1>
static int dtbl_compute(const uint8_t* seq, int seq_ld, int off, const T1* dpre1, int B, int T, int H, int FS, int* iwork,
                        float* partial, float* dTbl, float* dTblT, cudaStream_t st) {
    if (FS > DT_MAXFS) return fail(SRNN_ERR_UNSUPPORTED, "frame size %d > %d", FS, DT_MAXFS);
    const int W = T + FS - 1, N = B * W, nchunk = cdiv(N, QS_CHUNK);
    if (H % 4) return fail(SRNN_ERR_UNSUPPORTED, "dim %d must be a multiple of 4", H);
    int* counts = iwork;
    int* starts = iwork + 256;
    int* pos = iwork + 768;
    int* hist = iwork + 1024 + (size_t)B * (T + FS);           // (Q, nchunk)
    SRNN_LAUNCH(k_qs_hist, cdiv(nchunk, 4), 128, 0, st, seq, seq_ld, off, N, W, nchunk, hist);
    SRNN_LAUNCH(k_qs_scan_chunks, SRNN_Q, 256, 0, st, hist, nchunk, counts);
    SRNN_LAUNCH(k_q_prefix, 1, 32, 0, st, counts, starts);
    SRNN_LAUNCH(k_qs_scatter, cdiv(nchunk, 4), 128, 0, st, seq, seq_ld, off, N, W, nchunk, hist, starts, pos);
    const dim3 grid(SRNN_Q, cdiv(H, 1024), DT_SEG);          // FSMAX = accumulator rows held in registers
    if (FS <= 4) SRNN_LAUNCH((k_dtbl_accum<T1, 4>), grid, 256, 0, st, pos, starts, dpre1, T, W, H, FS, B, partial);
    else if (FS <= 8) SRNN_LAUNCH((k_dtbl_accum<T1, 8>), grid, 256, 0, st, pos, starts, dpre1, T, W, H, FS, B, partial);
    else if (FS <= 16) SRNN_LAUNCH((k_dtbl_accum<T1, 16>), grid, 256, 0, st, pos, starts, dpre1, T, W, H, FS, B, partial);
    else if (FS <= 20) SRNN_LAUNCH((k_dtbl_accum<T1, 20>), grid, 256, 0, st, pos, starts, dpre1, T, W, H, FS, B, partial);
    else SRNN_LAUNCH((k_dtbl_accum<T1, DT_MAXFS>), grid, 256, 0, st, pos, starts, dpre1, T, W, H, FS, B, partial);
    SRNN_LAUNCH(k_dtbl_final, gsz((size_t)FS * SRNN_Q * H), 256, 0, st, partial, FS, H, dTbl, dTblT);
    return SRNN_OK;
}

// fold dTbl back onto the MLP input weights (H,Q,FS) and the embedding (Q,Q):  Tbl[j][q][h] = sum_e Wm[h,e,j] E[q,e]
static int tbl_foldback(srnn_ctx* ctx, const srnn_params* P, const srnn_params* G, const float* dTblT, float* scratch2HQF,
                        float* dWmt, float* dWm, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// GRU cell backward for one frame (all utterances): recomputes r, z, n from the saved projections
// ------------------------------------------------------------------------------------------------
__global__ void k_gru_bwd_gates(const float* __restrict__ gi, const float* __restrict__ gh, int g_ld,
                                const float* __restrict__ h_prev, int hp_ld, const float* __restrict__ dy, int dy_ld,
                                const float* __restrict__ dh_carry, float* __restrict__ dgi, float* __restrict__ dgh,
                                float* __restrict__ dh_prev, int H, __nv_bfloat16* __restrict__ dgi16 = nullptr,
                                __nv_bfloat16* __restrict__ dgh16 = nullptr) {
    const int b = blockIdx.y;
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= H) return;
    const float* gir = gi + (size_t)b * g_ld;
    const float* ghr = gh + (size_t)b * g_ld;
    const float r = 1.f / (1.f + expf(-(gir[u] + ghr[u])));
    const float z = 1.f / (1.f + expf(-(gir[H + u] + ghr[H + u])));
    const float ghn = ghr[2 * H + u];
    const float n = tanhf(gir[2 * H + u] + r * ghn);
    const float hp = h_prev[(size_t)b * hp_ld + u];
    float dh = dy[(size_t)b * dy_ld + u];
    if (dh_carry) dh += dh_carry[(size_t)b * H + u];
    const float dn = dh * (1.f - z);
    const float dz = dh * (hp - n);
    const float dpn = dn * (1.f - n * n);
    const float dr = dpn * ghn;
    const float dpr = dr * r * (1.f - r);
    const float dpz = dz * z * (1.f - z);
    float* dgir = dgi + (size_t)b * g_ld;
    float* dghr = dgh + (size_t)b * g_ld;
    dgir[u] = dpr;          dghr[u] = dpr;
    dgir[H + u] = dpz;      dghr[H + u] = dpz;
    dgir[2 * H + u] = dpn;  dghr[2 * H + u] = dpn * r;
    if (dgi16) {
        __nv_bfloat16* a = dgi16 + (size_t)b * g_ld;
        __nv_bfloat16* c = dgh16 + (size_t)b * g_ld;
        a[u] = __float2bfloat16(dpr);          c[u] = __float2bfloat16(dpr);
        a[H + u] = __float2bfloat16(dpz);      c[H + u] = __float2bfloat16(dpz);
        a[2 * H + u] = __float2bfloat16(dpn);  c[2 * H + u] = __float2bfloat16(dpn * r);
    }
    dh_prev[(size_t)b * H + u] = dh * z;
}

// Hprev[(b,f), :] = f ? Y[(b,f-1), :] : h0[b, :]   (the recurrent input of every frame, for dW_hh)
__global__ void k_build_hprev(const float* __restrict__ Y, const float* __restrict__ h0, float* __restrict__ hp, int F, int H) {
    const int r = blockIdx.x, b = r / F, f = r % F;
    const float* src = f ? Y + (size_t)(r - 1) * H : h0 + (size_t)b * H;
    for (int u = threadIdx.x; u < H; u += blockDim.x) hp[(size_t)r * H + u] = src[u];
}

// ------------------------------------------------------------------------------------------------
// weight-norm backward per row: w = g v/||v||  =>  dg = <dw, v>/||v||,  dv = g/||v|| (dw - v <dw,v>/||v||^2)
// (plain weight: dweight = dw)
// ------------------------------------------------------------------------------------------------
__global__ void k_wn_bwd(const float* __restrict__ dw, const float* __restrict__ g, const float* __restrict__ v,
                         float* __restrict__ dweight, float* __restrict__ dg, float* __restrict__ dv, int cols) {
    const int r = blockIdx.x;
    const float* dwr = dw + (size_t)r * cols;
    if (dweight) {
        for (int c = threadIdx.x; c < cols; c += blockDim.x) dweight[(size_t)r * cols + c] = dwr[c];
        return;
    }
    __shared__ float red0[32], red1[32];
    __shared__ float s_vv, s_dv;
    const float* vr = v + (size_t)r * cols;
    float vv = 0.f, dvv = 0.f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        vv = fmaf(vr[c], vr[c], vv);
        dvv = fmaf(dwr[c], vr[c], dvv);
    }
    for (int o = 16; o; o >>= 1) {
        vv += __shfl_xor_sync(0xffffffffu, vv, o);
        dvv += __shfl_xor_sync(0xffffffffu, dvv, o);
    }
    if ((threadIdx.x & 31) == 0) { red0[threadIdx.x >> 5] = vv; red1[threadIdx.x >> 5] = dvv; }
    __syncthreads();
    if (threadIdx.x < 32) {
        float a = threadIdx.x < (blockDim.x >> 5) ? red0[threadIdx.x] : 0.f;
        float b = threadIdx.x < (blockDim.x >> 5) ? red1[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (threadIdx.x == 0) { s_vv = a; s_dv = b; }
    }
    __syncthreads();
    const float nrm = sqrtf(s_vv), gg = g[r];
    if (threadIdx.x == 0 && dg) dg[r] = s_dv / nrm;
    if (dv) {
        const float sc = gg / nrm, proj = s_dv / s_vv;
        for (int c = threadIdx.x; c < cols; c += blockDim.x) dv[(size_t)r * cols + c] = sc * (dwr[c] - vr[c] * proj);
    }
}
static int wn_bwd(const float* dw, const srnn_conv_params& p, const srnn_conv_params& g, int rows, int cols, cudaStream_t st) {
    float* dweight = (float*)g.weight;
    float* dg = (float*)g.weight_g;
    float* dv = (float*)g.weight_v;
    if (p.weight) {
        if (!dweight) return SRNN_OK;
        SRNN_LAUNCH(k_wn_bwd, rows, 256, 0, st, dw, nullptr, nullptr, dweight, nullptr, nullptr, cols);
    } else {
        if (!dg && !dv) return SRNN_OK;
        SRNN_LAUNCH(k_wn_bwd, rows, 256, 0, st, dw, p.weight_g, p.weight_v, nullptr, dg, dv, cols);
    }
    return SRNN_OK;
}

// top tier: gradient of the K-concatenated matrix (H, n + cond_dim + spk_dim) -> its three sources
__global__ void k_unpack_top_in(const float* __restrict__ dcomb, float* __restrict__ d_in, float* __restrict__ d_c,
                                float* __restrict__ d_s /* (H, spk_dim) wrt folded spk_expand */, const float* __restrict__ emb,
                                int n, int cond_dim, int spk_dim) {
    const int h = blockIdx.x, kin = n + cond_dim + spk_dim;
    const float* r = dcomb + (size_t)h * kin;
    for (int i = threadIdx.x; i < n; i += blockDim.x) d_in[(size_t)h * n + i] = r[i];
    for (int i = threadIdx.x; i < cond_dim; i += blockDim.x) d_c[(size_t)h * cond_dim + i] = r[n + i];
    for (int e = threadIdx.x; e < spk_dim; e += blockDim.x) {      // comb[h,s] = sum_e w_s[h,e] E[s,e]
        float a = 0.f;
        for (int s = 0; s < spk_dim; ++s) a = fmaf(r[n + cond_dim + s], emb[s * spk_dim + e], a);
        d_s[(size_t)h * spk_dim + e] = a;
    }
}
// dEmb[s,e] = sum_h dcomb[h, n+cond_dim+s] * w_s[h,e]
__global__ void k_spk_emb_grad(const float* __restrict__ dcomb, const float* __restrict__ w_s, float* __restrict__ demb,
                               int H, int kin, int off, int spk_dim) {
    const int s = blockIdx.x / spk_dim, e = blockIdx.x % spk_dim;
    __shared__ float red[128];
    float a = 0.f;
    for (int h = threadIdx.x; h < H; h += blockDim.x) a = fmaf(dcomb[(size_t)h * kin + off + s], w_s[(size_t)h * spk_dim + e], a);
    red[threadIdx.x] = a;
    __syncthreads();
    for (int o = 64; o; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) demb[s * spk_dim + e] = red[0];
}

// (FS, H, Q) -> (H, Q, FS)
__global__ void k_untranspose_mlp_in(const float* __restrict__ wt, float* __restrict__ w, int H, int Q, int FS) {
    const int h = blockIdx.x, j = blockIdx.y;
    for (int e = threadIdx.x; e < Q; e += blockDim.x) w[((size_t)h * Q + e) * FS + j] = wt[((size_t)j * H + h) * Q + e];
}


static int tbl_foldback(srnn_ctx* ctx, const srnn_params* P, const srnn_params* G, const float* dTblT, float* scratch2HQF,
                        float* dWmt, float* dWm, cudaStream_t st) {
    const int H = ctx->H, Q = ctx->Q, FS0 = ctx->FS0;
    float* wm_fold = scratch2HQF;                       // (H, Q, FS)
    float* wm_t = scratch2HQF + (size_t)H * Q * FS0;    // (FS, H, Q)
    SRNN_TRY(wn_fold(P->mlp_input, wm_fold, H, Q * FS0, st));
    SRNN_TRY(transpose_mlp_in(wm_fold, wm_t, H, Q, FS0, st));
    // dWm_t[(j,h), e] = sum_q dTblT[(j,h), q] E[q, e]        (one GEMM over all taps)
    SRNN_TRY(gemm_s(FS0 * H, Q, Q, dTblT, Q, 1, P->embedding, 1, Q, nullptr, 0, dWmt, Q, st));
    // dE[q, e] = sum_{(j,h)} dTblT[(j,h), q] wm_t[(j,h), e]
    float* dE = (float*)G->embedding;
    if (dE)
        SRNN_TRY(gemm_s_splitk(Q, Q, FS0 * H, dTblT, 1, Q, wm_t, 1, Q, dE, Q, scratch2HQF + 2 * (size_t)H * Q * FS0,
                               2 * (size_t)H * Q * FS0, st));
    SRNN_LAUNCH(k_untranspose_mlp_in, dim3(H, FS0), 256, 0, st, dWmt, dWm, H, Q, FS0);
    SRNN_TRY(wn_bwd(dWm, P->mlp_input, G->mlp_input, H, Q * FS0, st));
    return SRNN_OK;
}

// BPTT of one GRU layer, frame by frame in fp32 (the SRNN_MODE_FP32 schedule; also the srnn_gru_seq_bwd test hook).
// scratch: 2 * B * H floats.
int gru_seq_bwd_f32(int B, int Fr, int H, const float* GI, const float* GH, const float* Y, const float* h0, const float* dY,
                    const float* w_hh, float* dGI, float* dGH, float* dh0, float* scratch, cudaStream_t st) {
    float* dhc0 = scratch;
    float* dhc1 = scratch + (size_t)B * H;
    float* carry = nullptr;
    float* cnext = dhc0;
    for (int f = Fr - 1; f >= 0; --f) {
        const float* hp = f ? Y + (size_t)(f - 1) * H : h0;
        const int hp_ld = f ? Fr * H : H;
        float* part = (cnext == dhc0) ? dhc1 : dhc0;
        SRNN_LAUNCH(k_gru_bwd_gates, dim3(cdiv(H, 128), B), 128, 0, st, GI + (size_t)f * 3 * H, GH + (size_t)f * 3 * H,
                    Fr * 3 * H, hp, hp_ld, dY + (size_t)f * H, Fr * H, carry, dGI + (size_t)f * 3 * H,
                    dGH + (size_t)f * 3 * H, part, H);
        SRNN_TRY(gemm_s(B, H, 3 * H, dGH + (size_t)f * 3 * H, (long long)Fr * 3 * H, 1, w_hh, 1, H, part, H, cnext, H, st));
        carry = cnext;
        cnext = part;
    }
    if (dh0) SRNN_TRY(copy_f32(carry, dh0, (size_t)B * H, st));
    return SRNN_OK;
}

// ------------------------------------------------------------------------------------------------
// orchestration
// ------------------------------------------------------------------------------------------------
struct Bump2 {
    char* base;
    size_t off;
    Bump2(void* b, size_t o) : base((char*)b), off(o) {}
    template <typename T>
    T* take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? (T*)(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

size_t backward_scratch_bytes(const srnn_ctx* ctx, int B, int T) {
    const srnn_config& c = ctx->cfg;
    const size_t H = ctx->H, Q = ctx->Q, R = (size_t)B * T, FS0 = ctx->FS0;
    size_t maxM = 0, maxfs = 1, maxkin = 1;
    for (int i = 0; i < c.n_tiers; ++i) {
        const size_t M = (size_t)B * (T / ctx->tiers[i].n);
        if (M > maxM) maxM = M;
        if ((size_t)ctx->tiers[i].fs > maxfs) maxfs = ctx->tiers[i].fs;
        if ((size_t)ctx->tiers[i].kin > maxkin) maxkin = ctx->tiers[i].kin;
    }
    size_t f = R * Q + 2 * R * H                                // dlogits, dA, dB
               + 3 * maxM * H + 2 * maxM * 3 * H + 2 * (size_t)B * H   // dY ping/pong, dX, dGI, dGH, carries
               + maxfs * H * H * 2 + maxfs * H + H * maxkin + H * (size_t)c.spk_dim + 3 * H * (H > maxkin ? H : maxkin)   // weight-grad staging
               + (DT_SEG + 2) * FS0 * Q * H + 2 * FS0 * H * Q + (1024 + (size_t)B * (T + FS0) + 256 * ((size_t)B * (T + FS0) / 512 + 2))             // dTbl partials + final, dWm_t, dWm
               + (size_t)CS_CHUNKS * (maxfs * H > 3 * H ? maxfs * H : 3 * H) + 3 * H + Q * H + Q + 4096;
    return f * sizeof(float) + 64 * 256;
}

int predict_bwd_f32(srnn_ctx* ctx, const float* logp, const float* dlogp, const srnn_params* P, const srnn_params* G,
                    cudaStream_t st, const int64_t* nll_target, const float* nll_gscale) {
    const FwdPlan& F = ctx->fwd;
    const srnn_config& c = ctx->cfg;
    const int H = ctx->H, Q = ctx->Q, NT = c.n_tiers, NL = c.n_rnn, lookback = ctx->lookback, FS0 = ctx->FS0;
    const int B = F.B, T = F.T, R = B * T, Lseq = lookback + T - 1;
    if (F.bytes + backward_scratch_bytes(ctx, B, T) > ctx->ws_bytes)
        return fail(SRNN_ERR_STATE, "backward scratch was not reserved by the forward pass");
    int maxfs = 1, maxkin = 1;
    size_t maxM = 0;
    for (int i = 0; i < NT; ++i) {
        const size_t M = (size_t)B * (T / ctx->tiers[i].n);
        if (M > maxM) maxM = M;
        if (ctx->tiers[i].fs > maxfs) maxfs = ctx->tiers[i].fs;
        if (ctx->tiers[i].kin > maxkin) maxkin = ctx->tiers[i].kin;
    }
    Bump2 b(ctx->ws, F.bytes);
    float* dlogits = b.take<float>((size_t)R * Q);
    float* dA = b.take<float>((size_t)R * H);
    float* dB = b.take<float>((size_t)R * H);
    float* dYa = b.take<float>(maxM * H);
    float* dYb = b.take<float>(maxM * H);
    float* dXbuf[2] = {b.take<float>(maxM * H), nullptr};
    float* dGI = b.take<float>(maxM * 3 * H);
    float* dGH = b.take<float>(maxM * 3 * H);
    float* dhc0 = b.take<float>((size_t)B * H);
    float* dhc1 = b.take<float>((size_t)B * H);
    float* dWup = b.take<float>((size_t)maxfs * H * H);
    float* dwf = b.take<float>((size_t)maxfs * H * H);
    float* dbup = b.take<float>((size_t)maxfs * H);
    float* dWin = b.take<float>((size_t)H * maxkin);
    float* wsf = b.take<float>((size_t)H * c.spk_dim);
    const size_t stg = (size_t)H * (H > maxkin ? H : maxkin);
    float* t_in = b.take<float>(stg);
    float* t_c = b.take<float>(stg);
    float* t_s = b.take<float>(stg);
    float* dTblP = b.take<float>((size_t)DT_SEG * FS0 * Q * H);
    float* dTbl = b.take<float>((size_t)FS0 * Q * H);
    float* dTblT = b.take<float>((size_t)FS0 * Q * H);
    int* iwork = b.take<int>(1024 + (size_t)B * (T + FS0) + 256 * ((size_t)B * (T + FS0) / 512 + 2));
    float* dWmt = b.take<float>((size_t)FS0 * H * Q);
    float* dWm = b.take<float>((size_t)FS0 * H * Q);
    const size_t csp_floats = (size_t)CS_CHUNKS * (maxfs * H > 3 * H ? maxfs * H : 3 * H);
    const ColsumScratch csp{b.take<float>(csp_floats), csp_floats};
    float* dbtmp = b.take<float>((size_t)3 * H);
    float* dWo = b.take<float>((size_t)Q * H);
    float* wmf = dWm;   // folded mlp-input weights (H,Q,FS0) are rebuilt into dWm's storage before it is needed (see below)
    (void)wmf;

    // ---- log_softmax backward ----
    if (nll_target) SRNN_LAUNCH(k_nll_logsoftmax_bwd, cdiv(R, 8), 256, 0, st, logp, nll_target, nll_gscale, dlogits, (__nv_bfloat16*)nullptr, R);
    else SRNN_LAUNCH(k_logsoftmax_bwd, cdiv(R, 8), 256, 0, st, dlogp, logp, dlogits, R);
    // ---- output layer: logits = x2 W_o^T + b_o ----
    SRNN_TRY(gemm_dw(Q, H, R, dlogits, Q, F.X2, H, dWo, H, st));
    SRNN_TRY(wn_bwd(dWo, P->mlp_output, G->mlp_output, Q, H, st));
    if (G->mlp_output.bias) SRNN_TRY(colsum(dlogits, R, Q, Q, csp, (float*)G->mlp_output.bias, st));
    SRNN_TRY(gemm_dx(R, H, Q, dlogits, Q, ctx->w_out, H, nullptr, 0, dA, H, st));               // dx2
    SRNN_LAUNCH(k_relu_mask, gsz((size_t)R * H), 256, 0, st, dA, F.X2, dA, (size_t)R * H);        // dpre2
    // ---- hidden layer ----
    SRNN_TRY(gemm_dw(H, H, R, dA, H, F.X1, H, t_in, H, st));
    SRNN_TRY(wn_bwd(t_in, P->mlp_hidden, G->mlp_hidden, H, H, st));
    if (G->mlp_hidden.bias) SRNN_TRY(colsum(dA, R, H, H, csp, (float*)G->mlp_hidden.bias, st));
    SRNN_TRY(gemm_dx(R, H, H, dA, H, ctx->w_hid, H, nullptr, 0, dB, H, st));                    // dx1
    SRNN_LAUNCH(k_relu_mask, gsz((size_t)R * H), 256, 0, st, dB, F.X1, dB, (size_t)R * H);        // dpre1 = dc0
    // ---- folded table: dTbl, then back onto W_in (H,Q,FS) and E (Q,Q) ----
    SRNN_TRY(dtbl_compute(F.seq, Lseq, lookback - FS0, dB, B, T, H, FS0, iwork, dTblP, dTbl, dTblT, st));
    SRNN_TRY(tbl_foldback(ctx, P, G, dTblT, dTblP, dWmt, dWm, st));
    SRNN_CUDA(cudaEventRecord(ctx->ev_stage[0], st));                           // MLP + embedding gradients are final
    // ---- frame tiers, lowest first: each receives dUP (M, fs*H) from below ----
    const float* dUP = dB;                            // tier 0's upsampled output is the MLP conditioning c0
    int xb = 0;
    for (int i = 0; i < NT; ++i) {
        const TierPacked& t = ctx->tiers[i];
        const srnn_tier_params& tp = P->tiers[i];
        const srnn_tier_params& tg = G->tiers[i];
        const int Fr = T / t.n, M = B * Fr;
        // upsampling: UP = Y_last W_up^T + b_up
        const float* Ylast = F.Y[i][NL - 1];
        SRNN_TRY(gemm_dw(t.fs * H, H, M, dUP, t.fs * H, Ylast, H, dWup, H, st));
        SRNN_TRY(colsum(dUP, M, t.fs * H, t.fs * H, csp, dbup, st));
        SRNN_TRY(unpack_up_grad(dWup, dbup, dwf, (float*)tg.upsampling.bias, H, t.fs, st));
        SRNN_TRY(wn_bwd(dwf, tp.upsampling, tg.upsampling, H, H * t.fs, st));
        SRNN_CUDA(cudaEventRecord(ctx->ev_stage[1 + 2 * i], st));                 // tier i's upsampling gradients are final
        float* dY = dYa;
        float* dYn = dYb;
        SRNN_TRY(gemm_dx(M, H, t.fs * H, dUP, t.fs * H, t.w_up, H, nullptr, 0, dY, H, st));
        // GRU layers, last first (BPTT inside each)
        for (int l = NL - 1; l >= 0; --l) {
            const float* GI = F.GI[i][l];
            const float* GH = F.GH[i][l];
            const float* Y = F.Y[i][l];
            const float* h0 = F.H0[i] + (size_t)l * B * H;
            float* carry = nullptr;
            float* cnext = dhc0;
            for (int f = Fr - 1; f >= 0; --f) {
                const float* hp = f ? Y + (size_t)(f - 1) * H : h0;
                const int hp_ld = f ? Fr * H : H;
                float* part = (cnext == dhc0) ? dhc1 : dhc0;      // dh*z lands here, the GEMM adds dGH.W_hh into cnext
                SRNN_LAUNCH(k_gru_bwd_gates, dim3(cdiv(H, 128), B), 128, 0, st, GI + (size_t)f * 3 * H, GH + (size_t)f * 3 * H,
                            Fr * 3 * H, hp, hp_ld, dY + (size_t)f * H, Fr * H, carry, dGI + (size_t)f * 3 * H,
                            dGH + (size_t)f * 3 * H, part, H);
                // dh_{f-1} = dh*z + dGH_f . W_hh
                SRNN_TRY(gemm_s(B, H, 3 * H, dGH + (size_t)f * 3 * H, (long long)Fr * 3 * H, 1, t.w_hh[l], 1, H, part, H, cnext, H, st));
                carry = cnext;
                cnext = part;
            }
            // carry = dL/dh_{-1} (B,H): gradient of the learned initial state when this pass started from it
            float* dh0 = (float*)tg.h0;
            if (dh0) {
                if ((F.reset_mask >> i) & 1) SRNN_TRY(colsum(carry, B, H, H, csp, dh0 + (size_t)l * H, st));
                else SRNN_CUDA(cudaMemsetAsync(dh0 + (size_t)l * H, 0, sizeof(float) * H, st));   // carried state is detached
            }
            // recurrent weights: dW_hh = dGH^T . Hprev, rows (b,f) with Hprev = h_{f-1}
            float* dWhh = (float*)tg.weight_hh[l];
            if (dWhh) {
                SRNN_LAUNCH(k_build_hprev, M, 128, 0, st, Y, h0, dYn, Fr, H);      // dYn is free until `din` below
                SRNN_TRY(gemm_dw(3 * H, H, M, dGH, 3 * H, dYn, H, dWhh, H, st));
            }
            if (tg.bias_hh[l]) SRNN_TRY(colsum(dGH, M, 3 * H, 3 * H, csp, (float*)tg.bias_hh[l], st));
            // input weights
            const float* in = l ? F.Y[i][l - 1] : F.X[i];
            if (tg.weight_ih[l]) SRNN_TRY(gemm_dw(3 * H, H, M, dGI, 3 * H, in, H, (float*)tg.weight_ih[l], H, st));
            if (tg.bias_ih[l]) SRNN_TRY(colsum(dGI, M, 3 * H, 3 * H, csp, (float*)tg.bias_ih[l], st));
            // gradient wrt this layer's input = dY of the layer below (or dX)
            float* din = l ? dYn : dXbuf[0];
            SRNN_TRY(gemm_dx(M, H, 3 * H, dGI, 3 * H, t.w_ih[l], H, nullptr, 0, din, H, st));
            if (l) { float* tmp = dY; dY = dYn; dYn = tmp; }
        }
        float* dX = dXbuf[0];
        // input expansion: X = A W_in^T + b_in (+ upper)
        SRNN_TRY(gemm_s_splitk(H, t.kin, M, dX, 1, H, F.A[i], 1, t.kin, dWin, t.kin, dTblP, (size_t)DT_SEG * FS0 * Q * H, st));
        SRNN_TRY(colsum(dX, M, H, H, csp, dbtmp, st));
        if (t.top) {
            SRNN_TRY(wn_fold(tp.spk_expand, wsf, H, c.spk_dim, st));
            SRNN_LAUNCH(k_unpack_top_in, H, 128, 0, st, dWin, t_in, t_c, t_s, tp.spk_embedding, t.n, c.cond_dim, c.spk_dim);
            SRNN_TRY(wn_bwd(t_in, tp.input_expand, tg.input_expand, H, t.n, st));
            SRNN_TRY(wn_bwd(t_c, tp.cond_expand, tg.cond_expand, H, c.cond_dim, st));
            SRNN_TRY(wn_bwd(t_s, tp.spk_expand, tg.spk_expand, H, c.spk_dim, st));
            if (tg.spk_embedding)
                SRNN_LAUNCH(k_spk_emb_grad, c.spk_dim * c.spk_dim, 128, 0, st, dWin, wsf, (float*)tg.spk_embedding, H, t.kin,
                            t.n + c.cond_dim, c.spk_dim);
            if (tg.input_expand.bias) SRNN_TRY(copy_f32(dbtmp, (float*)tg.input_expand.bias, H, st));
            if (tg.cond_expand.bias) SRNN_TRY(copy_f32(dbtmp, (float*)tg.cond_expand.bias, H, st));
            if (tg.spk_expand.bias) SRNN_TRY(copy_f32(dbtmp, (float*)tg.spk_expand.bias, H, st));
        } else {
            SRNN_TRY(wn_bwd(dWin, tp.input_expand, tg.input_expand, H, t.n, st));
            if (tg.input_expand.bias) SRNN_TRY(copy_f32(dbtmp, (float*)tg.input_expand.bias, H, st));
            // d upper = dX: same memory viewed as (M_{i+1}, fs_{i+1}*H).  Keep it alive while the tier above runs.
            SRNN_TRY(copy_f32(dX, dA, (size_t)M * H, st));     // dA (R*H floats) is free by now
            dUP = dA;
        }
        // every gradient below the top tier is complete: a data-parallel caller may start reducing those while the top
        // tier's backward pass runs (srnn_bwd_wait_early)
        SRNN_CUDA(cudaEventRecord(ctx->ev_stage[2 + 2 * i], st));
        (void)xb;
    }
    return SRNN_OK;
}

// ------------------------------------------------------------------------------------------------
// bf16 / tcgen05 backward (SRNN_MODE_BF16 forward passes): same chain rule as predict_bwd_f32, with every
// H-wide contraction on the tensor cores through gemm_umma_multi:
//   dIn (rows, K)  = dOut (rows, N) . W (N, K)   : A operand = W^T (K, N) bf16 (packed once), B = dOut bf16
//   dW  (N, K)     = dOut^T . In                 : A = In^T (K, rows_p), B = dOut^T (N, rows_p), rows zero-padded to 64
// The transposed operands are produced by a tiled transpose+convert kernel (their HBM traffic is a few percent of the
// GEMM time at C3).  Element-wise pieces (GRU cell, masks, reductions) stay fp32.
// ------------------------------------------------------------------------------------------------
typedef __nv_bfloat16 bf;
static inline int rup64(long long x) { return (int)((x + 63) / 64 * 64); }
static inline int pick_bn2(int rows) { return rows <= 32 ? 32 : (rows <= 64 ? 64 : (rows <= 128 ? 128 : 256)); }

static int tc_dx(int rows, int Kin, int Nout, const bf* dOut16, int ld_do, const bf* Wt16, const float* addend, int ld_add,
                 float* outf, bf* outb, const bf* mask, int ld_out, cudaStream_t st) {
    GemmOperands o{Wt16, dOut16, nullptr, addend, outf, outb, Kin, Nout, ld_do, ld_add, ld_out, 0, mask};
    if (rows >= 256 && ld_out % 8 == 0 && ld_add % 4 == 0) return gemm_umma_rows(o, rows, Nout, 1, nullptr, st);
    return gemm_umma_multi(&o, 1, rows, Nout, 128, pick_bn2(rows), st);
}
// scratch: split-K partials (the K of a weight gradient is the token count: few output tiles, very long K loops)
static int tc_dw(int Nout, int Kin, int Kp, const bf* InT, const bf* dOutT, float* dW, float* scratch, size_t scratch_floats,
                 cudaStream_t st) {
    GemmOperands o{InT, dOutT, nullptr, nullptr, dW, nullptr, Kin, Kp, Kp, 0, Kin, 0, nullptr};
    if (Nout >= 128 && Kin % 8 == 0) {
        const int tiles = cdiv(Nout, 128) * cdiv(Kin, Kin <= 128 ? 128 : 256);
        int ks = cdiv(296, tiles);
        const size_t cap = scratch ? scratch_floats / ((size_t)Nout * Kin) : 0;
        if ((size_t)ks > cap) ks = (int)cap;
        if (ks > Kp / 512) ks = Kp / 512;              // at least 8 k-blocks per split
        return gemm_umma_rows(o, Nout, Kp, ks < 2 ? 1 : ks, scratch, st);
    }
    return gemm_umma_multi(&o, 1, Nout, Kp, 128, pick_bn2(Nout), st);
}

// dW (Nout, Kin) = dOut^T . In with dOut (rows, Nout) and In (rows, Kin) bf16 row-major read IN PLACE (MN-major UMMA operands):
// no transposed copies.  Split-K as in tc_dw.
static int tc_dw_tn(int Nout, int Kin, int rows, const bf* dOut16, int ld_do, const bf* In16, int ld_in, float* dW, float* scratch,
                    size_t scratch_floats, cudaStream_t st) {
    const int tiles = cdiv(Nout, 128) * cdiv(Kin, Kin <= 128 ? 128 : 256);
    int ks = cdiv(296, tiles);
    const size_t cap = scratch ? scratch_floats / ((size_t)Nout * Kin) : 0;
    if ((size_t)ks > cap) ks = (int)cap;
    if (ks > rows / 512) ks = rows / 512;                  // at least 8 k-blocks per split
    return gemm_umma_tn(dOut16, ld_do, In16, ld_in, Nout, Kin, rows, dW, Kin, ks < 2 ? 1 : ks, scratch, st);
}
// tbl_foldback on the tensor cores (bf16 training path): the two contractions are fp32 gradients of fp32 parameters, so they
// run as split-bf16 products (hi/lo pieces, see split3_bf16: ~2^-16 relative per product, fp32 accumulation) instead of FFMA
// GEMMs (0.45 ms of the C3 step).  Scratch comes from the stream-ordered pool (3 x 31 MB at C3).
static int tbl_foldback_x3(srnn_ctx* ctx, const srnn_params* P, const srnn_params* G, const float* dTblT, float* scratch2HQF,
                           size_t scratch_floats, float* dWmt, float* dWm, cudaStream_t st) {
    const int H = ctx->H, Q = ctx->Q, FS0 = ctx->FS0, KH = FS0 * H;
    const size_t n = (size_t)KH * Q;
    float* wm_fold = scratch2HQF;                       // (H, Q, FS)
    float* wm_t = scratch2HQF + n;                      // (FS, H, Q) = ((j,h), e)
    float* et = scratch2HQF + 2 * n;                    // (Q, Q) E^T
    float* splitk = et + (size_t)Q * Q;
    const size_t splitk_floats = scratch_floats - 2 * n - (size_t)Q * Q;
    bf *a3 = nullptr, *b3 = nullptr, *et3 = nullptr;
    SRNN_CUDA(cudaMallocAsync((void**)&a3, sizeof(bf) * 3 * n, st));
    SRNN_CUDA(cudaMallocAsync((void**)&b3, sizeof(bf) * 3 * n, st));
    SRNN_CUDA(cudaMallocAsync((void**)&et3, sizeof(bf) * 3 * (size_t)Q * Q, st));
    int rc = wn_fold(P->mlp_input, wm_fold, H, Q * FS0, st);
    if (rc == SRNN_OK) rc = transpose_mlp_in(wm_fold, wm_t, H, Q, FS0, st);
    // dWm_t[(j,h), e] = sum_q dTblT[(j,h), q] E[q, e]: rows (j,h), K = q, "weights" E^T (e, q)
    if (rc == SRNN_OK) rc = transpose_f32(P->embedding, et, Q, Q, st);
    if (rc == SRNN_OK) rc = split3_bf16(et, Q, Q, Q, et3, 1, st);
    if (rc == SRNN_OK) rc = gemm_x3(KH, Q, Q, dTblT, Q, et3, nullptr, 0, dWmt, Q, a3, 0, 0, st);
    // dE[q, e] = sum_{(j,h)} dTblT[(j,h), q] wm_t[(j,h), e]: both operands MN-major, the three pieces stacked along K
    float* dE = (float*)G->embedding;
    if (rc == SRNN_OK && dE) {
        rc = split3_planes_bf16(dTblT, n, a3, 1, st);
        if (rc == SRNN_OK) rc = split3_planes_bf16(wm_t, n, b3, 0, st);
        if (rc == SRNN_OK) rc = tc_dw_tn(Q, Q, 3 * KH, a3, Q, b3, Q, dE, splitk, splitk_floats, st);
    }
    cudaFreeAsync(a3, st);
    cudaFreeAsync(b3, st);
    cudaFreeAsync(et3, st);
    SRNN_TRY(rc);
    SRNN_LAUNCH(k_untranspose_mlp_in, dim3(H, FS0), 256, 0, st, dWmt, dWm, H, Q, FS0);
    SRNN_TRY(wn_bwd(dWm, P->mlp_input, G->mlp_input, H, Q * FS0, st));
    return SRNN_OK;
}

// HP16[(b,f), :] = f ? Y16[(b,f-1), :] : h0_16[b, :]   (bf16 recurrent input of every frame, for dW_hh)
__global__ void k_build_hprev16(const bf* __restrict__ Y16, const bf* __restrict__ h0, bf* __restrict__ hp, int F, int H) {
    const int r = blockIdx.x, b = r / F, f = r % F;
    const uint4* src = reinterpret_cast<const uint4*>(f ? Y16 + (size_t)(r - 1) * H : h0 + (size_t)b * H);
    uint4* dst = reinterpret_cast<uint4*>(hp + (size_t)r * H);
    for (int u = threadIdx.x; u < H / 8; u += blockDim.x) dst[u] = src[u];
}

size_t backward_scratch_bytes_bf16(const srnn_ctx* ctx, int B, int T) {
    const srnn_config& c = ctx->cfg;
    const size_t H = ctx->H, Q = ctx->Q, R = (size_t)B * T, Rp = rup64(R), FS0 = ctx->FS0;
    size_t maxM = 0, maxfs = 1, maxkin = 1;
    for (int i = 0; i < c.n_tiers; ++i) {
        const size_t M = (size_t)B * (T / ctx->tiers[i].n);
        if (M > maxM) maxM = M;
        if ((size_t)ctx->tiers[i].fs > maxfs) maxfs = ctx->tiers[i].fs;
        if ((size_t)ctx->tiers[i].kin > maxkin) maxkin = ctx->tiers[i].kin;
    }
    const size_t Mp = rup64(maxM);
    size_t f32 = R * Q + 3 * maxM * H + 2 * maxM * 3 * H + 2 * (size_t)B * H + 2 * maxfs * H * H + maxfs * H + H * maxkin +
                 H * (size_t)c.spk_dim + 3 * H * (H > maxkin ? H : maxkin) + (DT_SEG + 2) * FS0 * Q * H + 2 * FS0 * H * Q + (1024 + (size_t)B * (T + FS0) + 256 * ((size_t)B * (T + FS0) / 512 + 2))  +
                 (size_t)CS_CHUNKS * (maxfs * H > 3 * H ? maxfs * H : 3 * H) + 3 * H + Q * H + 3 * H * H;
    size_t b16 = R * Q + 2 * R * H                                            // D16, DP2, DP1
                 + 2 * maxM * 3 * H + 2 * maxM * H + 64;                       // dGI16, dGH16, HP16, DX16
    return f32 * sizeof(float) + b16 * sizeof(bf) + 96 * 256;
}

int predict_bwd_bf16(srnn_ctx* ctx, const float* logp, const float* dlogp, const srnn_params* P, const srnn_params* G,
                     cudaStream_t st, const int64_t* nll_target, const float* nll_gscale) {
    const FwdPlan& F = ctx->fwd;
    const srnn_config& c = ctx->cfg;
    const int H = ctx->H, Q = ctx->Q, NT = c.n_tiers, NL = c.n_rnn, lookback = ctx->lookback, FS0 = ctx->FS0;
    const int B = F.B, T = F.T, R = B * T, Rp = rup64(R), Lseq = lookback + T - 1;
    if (F.bytes + backward_scratch_bytes_bf16(ctx, B, T) > ctx->ws_bytes)
        return fail(SRNN_ERR_STATE, "backward scratch was not reserved by the forward pass");
    int maxfs = 1, maxkin = 1;
    size_t maxM = 0;
    for (int i = 0; i < NT; ++i) {
        const size_t M = (size_t)B * (T / ctx->tiers[i].n);
        if (M > maxM) maxM = M;
        if (ctx->tiers[i].fs > maxfs) maxfs = ctx->tiers[i].fs;
        if (ctx->tiers[i].kin > maxkin) maxkin = ctx->tiers[i].kin;
    }
    const size_t Mpmax = rup64(maxM);
    Bump2 b(ctx->ws, F.bytes);
    float* dlogits = b.take<float>((size_t)R * Q);
    float* dYa = b.take<float>(maxM * H);
    float* dYb = b.take<float>(maxM * H);
    float* dXf = b.take<float>(maxM * H);
    float* dGI = b.take<float>(maxM * 3 * H);
    float* dGH = b.take<float>(maxM * 3 * H);
    float* dhc0 = b.take<float>((size_t)B * H);
    float* dhc1 = b.take<float>((size_t)B * H);
    float* dWup = b.take<float>((size_t)maxfs * H * H);
    float* dwf = b.take<float>((size_t)maxfs * H * H);
    float* dbup = b.take<float>((size_t)maxfs * H);
    float* dWin = b.take<float>((size_t)H * maxkin);
    float* wsf = b.take<float>((size_t)H * c.spk_dim);
    const size_t stg = (size_t)H * (H > maxkin ? H : maxkin);
    float* t_in = b.take<float>(stg);
    float* t_c = b.take<float>(stg);
    float* t_s = b.take<float>(stg);
    float* dTblP = b.take<float>((size_t)DT_SEG * FS0 * Q * H);
    float* dTbl = b.take<float>((size_t)FS0 * Q * H);
    float* dTblT = b.take<float>((size_t)FS0 * Q * H);
    int* iwork = b.take<int>(1024 + (size_t)B * (T + FS0) + 256 * ((size_t)B * (T + FS0) / 512 + 2));
    float* dWmt = b.take<float>((size_t)FS0 * H * Q);
    float* dWm = b.take<float>((size_t)FS0 * H * Q);
    const size_t csp_floats = (size_t)CS_CHUNKS * (maxfs * H > 3 * H ? maxfs * H : 3 * H);
    const ColsumScratch csp{b.take<float>(csp_floats), csp_floats};
    float* dbtmp = b.take<float>((size_t)3 * H);
    float* dWo = b.take<float>((size_t)Q * H);
    float* dWtmp = b.take<float>((size_t)3 * H * H);
    const size_t dtblp_floats = (size_t)DT_SEG * FS0 * Q * H;   // dTblP doubles as split-K scratch outside dtbl_compute/foldback
    bf* D16 = b.take<bf>((size_t)R * Q);
    bf* DP2 = b.take<bf>((size_t)R * H);
    bf* DP1 = b.take<bf>((size_t)R * H);
    bf* dGI16 = b.take<bf>(maxM * 3 * H);
    bf* dGH16 = b.take<bf>(maxM * 3 * H);
    bf* HP16 = b.take<bf>(maxM * H);
    bf* DX16 = b.take<bf>(maxM * H);

    // ---- log_softmax backward; bf16 copies of dlogits in both orientations ----
    if (nll_target) {                                // fused loss: dlogits (fp32 + bf16) straight from logp and the targets
        SRNN_LAUNCH(k_nll_logsoftmax_bwd, cdiv(R, 8), 256, 0, st, logp, nll_target, nll_gscale, dlogits, D16, R);
    } else {
        SRNN_LAUNCH(k_logsoftmax_bwd, cdiv(R, 8), 256, 0, st, dlogp, logp, dlogits, R);
        SRNN_TRY(f32_to_bf16_pad(dlogits, R, Q, Q, D16, R, Q, st));
    }
    // ---- output layer ----
    SRNN_TRY(tc_dw_tn(Q, H, R, D16, Q, F.X2h, H, dWo, dTblP, dtblp_floats, st));
    SRNN_TRY(wn_bwd(dWo, P->mlp_output, G->mlp_output, Q, H, st));
    if (G->mlp_output.bias) SRNN_TRY(colsum(dlogits, R, Q, Q, csp, (float*)G->mlp_output.bias, st));
    SRNN_TRY(tc_dx(R, H, Q, D16, Q, ctx->w_out16_t, nullptr, 0, nullptr, DP2, F.X2h, H, st));   // dpre2 = dx2 * (x2 > 0)
    // ---- hidden layer ----
    SRNN_TRY(tc_dw_tn(H, H, R, DP2, H, F.X1h, H, dWtmp, dTblP, dtblp_floats, st));
    SRNN_TRY(wn_bwd(dWtmp, P->mlp_hidden, G->mlp_hidden, H, H, st));
    if (G->mlp_hidden.bias) SRNN_TRY(colsum(DP2, R, H, H, csp, (float*)G->mlp_hidden.bias, st));
    SRNN_TRY(tc_dx(R, H, H, DP2, H, ctx->w_hid16_t, nullptr, 0, nullptr, DP1, F.X1h, H, st));   // dpre1 = dc0
    // ---- folded table ----
    SRNN_TRY(dtbl_compute(F.seq, Lseq, lookback - FS0, DP1, B, T, H, FS0, iwork, dTblP, dTbl, dTblT, st));
    const bool foldback_f32 = getenv("SRNN_FOLDBACK_F32") != nullptr;             // A/B switch: the FFMA form
    if (foldback_f32 || Q % 64) SRNN_TRY(tbl_foldback(ctx, P, G, dTblT, dTblP, dWmt, dWm, st));
    else SRNN_TRY(tbl_foldback_x3(ctx, P, G, dTblT, dTblP, dtblp_floats, dWmt, dWm, st));
    SRNN_CUDA(cudaEventRecord(ctx->ev_stage[0], st));                           // MLP + embedding gradients are final
    // ---- frame tiers, lowest first ----
    const bf* dUP = DP1;
    for (int i = 0; i < NT; ++i) {
        const TierPacked& t = ctx->tiers[i];
        const srnn_tier_params& tp = P->tiers[i];
        const srnn_tier_params& tg = G->tiers[i];
        const int Fr = T / t.n, M = B * Fr, Mp = rup64(M), NU = t.fs * H;
        // upsampling
        SRNN_TRY(tc_dw_tn(NU, H, M, dUP, NU, F.Y16[i][NL - 1], H, dWup, dTblP, dtblp_floats, st));
        SRNN_TRY(colsum(dUP, M, NU, NU, csp, dbup, st));
        SRNN_TRY(unpack_up_grad(dWup, dbup, dwf, (float*)tg.upsampling.bias, H, t.fs, st));
        SRNN_TRY(wn_bwd(dwf, tp.upsampling, tg.upsampling, H, H * t.fs, st));
        SRNN_CUDA(cudaEventRecord(ctx->ev_stage[1 + 2 * i], st));                 // tier i's upsampling gradients are final
        float* dY = dYa;
        float* dYn = dYb;
        SRNN_TRY(tc_dx(M, H, NU, dUP, NU, t.w_up16_t, nullptr, 0, dY, nullptr, nullptr, H, st));
        for (int l = NL - 1; l >= 0; --l) {
            const float* GI = F.GI[i][l];
            const float* GH = F.GH[i][l];
            const float* Y = F.Y[i][l];
            const float* h0 = F.H0[i] + (size_t)l * B * H;
            float* carry = nullptr;
            float* cnext = dhc0;
            const bool persist = gru_persist_supported(B, H, ctx->n_sms);
            if (persist) {                                   // BPTT over all frames of the layer in one persistent launch
                if (!ctx->gru_ctr) SRNN_TRY(ctx->weights.alloc((void**)&ctx->gru_ctr, 256));
                // the kernel also sums the bias gradients, so the fp32 copies of dGI / dGH (only ever column-summed) are not written
                SRNN_TRY(gru_persist_bwd(B, Fr, H, GI, GH, Y, h0, dY, t.w_hh16_t[l], nullptr, nullptr, dGI16, dGH16, dhc0,
                                         ctx->gru_ctr, st, csp, (float*)tg.bias_ih[l], (float*)tg.bias_hh[l]));
                carry = dhc0;
            }
            for (int f = Fr - 1; f >= 0 && !persist; --f) {
                const float* hp = f ? Y + (size_t)(f - 1) * H : h0;
                const int hp_ld = f ? Fr * H : H;
                float* part = (cnext == dhc0) ? dhc1 : dhc0;
                SRNN_LAUNCH(k_gru_bwd_gates, dim3(cdiv(H, 128), B), 128, 0, st, GI + (size_t)f * 3 * H, GH + (size_t)f * 3 * H,
                            Fr * 3 * H, hp, hp_ld, dY + (size_t)f * H, Fr * H, carry, dGI + (size_t)f * 3 * H,
                            dGH + (size_t)f * 3 * H, part, H, dGI16 + (size_t)f * 3 * H, dGH16 + (size_t)f * 3 * H);
                // dh_{f-1} = dh*z + dGH_f . W_hh      (tensor cores, fp32 addend)
                SRNN_TRY(tc_dx(B, H, 3 * H, dGH16 + (size_t)f * 3 * H, Fr * 3 * H, t.w_hh16_t[l], part, H, cnext, nullptr,
                               nullptr, H, st));
                carry = cnext;
                cnext = part;
            }
            float* dh0 = (float*)tg.h0;
            if (dh0) {
                if ((F.reset_mask >> i) & 1) SRNN_TRY(colsum(carry, B, H, H, csp, dh0 + (size_t)l * H, st));
                else SRNN_CUDA(cudaMemsetAsync(dh0 + (size_t)l * H, 0, sizeof(float) * H, st));
            }
            if (tg.weight_hh[l]) {
                SRNN_LAUNCH(k_build_hprev16, M, 128, 0, st, F.Y16[i][l], F.H016[i] + (size_t)l * B * H, HP16, Fr, H);
                SRNN_TRY(tc_dw_tn(3 * H, H, M, dGH16, 3 * H, HP16, H, (float*)tg.weight_hh[l], dTblP, dtblp_floats, st));
            }
            if (tg.bias_hh[l] && !persist) SRNN_TRY(colsum(dGH, M, 3 * H, 3 * H, csp, (float*)tg.bias_hh[l], st));
            if (tg.weight_ih[l])
                SRNN_TRY(tc_dw_tn(3 * H, H, M, dGI16, 3 * H, l ? F.Y16[i][l - 1] : F.X16[i], H, (float*)tg.weight_ih[l], dTblP,
                                  dtblp_floats, st));
            if (tg.bias_ih[l] && !persist) SRNN_TRY(colsum(dGI, M, 3 * H, 3 * H, csp, (float*)tg.bias_ih[l], st));
            float* din = l ? dYn : dXf;
            SRNN_TRY(tc_dx(M, H, 3 * H, dGI16, 3 * H, t.w_ih16_t[l], nullptr, 0, din, nullptr, nullptr, H, st));
            if (l) { float* tmp = dY; dY = dYn; dYn = tmp; }
        }
        // input expansion (K = kin is small: fp32 FFMA)
        SRNN_TRY(gemm_s_splitk(H, t.kin, M, dXf, 1, H, F.A[i], 1, t.kin, dWin, t.kin, dTblP, (size_t)DT_SEG * FS0 * Q * H, st));
        SRNN_TRY(colsum(dXf, M, H, H, csp, dbtmp, st));
        if (t.top) {
            SRNN_TRY(wn_fold(tp.spk_expand, wsf, H, c.spk_dim, st));
            SRNN_LAUNCH(k_unpack_top_in, H, 128, 0, st, dWin, t_in, t_c, t_s, tp.spk_embedding, t.n, c.cond_dim, c.spk_dim);
            SRNN_TRY(wn_bwd(t_in, tp.input_expand, tg.input_expand, H, t.n, st));
            SRNN_TRY(wn_bwd(t_c, tp.cond_expand, tg.cond_expand, H, c.cond_dim, st));
            SRNN_TRY(wn_bwd(t_s, tp.spk_expand, tg.spk_expand, H, c.spk_dim, st));
            if (tg.spk_embedding)
                SRNN_LAUNCH(k_spk_emb_grad, c.spk_dim * c.spk_dim, 128, 0, st, dWin, wsf, (float*)tg.spk_embedding, H, t.kin,
                            t.n + c.cond_dim, c.spk_dim);
            if (tg.input_expand.bias) SRNN_TRY(copy_f32(dbtmp, (float*)tg.input_expand.bias, H, st));
            if (tg.cond_expand.bias) SRNN_TRY(copy_f32(dbtmp, (float*)tg.cond_expand.bias, H, st));
            if (tg.spk_expand.bias) SRNN_TRY(copy_f32(dbtmp, (float*)tg.spk_expand.bias, H, st));
        } else {
            SRNN_TRY(wn_bwd(dWin, tp.input_expand, tg.input_expand, H, t.n, st));
            if (tg.input_expand.bias) SRNN_TRY(copy_f32(dbtmp, (float*)tg.input_expand.bias, H, st));
            SRNN_TRY(f32_to_bf16_pad(dXf, M, H, H, DX16, M, H, st));     // d upper for the tier above, (M_{i+1}, fs_{i+1}*H)
            dUP = DX16;
        }
        SRNN_CUDA(cudaEventRecord(ctx->ev_stage[2 + 2 * i], st));
    }
    return SRNN_OK;
}

// ------------------------------------------------------------------------------------------------
// fused element-wise clamp to [-1, 1] + Adam over up to 64 tensors in one launch
// ------------------------------------------------------------------------------------------------
constexpr int ADAM_MAX = 64, ADAM_CHUNK = 4096;
struct AdamArgs {
    float* p[ADAM_MAX];
    const float* g[ADAM_MAX];
    float* m[ADAM_MAX];
    float* v[ADAM_MAX];
    int first_chunk[ADAM_MAX + 1];
    long long n[ADAM_MAX];
    int count;
    float lr_over_bc1, inv_sqrt_bc2, beta1, beta2, eps, clamp, grad_scale;
};
__global__ void k_clamp_adam(const __grid_constant__ AdamArgs a) {
    int t = 0;
    while (t + 1 < a.count && (int)blockIdx.x >= a.first_chunk[t + 1]) ++t;
    const long long base = (long long)((int)blockIdx.x - a.first_chunk[t]) * ADAM_CHUNK;
    float* __restrict__ p = a.p[t];
    const float* __restrict__ g = a.g[t];
    float* __restrict__ m = a.m[t];
    float* __restrict__ v = a.v[t];
    for (int i = threadIdx.x; i < ADAM_CHUNK; i += blockDim.x) {
        const long long idx = base + i;
        if (idx >= a.n[t]) break;
        float gg = fminf(fmaxf(g[idx] * a.grad_scale, -a.clamp), a.clamp);  // (mean over ranks, then) optim.py:10-13 hardtanh
        const float mm = a.beta1 * m[idx] + (1.f - a.beta1) * gg;
        const float vv = a.beta2 * v[idx] + (1.f - a.beta2) * gg * gg;
        m[idx] = mm;
        v[idx] = vv;
        p[idx] -= a.lr_over_bc1 * mm / (sqrtf(vv) * a.inv_sqrt_bc2 + a.eps);
    }
}

int clamp_adam(int count, float* const* params, const float* const* grads, float* const* m, float* const* v,
               const long long* sizes, float lr, float beta1, float beta2, float eps, int step, float clamp, cudaStream_t st,
               float grad_scale) {
    for (int s0 = 0; s0 < count; s0 += ADAM_MAX) {
        AdamArgs a;
        memset(&a, 0, sizeof(a));
        const int cnt = count - s0 < ADAM_MAX ? count - s0 : ADAM_MAX;
        int chunks = 0;
        for (int i = 0; i < cnt; ++i) {
            a.p[i] = params[s0 + i];
            a.g[i] = grads[s0 + i];
            a.m[i] = m[s0 + i];
            a.v[i] = v[s0 + i];
            a.n[i] = sizes[s0 + i];
            a.first_chunk[i] = chunks;
            chunks += (int)((sizes[s0 + i] + ADAM_CHUNK - 1) / ADAM_CHUNK);
        }
        a.first_chunk[cnt] = chunks;
        a.count = cnt;
        const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
        a.lr_over_bc1 = (float)(lr / bc1);
        a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
        a.beta1 = beta1;
        a.beta2 = beta2;
        a.eps = eps;
        a.clamp = clamp;
        a.grad_scale = grad_scale;
        if (chunks) SRNN_LAUNCH(k_clamp_adam, chunks, 256, 0, st, a);
    }
    return SRNN_OK;
}

}  // namespace srnn
