// The defined sampler that replaces Tensor.multinomial (model.py:517).  One warp draws one 256-way row.
// The summation order is part of the definition (oracle/srnn_oracle.py:sample_rows mirrors it in numpy):
//   lane l owns entries 8l..8l+7 and forms sequential running sums loc[i];
//   lane totals go through a 5-step Kogge-Stone inclusive scan;  cdf[8l+i] = excl[l] + loc[i];
//   idx = min(255, #{k : cdf[k] <= u * total}).
// All adds/muls are explicit round-to-nearest intrinsics so the compiler can neither fuse nor reorder them.
#pragma once

namespace srnn {

__device__ __forceinline__ int sampler_warp(const float (&p)[8], float u, int lane) {
    float loc[8];
    float acc = p[0];
    loc[0] = acc;
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        acc = __fadd_rn(acc, p[i]);
        loc[i] = acc;
    }
    float S = acc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, S, d);
        if (lane >= d) S = __fadd_rn(S, t);
    }
    float excl = __shfl_up_sync(0xffffffffu, S, 1);
    if (lane == 0) excl = 0.f;
    const float total = __shfl_sync(0xffffffffu, S, 31);
    const float thr = __fmul_rn(u, total);
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) cnt += (__fadd_rn(excl, loc[i]) <= thr) ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);        // integer warp reduction in one instruction (redux.sync): exact
    return cnt > 255 ? 255 : cnt;
}

}  // namespace srnn
