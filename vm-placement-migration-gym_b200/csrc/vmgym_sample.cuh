// vmgym_sample.cuh — the Gumbel noise of the multi-categorical sampler, shared by the stand-alone heads kernel
// (vmgym_policy.cu) and the fused GEMM epilogue (vmgym_gemm.cu) so both draw the same actions from the same logits.
#pragma once
#include <stdint.h>

#include "vmgym_device.cuh"

namespace vmgym {

// One Philox4x32-10 call serves 4 consecutive columns of row (env, v): counter (v*64 + a/4, env, 4, call counter).
__device__ __forceinline__ Philox4 sample_block(int v, int a4, uint32_t env, unsigned long long seed, uint32_t counter)
{
    return philox4x32_10((uint32_t)(v * 64 + a4), env, 4u, counter, (uint32_t)seed, (uint32_t)(seed >> 32));
}
__device__ __forceinline__ float gumbel_from(const Philox4& r, int sub)
{
    const uint32_t bits = sub == 0 ? r.x : (sub == 1 ? r.y : (sub == 2 ? r.z : r.w));
    const float u = ((float)(bits >> 8) + 0.5f) * (1.0f / 16777216.0f);     // (0,1)
    return -logf(-logf(u));
}

}  // namespace vmgym
