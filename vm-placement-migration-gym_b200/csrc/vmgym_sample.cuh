// vmgym_sample.cuh — random numbers of the multi-categorical sampler, shared by the stand-alone heads kernel
// (vmgym_policy.cu) and the fused GEMM epilogue (vmgym_gemm.cu) so both draw the same actions from the same logits.
//
// Sampling scheme (both kernels, identical arithmetic): STREAMING INVERSE-CDF.  A row's columns are visited in chunks
// of 32 with a running maximum m and running sum S of e^(z - m).  Chunk c with weight w_c replaces the current choice
// with probability w_c / S (one uniform u_c: take iff u_c * S < w_c — weighted reservoir sampling), and the column
// inside the chunk is found by inverse CDF with the SAME uniform (conditionally on being taken, u_c * S is uniform on
// [0, w_c)): first over the 8 groups of 4 columns, then inside the group.  P(column a) = e_a / S exactly; one Philox
// call per row instead of one Gumbel variate (two logarithms + a quarter Philox call) per column.
#pragma once
#include <stdint.h>

#include "vmgym_device.cuh"

namespace vmgym {

// One Philox4x32-7 call (7 rounds: the fastest Crush-resistant variant of Salmon et al.) serves 4 consecutive columns
// of row (env, v): counter (v*64 + a/4, env, 4, call counter).
__device__ __forceinline__ Philox4 sample_block(int v, int a4, uint32_t env, unsigned long long seed, uint32_t counter)
{
    uint32_t c0 = (uint32_t)(v * 64 + a4), c1 = env, c2 = 4u, c3 = counter, k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 7; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0, hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}
// Gumbel(0,1) noise from 24 random bits; the fast log intrinsic is ample for sampling noise and is used by both kernels
__device__ __forceinline__ float gumbel_from(const Philox4& r, int sub)
{
    const uint32_t bits = sub == 0 ? r.x : (sub == 1 ? r.y : (sub == 2 ? r.z : r.w));
    const float u = ((float)(bits >> 8) + 0.5f) * (1.0f / 16777216.0f);     // (0,1)
    return -__logf(-__logf(u));
}

// e^x as ex2.approx.ftz(x * log2 e): two instructions (the `__expf` intrinsic adds denormal handling around the same MUFU).
// Both samplers use THIS function so that they see identical weights.
__device__ __forceinline__ float fast_exp(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}

// The migration-ratio gate of PPOAgent.act (ppo.py:153-155) draws one uniform per VM row: word v & 3 of the Philox4x32-7
// block keyed (v >> 2, env, 3, call counter) — one call serves four rows.  Used by heads_kernel and mask_bits_kernel alike.
__device__ __forceinline__ float gate_uniform(int v, uint32_t env, unsigned long long seed, uint32_t counter)
{
    uint32_t c0 = (uint32_t)(v >> 2), c1 = env, c2 = 3u, c3 = counter, k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 7; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0, hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const uint32_t bits = (v & 3) == 0 ? c0 : ((v & 3) == 1 ? c1 : ((v & 3) == 2 ? c2 : c3));
    return (float)(bits >> 8) * (1.0f / 16777216.0f);
}

// uniform in [0, 1) for chunk c of a row, from the row's block(s): 24 bits of word c & 3
__device__ __forceinline__ float chunk_uniform(const Philox4& r, int c)
{
    const uint32_t bits = (c & 3) == 0 ? r.x : ((c & 3) == 1 ? r.y : ((c & 3) == 2 ? r.z : r.w));
    return (float)(bits >> 8) * (1.0f / 16777216.0f);
}

}  // namespace vmgym
