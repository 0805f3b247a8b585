// vmgym_gemm.cu — the one dense contraction of the hot path on Blackwell tensor cores (sm_100a):
//   C[M, N] (fp32) = A[M, K] (bf16, K-major) · W[N, K]^T (bf16, K-major = nn.Linear.weight) + bias[N]
// used for the actor's output layer (src/agents/ppo.py:103-109, Linear(hidden, V·A): 512 -> 30 600 at 100 PMs),
// which is 95 % of the policy's FLOPs.
//
// Hand-written tcgen05 pipeline, one 128x128 output tile per CTA:
//   warp 0   : TMA producer  (cp.async.bulk.tensor.2d, 128B-swizzled 128x64 bf16 boxes of A and W, mbarrier tx-count)
//   warp 1   : MMA issuer    (one elected lane: tcgen05.mma.cta_group::1.kind::f16, M=128 N=128 K=16, accumulator in TMEM;
//                             tcgen05.commit releases smem stages / signals the epilogue)
//   warps 2-5: epilogue      (tcgen05.ld 32x32b.x32 TMEM -> registers, + bias, fp32 stores)
// SASS: UTMALDG (TMA), UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/vmgym.h"
#include "vmgym_sample.cuh"
#include "vmgym_tc.cuh"

extern "C" void vmgym_internal_set_error(const char* msg);

namespace vmgym_gemm {
// ---- the main loop both kernels share: one 128x128 fp32 accumulator tile in TMEM per CTA --------------------------------
// shared-memory carve-up: STAGES x (A tile | W tile), then the barriers, the TMEM base address and 1 KiB of kernel-specific tail
struct Pipe {
    unsigned char* smem;              // 1024-aligned stage ring
    uint64_t *full_bar, *empty_bar, *tmem_full_bar;
    uint32_t* tmem_ptr;
};
__device__ __forceinline__ Pipe pipe_carve(unsigned char* smem_raw)
{
    Pipe p;
    p.smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    p.full_bar = reinterpret_cast<uint64_t*>(p.smem + (size_t)STAGES * STAGE_BYTES);
    p.empty_bar = p.full_bar + STAGES;
    p.tmem_full_bar = p.empty_bar + STAGES;
    p.tmem_ptr = reinterpret_cast<uint32_t*>(p.tmem_full_bar + 1);
    return p;
}
// descriptor prefetch (warp 0), barrier init (warp 1), TMEM allocation (warp 2); the caller syncs the CTA afterwards
__device__ __forceinline__ void pipe_setup(const Pipe& p, const CUtensorMap* map_a, const CUtensorMap* map_w, int warp, int lane)
{
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(map_w) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&p.full_bar[s], 1); mbar_init(&p.empty_bar[s], 1); }
        mbar_init(p.tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // one warp allocates the accumulator columns and publishes the TMEM base address
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(p.tmem_ptr)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
}
__device__ __forceinline__ uint32_t pipe_sync_tmem_base(const Pipe& p)
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *p.tmem_ptr;
}
// TMA producer (one thread): K slices of the A rows [m0, m0+128) and the W rows [n0, n0+128) into the stage ring
__device__ __forceinline__ void pipe_produce(const Pipe& p, const CUtensorMap* map_a, const CUtensorMap* map_w, int m0, int n0, int k_blocks)
{
    for (int kb = 0; kb < k_blocks; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&p.empty_bar[s], ph ^ 1u);                    // slot free (passes immediately on the first round)
        unsigned char* sa = p.smem + (size_t)s * STAGE_BYTES;
        unsigned char* sb = sa + BM * BK * 2;
        mbar_expect_tx(&p.full_bar[s], STAGE_BYTES);
        tma_load_2d(sa, map_a, kb * BK, m0, &p.full_bar[s]);
        tma_load_2d(sb, map_w, kb * BK, n0, &p.full_bar[s]);
    }
}
// MMA issuer (one thread): 4 x (M128 N128 K16) per stage into the TMEM accumulator, stage release and final commit
__device__ __forceinline__ void pipe_mma(const Pipe& p, uint32_t tmem_base, int k_blocks)
{
    const uint32_t idesc = umma_idesc_bf16_f32(BM, BN);
    for (int kb = 0; kb < k_blocks; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&p.full_bar[s], ph);                          // TMA bytes have landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(p.smem + (size_t)s * STAGE_BYTES);
        const uint32_t sb = sa + BM * BK * 2;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; k++) {
            // advance 16 bf16 = 32 B inside the 128-B swizzle row
            umma_bf16(tmem_base, umma_desc_sw128(sa + k * UMMA_K * 2), umma_desc_sw128(sb + k * UMMA_K * 2), idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&p.empty_bar[s]);                           // frees the smem stage when these MMAs retire
    }
    umma_commit(p.tmem_full_bar);                               // accumulator complete
}
__device__ __forceinline__ void pipe_teardown(uint32_t tmem_base, int warp)
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) linear_bf16_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                  const __grid_constant__ CUtensorMap map_w,
                                                                  const float* __restrict__ bias, float* __restrict__ C, int M,
                                                                  int N, int K, long long ldc)
{
    extern __shared__ unsigned char smem_raw[];
    const Pipe pipe = pipe_carve(smem_raw);
    unsigned char* smem = pipe.smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_blocks = (K + BK - 1) / BK;
    pipe_setup(pipe, &map_a, &map_w, warp, lane);
    const uint32_t tmem_base = pipe_sync_tmem_base(pipe);

    if (warp == 0) {
        if (lane == 0) pipe_produce(pipe, &map_a, &map_w, m0, n0, k_blocks);
    } else if (warp == 1) {
        if (lane == 0) pipe_mma(pipe, tmem_base, k_blocks);
    } else {
        // ===== epilogue: warps 2..5 own TMEM lanes 32*(warp%4) .. +31 (= rows of the tile) =====
        const int q = warp & 3;
        mbar_wait(pipe.tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // All smem stages are free once the accumulator is complete (one tile per CTA): each epilogue warp uses
        // 32 rows x 36 floats of it to turn "thread = row" (TMEM layout) into "lanes = consecutive columns" so the
        // fp32 tile leaves as full 128-byte segments.
        float* tr = reinterpret_cast<float*>(smem) + q * (32 * 36);
        const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            tmem_ld_row32(taddr, r);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(tr + lane * 36 + j) =
                    make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            __syncwarp();
            // lanes 0-7 cover the 32 columns of one row with float4; 4 rows per instruction
            const int cl = (lane & 7) * 4, rsub = lane >> 3;
            const int col = n0 + c0 + cl;
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias) {
                bv.x = col < N ? bias[col] : 0.f; bv.y = col + 1 < N ? bias[col + 1] : 0.f;
                bv.z = col + 2 < N ? bias[col + 2] : 0.f; bv.w = col + 3 < N ? bias[col + 3] : 0.f;
            }
#pragma unroll
            for (int rr = 0; rr < 32; rr += 4) {
                const int rloc = rr + rsub, row = m0 + q * 32 + rloc;
                float4 v = *reinterpret_cast<const float4*>(tr + rloc * 36 + cl);
                v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                if (row < M) {
                    float* dst = C + (long long)row * ldc + col;
                    if (vec_ok && col + 3 < N) *reinterpret_cast<float4*>(dst) = v;
                    else {
                        if (col < N) dst[0] = v.x;
                        if (col + 1 < N) dst[1] = v.y;
                        if (col + 2 < N) dst[2] = v.z;
                        if (col + 3 < N) dst[3] = v.w;
                    }
                }
            }
        }
    }
    pipe_teardown(tmem_base, warp);
}

// ---------------------------------------------------------------------------------------------------
// Fused actor head: the same TMA / tcgen05 main loop, but the epilogue consumes the 128x128 fp32 accumulator
// straight from TMEM — one tile = the A_pad = 128 logits of ONE VM for 128 envs — and emits the sampled action,
// its log-prob and the row entropy (Network.get_action, ppo.py:115-126).  The logits never reach HBM.
// Weights are re-laid out once on the host so VM v occupies rows [128 v, 128 v + A) (zero rows above A).
// The invalid-action bits (env.py:45-53 + the gating of ppo.py:153-155) come packed, 4 words per (env, VM).
// ---------------------------------------------------------------------------------------------------
struct FusedOut {
    const uint32_t* mask_bits;   // [M, V, 4] or nullptr (unmasked)
    const float* bias_pad;       // [V * 128]
    const void* action_in;       // evaluate these actions (u8) or nullptr
    uint8_t* action_out;         // [M, V]
    float* logprob;              // [M, V]
    float* entropy;              // [M, V]
    int A, V;
    unsigned long long seed;
    uint32_t counter;
    // gradient mode (PPOAgent.update, ppo.py:257-285): with action_in, additionally emit the gradient of
    //   sum_e c_logprob[e] * logprob(e) + c_entropy * entropy(e)   w.r.t. the logits, as bf16 g_out[e, v * 128 + a]
    // (d logp_act / dz_a = [a == act] - p_a;  dH / dz_a = -p_a (log p_a + H);  masked and padding columns get 0)
    const float* c_logprob;      // [M] or nullptr
    float c_entropy;
    __nv_bfloat16* g_out;        // [M, ldg] or nullptr
    long long ldg;
    // softmax statistics per (env, VM) row, [M, V] each: row maximum and sum of e^(z - max).  Written by an evaluating forward
    // call (action_in given) when non-null; READ by the gradient mode, which then needs a single pass over the accumulator.
    float* stat_m;
    float* stat_s;
    int apad;                    // rows per VM in w_pad / bias_pad and columns per VM in g_out: vmgym_policy_fused_rows(A, K)
};

// Per-row running state of the fused epilogue.
struct RowState {
    float m, ssum, tsum, best_z, z_given;
    int best_a;
};

// One chunk of 32 accumulator columns of this thread's row.  FULL: all 32 columns are real actions — straight-line code,
// 32 independent exponentials in flight.  !FULL (the last chunk when A is not a multiple of 32): only the groups of 4
// columns that hold real actions are touched (A = 102: 2 of 8 groups).
template <bool FULL>
__device__ __forceinline__ void fused_chunk(const FusedOut& fo, const float* s_bias, const uint32_t (&r)[32], uint32_t iw, int c0, int A,
                                            int act_given, const vmgym::Philox4& rnd, RowState& st)
{
    const int ncol = FULL ? 32 : min(32, A - c0), ngr = FULL ? 8 : (ncol + 3) >> 2;
    float z[32];
    float cm = -1e30f;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        if (FULL || t < ngr) {
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const int j = 4 * t + jj;
                float x = __uint_as_float(r[j]) + s_bias[c0 + j];
                if ((iw >> j) & 1u) x = -1e7f;                       // ppo.py:119
                if (!FULL && j >= ncol) x = -1e30f;                  // padding columns inside the last group
                z[j] = x;
                cm = fmaxf(cm, x);
            }
        } else {
            z[4 * t] = z[4 * t + 1] = z[4 * t + 2] = z[4 * t + 3] = -1e30f;
        }
    }
    float m = st.m, ssum = st.ssum, tsum = st.tsum;
    if (cm > m) {
        if (c0 > 0) { const float d = cm - m, sc = vmgym::fast_exp(-d); tsum = (tsum - d * ssum) * sc; ssum *= sc; }
        m = cm;
    }
    // e^(z - m) summed as a binary tree over 8 groups of 4 columns (= the order of a warp butterfly, which is what the
    // stand-alone heads kernel uses): group sums g[t], chunk weight w
    float g[8];
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        g[t] = 0.f;
        if (FULL || t < ngr) {
            const float x0 = z[4 * t] - m, x1 = z[4 * t + 1] - m, x2 = z[4 * t + 2] - m, x3 = z[4 * t + 3] - m;
            const float e0 = vmgym::fast_exp(x0), e1 = vmgym::fast_exp(x1), e2 = vmgym::fast_exp(x2), e3 = vmgym::fast_exp(x3);
            g[t] = (e0 + e1) + (e2 + e3);
            t0 = __fmaf_rn(e0, x0, t0); t1 = __fmaf_rn(e1, x1, t1); t2 = __fmaf_rn(e2, x2, t2); t3 = __fmaf_rn(e3, x3, t3);
        }
    }
    const float w = ((g[0] + g[1]) + (g[2] + g[3])) + ((g[4] + g[5]) + (g[6] + g[7]));
    const float ssum_new = ssum + w;
    tsum += (t0 + t1) + (t2 + t3);
    if (fo.action_in) {
        // one-hot word of the stored action inside this chunk: moved to predicates 7 bits at a time (R2P), the select costs one
        // instruction per column instead of a compare + select
        const uint32_t hot = (unsigned)(act_given - c0) < 32u ? 1u << (act_given - c0) : 0u;
#pragma unroll
        for (int j = 0; j < 32; j++) if ((hot >> j) & 1u) st.z_given = z[j];
    } else {
        // streaming inverse-CDF (vmgym_sample.cuh): this chunk replaces the choice iff u * S < w; the column is where the
        // cumulative sum passes u * S — first over the groups, then inside the group.  Branch-free.
        const float target = vmgym::chunk_uniform(rnd, c0 >> 5) * ssum_new;
        float cum = 0.f, base = 0.f, blast = 0.f;
        int tsel = -1, tlast = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const float prev = cum;
            cum += g[t];
            if (g[t] > 0.f) { tlast = t; blast = prev; }
            if (tsel < 0 && cum > target) { tsel = t; base = prev; }
        }
        if (tsel < 0) { tsel = tlast; base = blast; }
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;          // the 4 logits of group tsel, without a dynamic register index
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const bool hit = t == tsel;
            a0 = hit ? z[4 * t] : a0; a1 = hit ? z[4 * t + 1] : a1; a2 = hit ? z[4 * t + 2] : a2; a3 = hit ? z[4 * t + 3] : a3;
            asm volatile("" : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3));      // keep the select chain a select chain
        }
        const float e0 = vmgym::fast_exp(a0 - m), e1 = vmgym::fast_exp(a1 - m), e2 = vmgym::fast_exp(a2 - m), e3 = vmgym::fast_exp(a3 - m);
        int k = -1, klast = 0;
        float c2 = base + e0;
        if (c2 > target) k = 0;
        c2 += e1; if (e1 > 0.f) klast = 1; if (k < 0 && c2 > target) k = 1;
        c2 += e2; if (e2 > 0.f) klast = 2; if (k < 0 && c2 > target) k = 2;
        c2 += e3; if (e3 > 0.f) klast = 3; if (k < 0 && c2 > target) k = 3;
        if (k < 0) k = klast;
        if (target < w) {
            st.best_a = c0 + 4 * tsel + k;
            st.best_z = k == 0 ? a0 : (k == 1 ? a1 : (k == 2 ? a2 : a3));
        }
    }
    st.m = m; st.ssum = ssum_new; st.tsum = tsum;
}

// The fused epilogue for ONE row (thread = env row e) of ONE VM tile: masked log-softmax statistics relative to a running
// max (s = sum e^(z-m), t = sum e^(z-m) (z-m), merged chunk by chunk with one rescale per chunk), streaming inverse-CDF
// sample (or evaluation of a stored action), log-prob and entropy, straight out of the TMEM accumulator `tmem_acc`
// (column 0 of this tile's accumulator; the lane offset of the calling warp is added here).
// What a row's epilogue needs from global memory and the RNG — fetched BEFORE waiting for the accumulator so that the
// loads (one 16-byte mask word and one action byte per row, 4.8 KB apart between rows) overlap the MMA.
struct RowIn {
    uint4 inv;
    int act_given;
    vmgym::Philox4 rnd;
    float m, s;                  // gradient mode: the row's softmax statistics from the forward call
};
__device__ __forceinline__ RowIn fused_epilogue_prefetch(const FusedOut& fo, int q, int lane, int m0, int v, int M)
{
    const int e = m0 + q * 32 + lane;
    RowIn in;
    in.inv = make_uint4(0u, 0u, 0u, 0u);
    if (fo.mask_bits && e < M) in.inv = reinterpret_cast<const uint4*>(fo.mask_bits)[(long long)e * fo.V + v];
    in.act_given = -1;
    if (fo.action_in && e < M) in.act_given = reinterpret_cast<const uint8_t*>(fo.action_in)[(long long)e * fo.V + v];
    // sampling: one Philox call per row, word c = the uniform of chunk c (vmgym_sample.cuh)
    in.rnd = vmgym::Philox4{0u, 0u, 0u, 0u};
    if (!fo.action_in) in.rnd = vmgym::sample_block(v, 0, (uint32_t)e, fo.seed, fo.counter);
    in.m = 0.f; in.s = 1.f;
    if (fo.g_out && e < M) { in.m = fo.stat_m[(long long)e * fo.V + v]; in.s = fo.stat_s[(long long)e * fo.V + v]; }
    return in;
}
__device__ __forceinline__ void fused_epilogue_row(const FusedOut& fo, const float* s_bias, uint32_t tmem_acc, const RowIn& in, int q,
                                                   int lane, int m0, int v, int M)
{
    const int e = m0 + q * 32 + lane;
    const int A = fo.A;
    const uint4 inv = in.inv;
    const int act_given = in.act_given;
    const vmgym::Philox4 rnd = in.rnd;
    RowState st = {-1e30f, 0.f, 0.f, 0.f, 0.f, 0};
#pragma unroll 1
    for (int c0 = 0; c0 < BN && c0 < A; c0 += 32) {
        uint32_t r[32];
        tmem_ld_row32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        const uint32_t iw = c0 == 0 ? inv.x : (c0 == 32 ? inv.y : (c0 == 64 ? inv.z : inv.w));
        if (c0 + 32 <= A) fused_chunk<true>(fo, s_bias, r, iw, c0, A, act_given, rnd, st);
        else fused_chunk<false>(fo, s_bias, r, iw, c0, A, act_given, rnd, st);
    }
    const float ls = __logf(st.ssum);
    // The reference normalises in float32: logits - logsumexp with logsumexp = fl(max + log(sum e^(z - max))) (torch Categorical).
    // For a row whose columns are ALL masked (max = -1e7, float32 spacing 1.0 there) that rounding is visible — log-prob -5.0
    // instead of -log(101) — so the same rounded log-sum-exp is used here; for ordinary rows it differs by ~1e-6.
    const float lse = st.m + ls;
    const float H = (lse - st.m) - st.tsum / st.ssum;                // -sum p log p
    if (e < M) {
        const int act = fo.action_in ? act_given : st.best_a;
        const float za = fo.action_in ? st.z_given : st.best_z;
        const long long o = (long long)e * fo.V + v;
        if (fo.action_out) fo.action_out[o] = (uint8_t)act;
        if (fo.logprob) fo.logprob[o] = ((unsigned)act < (unsigned)A) ? za - lse : 0.f;
        if (fo.entropy) fo.entropy[o] = H;
        if (fo.stat_m) { fo.stat_m[o] = st.m; fo.stat_s[o] = st.ssum; }       // for the gradient mode's single pass
    }
}

// ---------------------------------------------------------------------------------------------------
// Evaluate-mode epilogue (PPOAgent.update: stored actions under stored masks; ppo.py:257-258 and its backward).
// At saturation almost every column of a row is masked (a running VM may only stay or be suspended: 2 valid columns of 102), and a
// masked column contributes exactly 0 to the softmax sums and gets exactly 0 gradient.  The math therefore visits only the columns
// that are valid in AT LEAST ONE of the warp's 32 rows (warp-uniform loop over the OR of the rows' valid bits; the accumulator
// value of column j is picked from the tcgen05.ld registers by a 32-way switch, so no register array is indexed dynamically).
// A row without any valid column gets no gradient at all (its logits were overwritten by the constant -1e7).  One pass over the
// accumulator, using the row statistics (max, sum) the evaluating forward call stored.  Chunks in which many columns are in
// play (an env far from saturation) take straight-line code instead of the loop.
// (Measured: the forward is bound by its MMA / TMA pipeline, not by its epilogue — a valid-column forward was slower — so the
// forward keeps the straight-line epilogue of the rollout kernel and only stores the statistics.)
// ---------------------------------------------------------------------------------------------------
#define VMGYM_PICK32(r, j, x)                                                                                                            \
    switch (j) {                                                                                                                         \
    case 0: x = r[0]; break; case 1: x = r[1]; break; case 2: x = r[2]; break; case 3: x = r[3]; break;                                  \
    case 4: x = r[4]; break; case 5: x = r[5]; break; case 6: x = r[6]; break; case 7: x = r[7]; break;                                  \
    case 8: x = r[8]; break; case 9: x = r[9]; break; case 10: x = r[10]; break; case 11: x = r[11]; break;                              \
    case 12: x = r[12]; break; case 13: x = r[13]; break; case 14: x = r[14]; break; case 15: x = r[15]; break;                          \
    case 16: x = r[16]; break; case 17: x = r[17]; break; case 18: x = r[18]; break; case 19: x = r[19]; break;                          \
    case 20: x = r[20]; break; case 21: x = r[21]; break; case 22: x = r[22]; break; case 23: x = r[23]; break;                          \
    case 24: x = r[24]; break; case 25: x = r[25]; break; case 26: x = r[26]; break; case 27: x = r[27]; break;                          \
    case 28: x = r[28]; break; case 29: x = r[29]; break; case 30: x = r[30]; break; default: x = r[31]; break;                          \
    }
#define VMGYM_PUT16(gb, k, w)                                                                                                            \
    switch (k) {                                                                                                                         \
    case 0: gb[0] = w; break; case 1: gb[1] = w; break; case 2: gb[2] = w; break; case 3: gb[3] = w; break;                              \
    case 4: gb[4] = w; break; case 5: gb[5] = w; break; case 6: gb[6] = w; break; case 7: gb[7] = w; break;                              \
    case 8: gb[8] = w; break; case 9: gb[9] = w; break; case 10: gb[10] = w; break; case 11: gb[11] = w; break;                          \
    case 12: gb[12] = w; break; case 13: gb[13] = w; break; case 14: gb[14] = w; break; default: gb[15] = w; break;                      \
    }

constexpr int EVAL_DENSE = 9;       // columns in play per 32-column chunk above which the straight-line path is cheaper than the loop

__device__ __forceinline__ uint32_t eval_valid_bits(const FusedOut& fo, const uint4& inv, int c0, int A, bool row_live)
{
    const uint32_t iw = c0 == 0 ? inv.x : (c0 == 32 ? inv.y : (c0 == 64 ? inv.z : inv.w));
    const uint32_t real = A - c0 >= 32 ? 0xffffffffu : ((1u << (A - c0)) - 1u);
    return row_live ? (~iw & real) : 0u;
}

__device__ __forceinline__ void eval_epilogue_row(const FusedOut& fo, const float* s_bias, uint32_t tmem_acc, const RowIn& in, int q,
                                                  int lane, int m0, int v, int M, uint32_t stage_addr)
{
    const int e = m0 + q * 32 + lane;
    const int A = fo.A;
    const bool live = e < M;
    const uint4 inv = in.inv;
    const int act = in.act_given;
    const uint32_t trow = tmem_acc + ((uint32_t)(q * 32) << 16);
    const bool any_valid = (eval_valid_bits(fo, inv, 0, A, live) | (A > 32 ? eval_valid_bits(fo, inv, 32, A, live) : 0u) |
                            (A > 64 ? eval_valid_bits(fo, inv, 64, A, live) : 0u) | (A > 96 ? eval_valid_bits(fo, inv, 96, A, live) : 0u)) != 0u;
    // ---- gradient: one pass; g_a = c_lp ([a == act] - p_a) - c_ent p_a (log p_a + H) on the valid columns, 0 elsewhere ----
    const float m = in.m, ssum = in.s;
    const float ls = __logf(ssum), lse = m + ls, inv_s = 1.0f / ssum;
    const float clp = live ? fo.c_logprob[e] : 0.f, cen = fo.c_entropy;
    // H = (lse - m) - tsum / ssum needs tsum: recomputed here from the same pass would need two passes, so the forward's
    // entropy output is used instead (it is the same number)
    const float H = (live && fo.entropy) ? fo.entropy[(long long)e * fo.V + v] : 0.f;
    __nv_bfloat16* grow = fo.g_out + (long long)e * fo.ldg + (long long)v * fo.apad;
#pragma unroll 1
    for (int c0 = 0; c0 < fo.apad; c0 += 32) {
        uint32_t gb[16];
#pragma unroll
        for (int k = 0; k < 16; k++) gb[k] = 0u;
        if (c0 < A) {
            uint32_t r[32];
            tmem_ld_row32(trow + (uint32_t)c0, r);
            const uint32_t vb = any_valid ? eval_valid_bits(fo, inv, c0, A, live) : 0u;
            uint32_t u = __reduce_or_sync(0xffffffffu, vb);
            if (__popc(u) > EVAL_DENSE) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float gv[2];
#pragma unroll
                    for (int t = 0; t < 2; t++) {
                        const float x = __uint_as_float(r[j + t]) + s_bias[c0 + j + t];
                        float gg = 0.f;
                        if ((vb >> (j + t)) & 1u) {
                            const float pa = vmgym::fast_exp(x - m) * inv_s;
                            gg = clp * ((c0 + j + t == act ? 1.f : 0.f) - pa) + (pa > 0.f ? cen * (-pa * ((x - lse) + H)) : 0.f);
                        }
                        gv[t] = gg;
                    }
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(gv[0], gv[1]);
                    gb[j >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                u = 0u;
            }
            while (u) {
                const int j = __ffs(u) - 1;
                u &= u - 1;
                uint32_t xr;
                VMGYM_PICK32(r, j, xr);
                const float x = __uint_as_float(xr) + s_bias[c0 + j];
                float gg = 0.f;
                if ((vb >> j) & 1u) {
                    const float pa = vmgym::fast_exp(x - m) * inv_s;
                    gg = clp * ((c0 + j == act ? 1.f : 0.f) - pa) + (pa > 0.f ? cen * (-pa * ((x - lse) + H)) : 0.f);
                }
                // the column's bf16 goes into its half of word j / 2 (the other half may already hold its neighbour)
                const uint32_t hb = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(gg));
                const int k = j >> 1;
                uint32_t wcur;
                switch (k) {
                case 0: wcur = gb[0]; break; case 1: wcur = gb[1]; break; case 2: wcur = gb[2]; break; case 3: wcur = gb[3]; break;
                case 4: wcur = gb[4]; break; case 5: wcur = gb[5]; break; case 6: wcur = gb[6]; break; case 7: wcur = gb[7]; break;
                case 8: wcur = gb[8]; break; case 9: wcur = gb[9]; break; case 10: wcur = gb[10]; break; case 11: wcur = gb[11]; break;
                case 12: wcur = gb[12]; break; case 13: wcur = gb[13]; break; case 14: wcur = gb[14]; break; default: wcur = gb[15]; break;
                }
                const uint32_t wnew = (j & 1) ? ((wcur & 0x0000ffffu) | (hb << 16)) : ((wcur & 0xffff0000u) | hb);
                VMGYM_PUT16(gb, k, wnew);
            }
        }
        if (stage_addr) {
            // coalesced store: the warp's 32 rows x 64 B go through its 2 KB staging buffer (16-byte pieces XOR-swizzled: no bank
            // conflicts either way) and leave as 8 rows x 64 contiguous bytes per store instruction — full 32-byte sectors.  One row
            // per thread (32 rows x 16 B, 67 KB apart) wrote half sectors: 0.7 ms of a 2.0 ms call (profiles/r2_fused_head.md)
            const uint32_t wr = stage_addr + (uint32_t)lane * 64u, sw = (uint32_t)(lane >> 1) & 3u;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wr + ((0u ^ sw) << 4)), "r"(gb[0]), "r"(gb[1]), "r"(gb[2]), "r"(gb[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wr + ((1u ^ sw) << 4)), "r"(gb[4]), "r"(gb[5]), "r"(gb[6]), "r"(gb[7]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wr + ((2u ^ sw) << 4)), "r"(gb[8]), "r"(gb[9]), "r"(gb[10]), "r"(gb[11]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wr + ((3u ^ sw) << 4)), "r"(gb[12]), "r"(gb[13]), "r"(gb[14]), "r"(gb[15]) : "memory");
            __syncwarp();
            const int pc = lane & 3;
            const bool pc_ok = c0 + pc * 8 < fo.apad;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int row = i * 8 + (lane >> 2);
                uint4 w;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
                             : "r"(stage_addr + (uint32_t)row * 64u + ((((uint32_t)pc) ^ ((uint32_t)(row >> 1) & 3u)) << 4)));
                const int er = m0 + q * 32 + row;
                if (er < M && pc_ok)
                    *reinterpret_cast<uint4*>(fo.g_out + (long long)er * fo.ldg + (long long)v * fo.apad + c0 + pc * 8) = w;
            }
            __syncwarp();
        } else if (live) {
            uint4* dst = reinterpret_cast<uint4*>(grow + c0);
            const int n8 = min(4, (fo.apad - c0) >> 3);                  // 16-byte pieces of this chunk inside the VM's apad columns
            dst[0] = make_uint4(gb[0], gb[1], gb[2], gb[3]);
            if (n8 > 1) dst[1] = make_uint4(gb[4], gb[5], gb[6], gb[7]);
            if (n8 > 2) dst[2] = make_uint4(gb[8], gb[9], gb[10], gb[11]);
            if (n8 > 3) dst[3] = make_uint4(gb[12], gb[13], gb[14], gb[15]);
        }
    }
}

__global__ void __launch_bounds__(THREADS, 1) policy_fused_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                   const __grid_constant__ CUtensorMap map_w, FusedOut fo, int M, int K)
{
    extern __shared__ unsigned char smem_raw[];
    const Pipe pipe = pipe_carve(smem_raw);
    float* s_bias = reinterpret_cast<float*>(pipe.tmem_ptr + 4);     // 128 floats (inside the 1 KiB tail region)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int v = blockIdx.x, m0 = blockIdx.y * BM, n0 = v * BN;
    const int k_blocks = (K + BK - 1) / BK;
    pipe_setup(pipe, &map_a, &map_w, warp, lane);
    if (warp >= 2) { const int t = threadIdx.x - 64; s_bias[t] = fo.bias_pad[n0 + t]; }
    const uint32_t tmem_base = pipe_sync_tmem_base(pipe);

    if (warp == 0) {
        if (lane == 0) pipe_produce(pipe, &map_a, &map_w, m0, n0, k_blocks);
    } else if (warp == 1) {
        if (lane == 0) pipe_mma(pipe, tmem_base, k_blocks);
    } else {
        // ===== fused epilogue: thread = env row, registers = this VM's logits =====
        const RowIn in = fused_epilogue_prefetch(fo, warp & 3, lane, m0, v, M);
        mbar_wait(pipe.tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (fo.g_out) eval_epilogue_row(fo, s_bias, tmem_base, in, warp & 3, lane, m0, v, M, 0u);
        else fused_epilogue_row(fo, s_bias, tmem_base, in, warp & 3, lane, m0, v, M);
    }
    pipe_teardown(tmem_base, warp);
}

// ---------------------------------------------------------------------------------------------------
// Persistent fused actor head (K <= 512): one CTA per SM walks "units" = (env tile, chunk of VMs).
//   * the 128 x K activation tile of the unit stays RESIDENT in shared memory (8 K-slices of 16 KB, loaded once per unit):
//     only the W_v tiles stream through a 4-stage ring, which halves the L2 -> SMEM operand traffic of the tile-per-CTA kernel;
//   * P_GROUPS accumulators in TMEM (128 columns each) and as many epilogue warpgroups: while group g samples tile t out
//     of accumulator g, the MMA warp already fills the other accumulators with tiles t+1, ...
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2+4g .. 5+4g epilogue group g
// (a warp may touch TMEM lanes 32 (warp % 4) .. +31, so both groups cover all 128 lanes).
// ---------------------------------------------------------------------------------------------------
constexpr int P_GROUPS = 4;                                    // epilogue warpgroups = TMEM accumulators in flight
constexpr int P_THREADS = 64 + 128 * P_GROUPS + 32;           // + the second MMA issuer (last warp)
constexpr int P_TMEM_COLS = P_GROUPS <= 2 ? 256 : 512;         // power of two >= P_GROUPS * 128
static_assert(P_TMEM_COLS == 512, "the MMA warp assumes the whole TMEM (allocation base 0)");
constexpr int P_WSTAGES = 4;
constexpr int P_KSLICES = 8;                                   // K <= 512
constexpr int P_SLICE_BYTES = BM * BK * 2;                     // 16 KiB: one K-slice of a 128-row operand tile
constexpr int P_VCHUNK = 10;                                   // VM tiles per unit
constexpr int P_STAGE_BYTES = 2048;                            // per epilogue warp: 32 rows x 64 B of logit gradients on their way out
constexpr size_t P_SMEM_BYTES = (size_t)(P_KSLICES + P_WSTAGES) * P_SLICE_BYTES + (size_t)P_GROUPS * 4 * P_STAGE_BYTES +
                                1024 /* alignment */ + 2048 /* barriers, bias per group */;

__global__ void __launch_bounds__(P_THREADS, 1) policy_fused_persistent_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                                const __grid_constant__ CUtensorMap map_w, FusedOut fo,
                                                                                int M, int K)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* s_a = smem;                                              // P_KSLICES slices, resident per unit
    unsigned char* s_w = smem + (size_t)P_KSLICES * P_SLICE_BYTES;          // ring of W K-slices
    unsigned char* s_stage = s_w + (size_t)P_WSTAGES * P_SLICE_BYTES;       // P_STAGE_BYTES per epilogue warp
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + (size_t)P_GROUPS * 4 * P_STAGE_BYTES);
    uint64_t* w_full = bars;                 // [P_WSTAGES]
    uint64_t* w_empty = bars + P_WSTAGES;    // [P_WSTAGES]
    uint64_t* a_full = bars + 2 * P_WSTAGES; // [1]
    uint64_t* a_empty = a_full + 1;          // [1]
    uint64_t* acc_full = a_empty + 1;        // [P_GROUPS]
    uint64_t* acc_empty = acc_full + P_GROUPS;   // [P_GROUPS]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + P_GROUPS);
    float* s_bias = reinterpret_cast<float*>(tmem_ptr + 4);                 // [P_GROUPS][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k_blocks = (K + BK - 1) / BK;
    const int m_tiles = (M + BM - 1) / BM;
    const int v_chunks = (fo.V + P_VCHUNK - 1) / P_VCHUNK;
    const int n_units = m_tiles * v_chunks;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < P_WSTAGES; s++) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 2); }
        mbar_init(a_full, 1); mbar_init(a_empty, 2);                  // both MMA issuers release the unit's activation tile
        for (int g = 0; g < P_GROUPS; g++) { mbar_init(&acc_full[g], 1); mbar_init(&acc_empty[g], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(P_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t wcount = 0, ucount = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x, ucount++) {
                const int vc = u / m_tiles, mt = u % m_tiles;       // CTAs running side by side share the VM chunk's W tiles (L2)
                const int v0 = vc * P_VCHUNK, v1 = min(fo.V, v0 + P_VCHUNK);
                mbar_wait(a_empty, (ucount & 1u) ^ 1u);                        // previous unit's MMAs have read A
                mbar_expect_tx(a_full, (uint32_t)(k_blocks * P_SLICE_BYTES));
                for (int kb = 0; kb < k_blocks; kb++) tma_load_2d(s_a + (size_t)kb * P_SLICE_BYTES, &map_a, kb * BK, mt * BM, a_full);
                for (int v = v0; v < v1; v++) {
                    for (int kb = 0; kb < k_blocks; kb++, wcount++) {
                        const int s = wcount % P_WSTAGES;
                        mbar_wait(&w_empty[s], ((wcount / P_WSTAGES) & 1u) ^ 1u);
                        mbar_expect_tx(&w_full[s], (uint32_t)(fo.apad * BK * 2));
                        tma_load_2d(s_w + (size_t)s * P_SLICE_BYTES, &map_w, kb * BK, v * fo.apad, &w_full[s]);
                    }
                }
            }
        }
    } else if (warp == 1 || warp == P_THREADS / 32 - 1) {
        // ===== MMA issuers: two warps on different schedulers, issuer i takes the tiles with tile index % 2 == i.  A 128 x apad x 64
        // K block is ~224 tensor cycles but ~60 dependent SASS instructions to issue (barrier wait, descriptor adds, commit) for a
        // warp that shares its scheduler with four epilogue warps: one issuer needed ~560 cycles per K block and was the bottleneck
        // of the kernel (profiles/r2_fused_head.md).  The whole warp walks the loop (uniform control flow keeps the address arithmetic
        // on the uniform datapath), one elected lane issues.  All 512 TMEM columns are this CTA's, so the allocation starts at
        // column 0 and the accumulator addresses are plain loop arithmetic (checked once below). =====
        if (tmem_base != 0u) __trap();
        const uint32_t me = warp == 1 ? 0u : 1u;
        const uint32_t idesc = umma_idesc_bf16_f32(BM, fo.apad);         // N = the VM's (padded) row count: multiple of 16
        const uint32_t a_lo0 = umma_desc_lo(smem_u32(s_a)), w_lo0 = umma_desc_lo(smem_u32(s_w));
        uint32_t ws = 0, wphase = 0, ucount = 0, tcount = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ucount++) {
            const int vc = u / m_tiles;
            const int v0 = vc * P_VCHUNK, v1 = min(fo.V, v0 + P_VCHUNK);
            mbar_wait(a_full, ucount & 1u);
            for (int v = v0; v < v1; v++, tcount++) {
                if ((tcount & 1u) != me) {
                    // the other issuer's tile.  A parity wait only works for a waiter that sees EVERY phase of its barrier, so this
                    // warp still observes each ring slot fill and is one of the slot's two releases (w_empty counts 2 arrivals)
                    for (int kb = 0; kb < k_blocks; kb++) {
                        mbar_wait(&w_full[ws], wphase);
                        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&w_empty[ws])) : "memory");
                        __syncwarp();
                        if (++ws == P_WSTAGES) { ws = 0; wphase ^= 1u; }
                    }
                    continue;
                }
                const uint32_t g = tcount % P_GROUPS;
                mbar_wait(&acc_empty[g], ((tcount / P_GROUPS) & 1u) ^ 1u);     // epilogue group g has drained its accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = g * TMEM_COLS;
#pragma unroll 1
                for (int kb = 0; kb < k_blocks; kb++) {
                    mbar_wait(&w_full[ws], wphase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        umma_bf16_k64(acc, a_lo0 + (uint32_t)kb * (P_SLICE_BYTES >> 4), w_lo0 + ws * (P_SLICE_BYTES >> 4), idesc, kb ? 1u : 0u);
                        umma_commit(&w_empty[ws]);
                        if (kb == k_blocks - 1) umma_commit(&acc_full[g]);
                    }
                    __syncwarp();
                    if (++ws == P_WSTAGES) { ws = 0; wphase ^= 1u; }
                }
            }
            if (elect_one()) umma_commit(a_empty);                             // this issuer's MMAs on the unit's A have retired
            __syncwarp();
        }
    } else {
        // ===== epilogue groups: group g = warps 2+4g .. 5+4g handles the tiles with tile index % P_GROUPS == g =====
        const int g = (warp - 2) >> 2;
        const int q = warp & 3;
        const int tg = threadIdx.x - 64 - g * 128;                             // 0..127 inside the group
        float* bias_g = s_bias + g * 128;
        const uint32_t stage_w = smem_u32(s_stage + (size_t)(warp - 2) * P_STAGE_BYTES);
        uint32_t tcount = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int vc = u / m_tiles, mt = u % m_tiles;
            const int v0 = vc * P_VCHUNK, v1 = min(fo.V, v0 + P_VCHUNK);
            for (int v = v0; v < v1; v++, tcount++) {
                if (tcount % P_GROUPS != (uint32_t)g) continue;
                asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");    // previous tile's readers of bias_g are done
                bias_g[tg] = tg < fo.apad ? fo.bias_pad[v * fo.apad + tg] : 0.f;
                asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
                const RowIn in = fused_epilogue_prefetch(fo, q, lane, mt * BM, v, M);
                mbar_wait(&acc_full[g], (tcount / P_GROUPS) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (fo.g_out) eval_epilogue_row(fo, bias_g, tmem_base + g * TMEM_COLS, in, q, lane, mt * BM, v, M, stage_w);
                else fused_epilogue_row(fo, bias_g, tmem_base + g * TMEM_COLS, in, q, lane, mt * BM, v, M);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[g])) : "memory");
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(P_TMEM_COLS) : "memory");
}



// cudaFuncSetAttribute is per device: remember which devices have been configured for a kernel (slot 0..2)
static bool attr_done(int slot, bool mark)
{
    static bool done[3][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (mark) done[slot][dev] = true;
    return done[slot][dev];
}

}  // namespace vmgym_gemm

extern "C" int vmgym_linear_bf16(const void* d_a_bf16, const void* d_w_bf16, const float* d_bias, float* d_c, int64_t M, int64_t N,
                                 int64_t K, int64_t ldc, void* stream)
{
    using namespace vmgym_gemm;
    if (!d_a_bf16 || !d_w_bf16 || !d_c || M < 0 || N < 0 || K <= 0) { vmgym_internal_set_error("vmgym_linear_bf16: null operand"); return VMGYM_EINVAL; }
    if (M == 0 || N == 0) return VMGYM_OK;
    if (K % 8 != 0 || ((uintptr_t)d_a_bf16 & 15) || ((uintptr_t)d_w_bf16 & 15)) {
        vmgym_internal_set_error("vmgym_linear_bf16: K must be a multiple of 8 and operands 16-byte aligned (TMA)");
        return VMGYM_EINVAL;
    }
    CUtensorMap map_a, map_w;
    if (make_map(&map_a, d_a_bf16, (int)M, (int)K, BM) || make_map(&map_w, d_w_bf16, (int)N, (int)K, BN)) {
        vmgym_internal_set_error("vmgym_linear_bf16: cuTensorMapEncodeTiled failed");
        return VMGYM_ECUDA;
    }
    if (!attr_done(0, false)) {
        cudaError_t e = cudaFuncSetAttribute(linear_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) { vmgym_internal_set_error(cudaGetErrorString(e)); return VMGYM_ECUDA; }
        attr_done(0, true);
    }
    dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM));
    linear_bf16_kernel<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(map_a, map_w, d_bias, d_c, (int)M, (int)N, (int)K, ldc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { vmgym_internal_set_error(cudaGetErrorString(e)); return VMGYM_ECUDA; }
    return VMGYM_OK;
}

// rows per VM of the padded output-layer weights / bias (and columns per VM of the logit gradients): the persistent kernel (K <= 512)
// issues MMAs of N = align16(A) columns, the tile-per-CTA fallback keeps full 128-column tiles
extern "C" int vmgym_policy_fused_rows(int64_t A, int64_t K)
{
    using namespace vmgym_gemm;
    static const int persistent = getenv("VMGYM_FUSED_PERSISTENT") ? atoi(getenv("VMGYM_FUSED_PERSISTENT")) : 1;
    if (A < 1 || A > 128) return 0;
    return (persistent && K <= P_KSLICES * BK) ? (int)((A + 15) / 16 * 16) : BN;
}

static int launch_policy_fused(const char* who, const void* d_h_bf16, const void* d_wpad_bf16, vmgym_gemm::FusedOut fo, int64_t M, int64_t V,
                               int64_t A, int64_t K, void* stream)
{
    using namespace vmgym_gemm;
    char msg[160];
    if (A < 1 || A > 128 || K % 8 != 0 || ((uintptr_t)d_h_bf16 & 15) || ((uintptr_t)d_wpad_bf16 & 15) || ((uintptr_t)fo.mask_bits & 15)) {
        snprintf(msg, sizeof(msg), "%s: needs action_dim <= 128, K %% 8 == 0 and 16-byte aligned operands", who);
        vmgym_internal_set_error(msg);
        return VMGYM_EUNSUPPORTED;
    }
    if (M == 0 || V == 0) return VMGYM_OK;
    fo.apad = vmgym_policy_fused_rows(A, K);
    CUtensorMap map_a, map_w;
    if (make_map(&map_a, d_h_bf16, (int)M, (int)K, BM) || make_map(&map_w, d_wpad_bf16, (int)(V * fo.apad), (int)K, fo.apad)) {
        snprintf(msg, sizeof(msg), "%s: cuTensorMapEncodeTiled failed", who);
        vmgym_internal_set_error(msg);
        return VMGYM_ECUDA;
    }
    if (!attr_done(1, false)) {
        cudaError_t e = cudaFuncSetAttribute(policy_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) { vmgym_internal_set_error(cudaGetErrorString(e)); return VMGYM_ECUDA; }
        attr_done(1, true);
    }
    // default: the persistent kernel whenever the activation tile fits (K <= 512); VMGYM_FUSED_PERSISTENT=0 selects the
    // tile-per-CTA kernel (A/B experiments, and the fallback for wider hidden layers)
    static const int persistent = getenv("VMGYM_FUSED_PERSISTENT") ? atoi(getenv("VMGYM_FUSED_PERSISTENT")) : 1;
    if (persistent && K <= P_KSLICES * BK) {
        static int n_sm = 0;
        if (!attr_done(2, false)) {
            cudaError_t e2 = cudaFuncSetAttribute(policy_fused_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM_BYTES);
            if (e2 != cudaSuccess) { vmgym_internal_set_error(cudaGetErrorString(e2)); return VMGYM_ECUDA; }
            attr_done(2, true);
        }
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm <= 0) n_sm = 148;
        const long long units = ((M + BM - 1) / BM) * ((V + P_VCHUNK - 1) / P_VCHUNK);
        const unsigned ctas = (unsigned)(units < n_sm ? units : n_sm);
        policy_fused_persistent_kernel<<<ctas, P_THREADS, P_SMEM_BYTES, (cudaStream_t)stream>>>(map_a, map_w, fo, (int)M, (int)K);
        cudaError_t e3 = cudaGetLastError();
        if (e3 != cudaSuccess) { vmgym_internal_set_error(cudaGetErrorString(e3)); return VMGYM_ECUDA; }
        return VMGYM_OK;
    }
    dim3 grid((unsigned)V, (unsigned)((M + BM - 1) / BM));
    policy_fused_kernel<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(map_a, map_w, fo, (int)M, (int)K);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { vmgym_internal_set_error(cudaGetErrorString(e)); return VMGYM_ECUDA; }
    return VMGYM_OK;
}

extern "C" int vmgym_policy_fused(const void* d_h_bf16, const void* d_wpad_bf16, const float* d_bias_pad, const uint32_t* d_mask_bits,
                                  const void* d_action_in, int64_t M, int64_t V, int64_t A, int64_t K, uint64_t seed,
                                  uint64_t counter, uint8_t* d_action_out, float* d_logprob, float* d_entropy, void* stream)
{
    using namespace vmgym_gemm;
    if (!d_h_bf16 || !d_wpad_bf16 || !d_bias_pad || !d_logprob || !d_entropy || (!d_action_in && !d_action_out) || M < 0) {
        vmgym_internal_set_error("vmgym_policy_fused: null operand");
        return VMGYM_EINVAL;
    }
    FusedOut fo;
    fo.mask_bits = d_mask_bits; fo.bias_pad = d_bias_pad; fo.action_in = d_action_in; fo.action_out = d_action_out;
    fo.logprob = d_logprob; fo.entropy = d_entropy; fo.A = (int)A; fo.V = (int)V; fo.seed = seed; fo.counter = (uint32_t)counter;
    fo.c_logprob = nullptr; fo.c_entropy = 0.f; fo.g_out = nullptr; fo.ldg = 0; fo.stat_m = nullptr; fo.stat_s = nullptr;
    return launch_policy_fused("vmgym_policy_fused", d_h_bf16, d_wpad_bf16, fo, M, V, A, K, stream);
}

extern "C" int vmgym_policy_fused_eval(const void* d_h_bf16, const void* d_wpad_bf16, const float* d_bias_pad, const uint32_t* d_mask_bits,
                                       const void* d_action_in, int64_t M, int64_t V, int64_t A, int64_t K, float* d_logprob, float* d_entropy,
                                       float* d_stat_max, float* d_stat_sum, void* stream)
{
    using namespace vmgym_gemm;
    if (!d_h_bf16 || !d_wpad_bf16 || !d_bias_pad || !d_action_in || !d_logprob || !d_entropy || !d_stat_max || !d_stat_sum || M < 0) {
        vmgym_internal_set_error("vmgym_policy_fused_eval: null operand");
        return VMGYM_EINVAL;
    }
    FusedOut fo;
    fo.mask_bits = d_mask_bits; fo.bias_pad = d_bias_pad; fo.action_in = d_action_in; fo.action_out = nullptr;
    fo.logprob = d_logprob; fo.entropy = d_entropy; fo.A = (int)A; fo.V = (int)V; fo.seed = 0; fo.counter = 0;
    fo.c_logprob = nullptr; fo.c_entropy = 0.f; fo.g_out = nullptr; fo.ldg = 0; fo.stat_m = d_stat_max; fo.stat_s = d_stat_sum;
    return launch_policy_fused("vmgym_policy_fused_eval", d_h_bf16, d_wpad_bf16, fo, M, V, A, K, stream);
}

extern "C" int vmgym_policy_fused_grad(const void* d_h_bf16, const void* d_wpad_bf16, const float* d_bias_pad, const uint32_t* d_mask_bits,
                                       const void* d_action_in, int64_t M, int64_t V, int64_t A, int64_t K, const float* d_c_logprob,
                                       float c_entropy, const float* d_entropy, const float* d_stat_max, const float* d_stat_sum,
                                       void* d_g_bf16, int64_t ldg, void* stream)
{
    using namespace vmgym_gemm;
    if (!d_h_bf16 || !d_wpad_bf16 || !d_bias_pad || !d_action_in || !d_c_logprob || !d_entropy || !d_stat_max || !d_stat_sum || !d_g_bf16 ||
        M < 0 || ldg < V * vmgym_policy_fused_rows(A, K) || (ldg & 7) || ((uintptr_t)d_g_bf16 & 15)) {
        vmgym_internal_set_error("vmgym_policy_fused_grad: null operand, or ldg < V * vmgym_policy_fused_rows(A, K) / not a multiple of 8, or unaligned output");
        return VMGYM_EINVAL;
    }
    FusedOut fo;
    fo.mask_bits = d_mask_bits; fo.bias_pad = d_bias_pad; fo.action_in = d_action_in; fo.action_out = nullptr;
    fo.logprob = nullptr; fo.entropy = const_cast<float*>(d_entropy); fo.A = (int)A; fo.V = (int)V; fo.seed = 0; fo.counter = 0;
    fo.c_logprob = d_c_logprob; fo.c_entropy = c_entropy; fo.g_out = (__nv_bfloat16*)d_g_bf16; fo.ldg = ldg;
    fo.stat_m = const_cast<float*>(d_stat_max); fo.stat_s = const_cast<float*>(d_stat_sum);
    return launch_policy_fused("vmgym_policy_fused_grad", d_h_bf16, d_wpad_bf16, fo, M, V, A, K, stream);
}
