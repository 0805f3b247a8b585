// vmgym_policy.cu — PPO-side kernels of the hot path (sm_100a): masked multi-categorical heads on the actor's
// logits (mask built on the fly from the env records, migration-ratio gating, Gumbel-max sampling with Philox,
// log-prob and entropy), their backward, and the GAE reverse scan.
//
// Reference: src/agents/ppo.py — Network.get_action (:115-126), PPOAgent.act gating (:151-155), update GAE (:237-242);
// vmenv/envs/env.py:45-53 (get_invalid_action_mask), :35-42 (validate).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "vmgym_env_kernels.cuh"
#include "vmgym_sample.cuh"

namespace vmgym {

}  // namespace vmgym
extern "C" void vmgym_internal_set_error(const char* msg);   // vmgym_env.cu: the message vmgym_last_error() returns
namespace vmgym {
static int pfail(int code, const char* msg)
{
    vmgym_internal_set_error(msg);
    return code;
}

struct HeadParams {
    DevLayout L;
    const unsigned char* state;       // env records (mask built on the fly) or nullptr
    const uint32_t* mask_in;          // packed invalid bits [n, V, W] or nullptr
    const float* logits;              // [n, V*A]
    long long n_envs;
    const void* action_in;            // evaluate these actions (update pass) or nullptr (sample)
    int action_dtype;
    float migration_ratio;            // >= 0: PPOAgent.act gating (ppo.py:153-155); < 0: none (training rollouts, ppo.py:196)
    unsigned long long seed, counter; // Philox key / per-call counter
    void* action_out;                 // [n, V] placement dtype
    float* logprob;                   // [n]   sum over VMs (ppo.py:126)
    float* entropy;                   // [n]
    uint32_t* mask_out;               // packed effective mask [n, V, W] or nullptr
    // backward
    const float* g_logprob;           // [n]
    const float* g_entropy;           // [n]
    float* g_logits;                  // [n, V*A]
    int masked;                       // 0: get_invalid_action_mask(masked=False) -> nothing is masked
};

constexpr float MASK_LOGIT = -1e7f;   // ppo.py:119

// Row v of the invalid-action mask for columns a = lane + 32 i, as a bit per i (1 = invalid), built from the env
// record exactly like env.py:45-53 / validate :35-42 (fp64 capacity checks against the current state).
template <typename PT>
__device__ __forceinline__ unsigned row_invalid_bits(const unsigned char* rec, const DevLayout& L, const double* s_cpu,
                                                     const double* s_mem, int v, int lane, int A)
{
    const PT* place = reinterpret_cast<const PT*>(rec + L.off_place);
    const int P = L.P;
    const int cur = (int)place[v];
    double vc = 0.0, vm = 0.0;
    if (cur == P) {
        vc = (double)(rec[L.off_cpuc + v] & 0x7f) / 100.0;
        vm = (double)rec[L.off_memc + v] / 100.0;
    }
    unsigned bits = 0;
    for (int i = 0; i * 32 < A; i++) {
        const int a = i * 32 + lane;
        if (a >= A) break;
        bool valid;
        if (a == cur) valid = true;
        else if (cur == P) valid = a < P && (s_cpu[a] + vc <= 1.0) && (s_mem[a] + vm <= 1.0);
        else if (cur < P) valid = (a == P);
        else valid = false;
        bits |= valid ? 0u : (1u << i);
    }
    return bits;
}

// One CTA per env; each warp walks rows v = warp, warp + W, ...  Per row: mask (+ gating), log-softmax, entropy,
// Gumbel-max sample or evaluation of a given action.  Per-env sums are reduced in a fixed order (deterministic).
// NIC: compile-time number of 32-column groups per row (0 = run time, <= 8) so that the per-lane arrays stay in registers
template <typename PT, bool BACKWARD, int NIC>
__global__ void __launch_bounds__(256) heads_kernel(const HeadParams p)
{
    extern __shared__ __align__(16) unsigned char hs[];
    const DevLayout& L = p.L;
    const int P = L.P, V = L.V, A = L.A;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const long long env = blockIdx.x;
    double* s_cpu = reinterpret_cast<double*>(hs);
    double* s_mem = s_cpu + L.Pp;
    float* s_part = reinterpret_cast<float*>(s_mem + L.Pp);      // [2][nwarps] partial sums
    const unsigned char* rec = p.state ? p.state + env * (long long)L.rec_bytes : nullptr;
    if (rec) {
        const double* g_cpu = reinterpret_cast<const double*>(rec);
        const double* g_mem = reinterpret_cast<const double*>(rec + L.off_mem);
        for (int q = threadIdx.x; q < P; q += blockDim.x) { s_cpu[q] = g_cpu[q]; s_mem[q] = g_mem[q]; }
    }
    __syncthreads();
    const int W = (A + 31) / 32;                  // mask words per row
    const int NI = NIC ? NIC : W;                 // columns per lane
    float lp_sum = 0.f, ent_sum = 0.f;
    const float glp = BACKWARD ? p.g_logprob[env] : 0.f, gent = BACKWARD ? p.g_entropy[env] : 0.f;
    for (int v = warp; v < V; v += nwarps) {
        // ---- mask bits for this lane's columns ----
        unsigned inv = 0;
        if (p.masked) {
            if (rec) inv = row_invalid_bits<PT>(rec, L, s_cpu, s_mem, v, lane, A);
            else if (p.mask_in) {
                const uint32_t* mw = p.mask_in + (env * V + v) * (long long)W;
                _Pragma("unroll") for (int i = 0; i < 8; i++) if (i < NI) inv |= ((mw[i] >> lane) & 1u) << i;
            }
        }
        if (p.migration_ratio >= 0.f && rec) {
            // ppo.py:153-155: if count_nonzero(invalid_row) > 1 and not invalid_row[P] and rand() > migration_ratio:
            //                     invalid_row[P] = True
            int cnt = __popc(inv);
            cnt = __reduce_add_sync(FULL, cnt);
            const int wl = P & 31, wi = P >> 5;
            const unsigned wait_inv = __shfl_sync(FULL, (inv >> wi) & 1u, wl);
            if (cnt > 1 && !wait_inv) {
                const float u = gate_uniform(v, (uint32_t)env, p.seed, (uint32_t)p.counter);
                if (u > p.migration_ratio && lane == wl) inv |= 1u << wi;
            }
        }
        if (p.mask_out) {
            uint32_t* mo = p.mask_out + (env * V + v) * (long long)W;
            _Pragma("unroll") for (int i = 0; i < 8; i++) if (i < NI) {
                const unsigned word = __ballot_sync(FULL, (inv >> i) & 1u);
                if (lane == 0) mo[i] = word;
            }
        }
        if (!p.logits) continue;                     // mask-only call (the fused policy kernel consumes the packed bits)
        // ---- masked logits, log-softmax ----
        const float* z = p.logits + (env * V + v) * (long long)A;
        float zl[8];
        float mx = -INFINITY;
        _Pragma("unroll") for (int i = 0; i < 8; i++) if (i < NI) {
            const int a = i * 32 + lane;
            float x = -INFINITY;
            if (a < A) x = ((inv >> i) & 1u) ? MASK_LOGIT : z[a];
            zl[i] = x;
            mx = fmaxf(mx, x);
        }
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
        // one exponential per element: e = exp(z - max);  p = e / S;  log p = (z - max) - log S;
        // entropy = -sum p log p = log S - (sum e (z - max)) / S        (masked columns: e = 0, contribute 0)
        float el[8];
        float se = 0.f, sez = 0.f;
        _Pragma("unroll") for (int i = 0; i < 8; i++) if (i < NI) {
            float e = 0.f;
            if (i * 32 + lane < A) {
                const float d = zl[i] - mx;
                e = fast_exp(d);
                sez += e * d;
            }
            el[i] = e;
            se += e;
        }
        for (int o = 16; o > 0; o >>= 1) { se += __shfl_xor_sync(FULL, se, o); sez += __shfl_xor_sync(FULL, sez, o); }
        const float log_se = logf(se), inv_se = 1.0f / se;
        const float lse = mx + log_se;
        // Categorical.entropy = -sum p (z - lse) with the float32 log-sum-exp the reference normalises with (0 contribution from
        // masked columns); (lse - mx) equals log_se except for fully masked rows, where max = -1e7 makes the rounding of lse visible
        const float ent = (lse - mx) - sez * inv_se;
        // ---- action: given or Gumbel-max sample ----
        int act;
        if (p.action_in) {
            const unsigned char* arow = reinterpret_cast<const unsigned char*>(p.action_in) + env * (long long)V * dtype_bytes(p.action_dtype);
            act = load_action(arow, p.action_dtype, v);
        } else {
            // streaming inverse-CDF sampling (vmgym_sample.cuh), the same arithmetic as the fused tcgen05 epilogue: chunk =
            // the 32 columns i*32 .. i*32+31 (one per lane), running max / sum with rescaling, butterfly sums = the fused
            // kernel's tree sums, group and in-group scans in the same sequential order
            float m_run = -1e30f, s_run = 0.f;
            int besta = 0;
            const Philox4 rnd0 = sample_block(v, 0, (uint32_t)env, p.seed, (uint32_t)p.counter);
            const Philox4 rnd1 = NI > 4 ? sample_block(v, 1, (uint32_t)env, p.seed, (uint32_t)p.counter) : rnd0;
            _Pragma("unroll") for (int i = 0; i < 8; i++) if (i < NI) {
                const float x = (i * 32 + lane < A) ? zl[i] : -1e30f;
                float cm = x;
                for (int o = 16; o > 0; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(FULL, cm, o));
                if (cm > m_run) {
                    if (i > 0) s_run *= fast_exp(-(cm - m_run));
                    m_run = cm;
                }
                const float e = fast_exp(x - m_run);
                float sm = e;
                sm += __shfl_xor_sync(FULL, sm, 1);
                sm += __shfl_xor_sync(FULL, sm, 2);
                const float gsum = sm;                                  // sum of this lane's group of 4 columns
                sm += __shfl_xor_sync(FULL, sm, 4);
                sm += __shfl_xor_sync(FULL, sm, 8);
                sm += __shfl_xor_sync(FULL, sm, 16);
                const float w = sm, s_new = s_run + w;
                const float target = chunk_uniform(i < 4 ? rnd0 : rnd1, i) * s_new;
                float cum = 0.f, base = 0.f, blast = 0.f;
                int tsel = -1, tlast = 0;
                for (int t = 0; t < 8; t++) {
                    const float gt = __shfl_sync(FULL, gsum, 4 * t), prev = cum;
                    cum += gt;
                    if (gt > 0.f) { tlast = t; blast = prev; }
                    if (tsel < 0 && cum > target) { tsel = t; base = prev; }
                }
                if (tsel < 0) { tsel = tlast; base = blast; }
                int k = -1, klast = 0;
                float c2 = base;
                for (int kk = 0; kk < 4; kk++) {
                    const float ek = __shfl_sync(FULL, e, 4 * tsel + kk);
                    c2 += ek;
                    if (kk > 0 && ek > 0.f) klast = kk;
                    if (k < 0 && c2 > target) k = kk;
                }
                if (k < 0) k = klast;
                if (target < w) besta = i * 32 + 4 * tsel + k;
                s_run = s_new;
            }
            act = besta;
            if (p.action_out && lane == 0) reinterpret_cast<PT*>(p.action_out)[env * V + v] = (PT)act;
        }
        float lpa = 0.f;
        if ((unsigned)act < (unsigned)A) {
            // z[act] again from memory (a broadcast L1 hit) rather than zl[act >> 5]: a dynamic index would push the
            // per-lane arrays into local memory
            const unsigned inv_a = __shfl_sync(FULL, inv, act & 31);
            lpa = (((inv_a >> (act >> 5)) & 1u) ? MASK_LOGIT : z[act]) - lse;
        }
        lp_sum += lpa;
        ent_sum += ent;
        if (BACKWARD) {
            // d(sum logprob)/dz_j = [j == a] - p_j ; d(sum entropy)/dz_j = -p_j (logp_j + H); masked columns get 0
            float* gz = p.g_logits + (env * V + v) * (long long)A;
            _Pragma("unroll") for (int i = 0; i < 8; i++) if (i < NI) {
                const int a = i * 32 + lane;
                if (a < A) {
                    const float l = zl[i] - lse, pr = el[i] * inv_se;
                    float g = glp * ((a == act ? 1.f : 0.f) - pr) - gent * pr * (l + ent);
                    if ((inv >> i) & 1u) g = 0.f;
                    gz[a] = g;
                }
            }
        }
    }
    if (!BACKWARD) {
        if (lane == 0) { s_part[warp] = lp_sum; s_part[nwarps + warp] = ent_sum; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.f, b = 0.f;
            for (int w = 0; w < nwarps; w++) { a += s_part[w]; b += s_part[nwarps + w]; }
            if (p.logprob) p.logprob[env] = a;
            if (p.entropy) p.entropy[env] = b;
        }
    }
}

// Evaluate-mode heads (the PPO update: log-prob / entropy of STORED actions under a PACKED mask, and their backward)
// with 8 lanes per VM row and 4 rows per warp in flight: a row of A <= 128 logits is 4 x 32-byte segments per load
// instruction, reductions take 3 shuffle steps instead of 5, and four independent rows hide each other's latency.
// Same arithmetic per row as heads_kernel (one exponential per element); per-env sums are reduced in a fixed order.
template <typename PT, bool BACKWARD, int NEL>
__global__ void __launch_bounds__(256) heads_eval_kernel(const HeadParams p)
{
    __shared__ float s_part[2][8];
    const int V = p.L.V, A = p.L.A;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int sub = lane & 7, grp = lane >> 3;
    const long long env = blockIdx.x;
    const int W = (A + 31) / 32;
    float lp_sum = 0.f, ent_sum = 0.f;
    const float glp = BACKWARD ? p.g_logprob[env] : 0.f, gent = BACKWARD ? p.g_entropy[env] : 0.f;
    const unsigned char* arow = reinterpret_cast<const unsigned char*>(p.action_in) + env * (long long)V * dtype_bytes(p.action_dtype);
    for (int v0 = warp * 4; v0 < V; v0 += nwarps * 4) {
        const int v = v0 + grp;
        const bool live = v < V;
        const float* z = p.logits + (env * V + (live ? v : 0)) * (long long)A;
        uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
        if (p.masked && p.mask_in && live) {
            const uint32_t* mw = p.mask_in + (env * V + v) * (long long)W;
            m0 = mw[0];
            if (W > 1) m1 = mw[1];
            if (W > 2) m2 = mw[2];
            if (W > 3) m3 = mw[3];
        }
        float zl[NEL], el[NEL];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < NEL; i++) {
            const int a = sub + 8 * i;
            const uint32_t mwd = (i >> 2) == 0 ? m0 : ((i >> 2) == 1 ? m1 : ((i >> 2) == 2 ? m2 : m3));   // word a >> 5 == i >> 2
            float x = -INFINITY;
            if (a < A && live) x = ((mwd >> (8 * (i & 3) + sub)) & 1u) ? MASK_LOGIT : z[a];
            zl[i] = x;
            mx = fmaxf(mx, x);
        }
        for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
        float se = 0.f, sez = 0.f;
#pragma unroll
        for (int i = 0; i < NEL; i++) {
            float e = 0.f;
            if (sub + 8 * i < A && live) {
                const float d = zl[i] - mx;
                e = fast_exp(d);
                sez += e * d;
            }
            el[i] = e;
            se += e;
        }
        for (int o = 4; o > 0; o >>= 1) { se += __shfl_xor_sync(FULL, se, o); sez += __shfl_xor_sync(FULL, sez, o); }
        if (!live) continue;                           // whole groups only; no shuffles below
        const float log_se = logf(se), inv_se = 1.0f / se;
        const float lse = mx + log_se;
        const float ent = (lse - mx) - sez * inv_se;           // see heads_kernel
        const int act = load_action(arow, p.action_dtype, v);
        float lpa = 0.f;
        if ((unsigned)act < (unsigned)A) {
            const int wa = act >> 5;
            const uint32_t mwa = wa == 0 ? m0 : (wa == 1 ? m1 : (wa == 2 ? m2 : m3));
            lpa = (((mwa >> (act & 31)) & 1u) ? MASK_LOGIT : z[act]) - lse;
        }
        if (sub == 0) { lp_sum += lpa; ent_sum += ent; }
        if (BACKWARD) {
            float* gz = p.g_logits + (env * V + v) * (long long)A;
#pragma unroll
            for (int i = 0; i < NEL; i++) {
                const int a = sub + 8 * i;
                if (a < A) {
                    const uint32_t mwd = (i >> 2) == 0 ? m0 : ((i >> 2) == 1 ? m1 : ((i >> 2) == 2 ? m2 : m3));
                    const float l = zl[i] - lse, pr = el[i] * inv_se;
                    float g = glp * ((a == act ? 1.f : 0.f) - pr) - gent * pr * (l + ent);
                    if ((mwd >> (8 * (i & 3) + sub)) & 1u) g = 0.f;
                    gz[a] = g;
                }
            }
        }
    }
    if (!BACKWARD) {
        // fixed-order reduction: groups of a warp (lanes 0, 8, 16, 24), then warps
        lp_sum += __shfl_xor_sync(FULL, lp_sum, 8);  ent_sum += __shfl_xor_sync(FULL, ent_sum, 8);
        lp_sum += __shfl_xor_sync(FULL, lp_sum, 16); ent_sum += __shfl_xor_sync(FULL, ent_sum, 16);
        if (lane == 0) { s_part[0][warp] = lp_sum; s_part[1][warp] = ent_sum; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.f, b = 0.f;
            for (int w = 0; w < nwarps; w++) { a += s_part[0][w]; b += s_part[1][w]; }
            if (p.logprob) p.logprob[env] = a;
            if (p.entropy) p.entropy[env] = b;
        }
    }
}

// Mask-only variant (the fused policy kernel consumes the packed bits): one CTA per env, one THREAD per VM row.
// The fp64 capacity test  cpu[a] + c/100.0 <= 1.0  (env.py:35-42) is monotone in the size code c, so each PM's largest
// admissible cpu / memory code is found once per env (binary search with the same fp64 expression) and every waiting
// row then reduces to byte-SIMD compares of its two codes against 4 PMs per instruction.  Bit-identical to
// row_invalid_bits; ~7x fewer instructions and no per-row global latency (placement / codes staged in SMEM).
// NW = mask words per row the kernel is compiled for (A <= 32 NW): the per-row loops are fully unrolled over NW words, so a 102-action
// row costs 4 words of work, not 8 (178 -> 131 us at 8192 envs of config/100.yml; ncu: the 8-word code ran 12.2 k warp-instructions per env with 3.4 no_instruction stalls per issue)
template <typename PT, int NW>
__global__ void __launch_bounds__(128) mask_bits_kernel(const HeadParams p)
{
    extern __shared__ __align__(16) unsigned char ms[];
    const DevLayout& L = p.L;
    const int P = L.P, V = L.V, A = L.A;
    const int W = (A + 31) / 32;
    const int P4 = (P + 3) / 4;                                   // packed capacity words
    double* s_tab = reinterpret_cast<double*>(ms);                // k / 100.0, k < 128
    uint32_t* s_kc = reinterpret_cast<uint32_t*>(s_tab + 128);    // [64] max cpu code per PM, 4 PMs per word
    uint32_t* s_km = s_kc + 64;
    uint32_t* s_out = s_km + 64;                                  // [V * W]
    const long long env = blockIdx.x;
    const unsigned char* rec = p.state + env * (long long)L.rec_bytes;
    for (int k = threadIdx.x; k < 128; k += blockDim.x) s_tab[k] = code_to_f64(k);
    for (int q = threadIdx.x; q < 128; q += blockDim.x) s_kc[q] = 0;          // zeroes s_kc and s_km (pad PMs admit nothing)
    __syncthreads();
    const double* g_cpu = reinterpret_cast<const double*>(rec);
    const double* g_mem = reinterpret_cast<const double*>(rec + L.off_mem);
    for (int q = threadIdx.x; q < 2 * P; q += blockDim.x) {
        const bool is_mem = q >= P;
        const int pm = is_mem ? q - P : q;
        const double load = is_mem ? g_mem[pm] : g_cpu[pm];
        int lo = -1, hi = 128;                                     // largest k with load + k/100 <= 1 (-1: not even k = 0)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (load + s_tab[mid] <= 1.0) lo = mid; else hi = mid;
        }
        // codes compare as unsigned bytes with c <= k; k = -1 cannot be expressed, but c = 0 never needs it: a load
        // above 1.0 cannot occur (every placement passed this same test), so lo >= 0 always
        reinterpret_cast<unsigned char*>(is_mem ? s_km : s_kc)[pm] = (unsigned char)max(lo, 0);
    }
    __syncthreads();
    const PT* place = reinterpret_cast<const PT*>(rec + L.off_place);
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        const int cur = (int)place[v];
        uint32_t bits[NW];
        #pragma unroll
        for (int w = 0; w < NW; w++) bits[w] = 0xffffffffu;
        if (p.masked) {
            if (cur == P) {
                const uint32_t c4 = (uint32_t)(rec[L.off_cpuc + v] & 0x7f) * 0x01010101u;
                const uint32_t m4 = (uint32_t)rec[L.off_memc + v] * 0x01010101u;
                #pragma unroll
                for (int w = 0; w < NW; w++) {
                    uint32_t word = 0;
                    #pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int g = w * 8 + j;
                        if (g < P4) {
                            const uint32_t ok = __vcmpleu4(c4, s_kc[g]) & __vcmpleu4(m4, s_km[g]);
                            const uint32_t inv4 = ((~ok & 0x08040201u) * 0x01010101u) >> 24;   // one bit per byte lane
                            word |= inv4 << (4 * j);
                        } else word |= 0xfu << (4 * j);
                    }
                    bits[w] = word;
                }
            }
            // columns >= P are invalid unless named below; a == cur is always valid; a placed VM may also go back to P
            #pragma unroll
            for (int w = 0; w < NW; w++) {
                const int base = w * 32;
                if (P - base < 32) bits[w] |= (P - base <= 0) ? 0xffffffffu : (0xffffffffu << (P - base));
                if (A - base < 32) bits[w] &= (A - base <= 0) ? 0u : ~(0xffffffffu << (A - base));
                if ((cur >> 5) == w && cur < A) bits[w] &= ~(1u << (cur & 31));
                if (cur < P && (P >> 5) == w) bits[w] &= ~(1u << (P & 31));
            }
        } else {
            #pragma unroll
            for (int w = 0; w < NW; w++) bits[w] = 0u;
        }
        if (p.migration_ratio >= 0.f) {
            // ppo.py:153-155: if count_nonzero(invalid_row) > 1 and not invalid_row[P] and rand() > migration_ratio
            int cnt = 0;
            #pragma unroll
            for (int w = 0; w < NW; w++) cnt += (w < W) ? __popc(bits[w]) : 0;
            uint32_t wait_inv = 0;
            #pragma unroll
            for (int w = 0; w < NW; w++) if ((P >> 5) == w) wait_inv = (bits[w] >> (P & 31)) & 1u;
            if (cnt > 1 && !wait_inv) {
                const float u = gate_uniform(v, (uint32_t)env, p.seed, (uint32_t)p.counter);
                if (u > p.migration_ratio) {
                    #pragma unroll
                    for (int w = 0; w < NW; w++) if ((P >> 5) == w) bits[w] |= 1u << (P & 31);
                }
            }
        }
        #pragma unroll
        for (int w = 0; w < NW; w++) if (w < W) s_out[v * W + w] = bits[w];
    }
    __syncthreads();
    uint32_t* mo = p.mask_out + env * (long long)V * W;
    for (int i = threadIdx.x; i < V * W; i += blockDim.x) mo[i] = s_out[i];
}

// GAE (ppo.py:237-242) per env along T:  A_t = delta_t + gamma*lambda*(1-done_t) * A_{t+1},
// delta_t = r_t + gamma*(1-done_t)*V(s_{t+1}) - V(s_t);  returns = A + V.   Tensors are [T, N] (time-major).
// One warp per env: the recurrence is a composition of affine maps x -> c*x + d, scanned in reverse with shuffles,
// 32 time steps per pass with a carry.
__global__ void gae_kernel(const float* rewards, const float* values, const float* next_values, const uint8_t* dones, int T,
                           long long N, float gamma, float lambda, float* adv, float* ret)
{
    const int lane = threadIdx.x & 31;
    const long long env = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (env >= N) return;
    float carry = 0.f;                                  // A_{t+1} beyond the current chunk
    for (int t1 = T; t1 > 0; t1 -= 32) {
        const int t = t1 - 1 - lane;                    // lane 0 holds the latest step of the chunk
        float c = 0.f, d = 0.f, v = 0.f;
        if (t >= 0) {
            const long long i = (long long)t * N + env;
            const float nd = 1.f - (float)dones[i];
            v = values[i];
            d = rewards[i] + nd * gamma * next_values[i] - v;
            c = nd * gamma * lambda;
        }
        // inclusive scan over lanes 0..lane of f_lane o ... o f_0 where f(x) = c*x + d, applied to `carry`
        for (int o = 1; o < 32; o <<= 1) {
            const float pc = __shfl_up_sync(FULL, c, o), pd = __shfl_up_sync(FULL, d, o);
            if (lane >= o) { d = c * pd + d; c = c * pc; }
        }
        const float a = c * carry + d;
        if (t >= 0) {
            const long long i = (long long)t * N + env;
            adv[i] = a;
            ret[i] = a + v;
        }
        carry = __shfl_sync(FULL, a, min(31, t1 - 1));  // A at the earliest step of this chunk
    }
}

}  // namespace vmgym

using namespace vmgym;

extern "C" {

static int heads_launch(HeadParams& hp, const vmgym_config* cfg, bool backward, void* stream)
{
    vmgym_layout pub;
    int rc = vmgym_get_layout(cfg, &pub);
    if (rc) return pfail(rc, vmgym_last_error());
    // rebuild the device layout from the public one (only the fields the heads kernel reads)
    DevLayout& L = hp.L;
    L.P = cfg->pms; L.V = cfg->vms; L.A = pub.action_dim; L.Pp = pub.pms_padded; L.Vp = pub.vms_padded; L.D = pub.obs_dim;
    L.off_mem = pub.off_memory; L.off_rem = pub.off_remaining; L.off_place = pub.off_placement; L.off_cpuc = pub.off_cpu_code;
    L.off_memc = pub.off_mem_code; L.off_cap = pub.off_capacity; L.off_scal = pub.off_scalars; L.rec_bytes = pub.record_bytes;
    if (L.A > 256) return pfail(VMGYM_EUNSUPPORTED, "action_dim > 256 not supported by the heads kernel yet");
    if ((!hp.logits && !hp.mask_out) || hp.n_envs < 0) return pfail(VMGYM_EINVAL, "null logits");
    if (hp.n_envs == 0) return VMGYM_OK;
    const int threads = 256;
    const size_t smem = (size_t)16 * L.Pp + 2 * (threads / 32) * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    const bool small = L.P <= 253;
    if (!backward && !hp.logits && hp.state) {
        // mask-only call: thread-per-row kernel
        const size_t msmem = 128 * sizeof(double) + 128 * sizeof(uint32_t) + (size_t)L.V * ((L.A + 31) / 32) * sizeof(uint32_t);
        if (msmem > 200 * 1024) return pfail(VMGYM_EUNSUPPORTED, "vms too large for the mask kernel");
        const int words = (L.A + 31) / 32;
        if (words > 8) return pfail(VMGYM_EUNSUPPORTED, "the mask kernel supports action_dim <= 256");
#define VMGYM_MASK_LAUNCH(PT_, NW_) do {                                                                                              \
            if (msmem > 48 * 1024) cudaFuncSetAttribute(mask_bits_kernel<PT_, NW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem); \
            mask_bits_kernel<PT_, NW_><<<(unsigned)hp.n_envs, 128, msmem, st>>>(hp); } while (0)
#define VMGYM_MASK_WORDS(PT_) do { if (words <= 1) VMGYM_MASK_LAUNCH(PT_, 1); else if (words <= 2) VMGYM_MASK_LAUNCH(PT_, 2);            \
                                   else if (words <= 4) VMGYM_MASK_LAUNCH(PT_, 4); else VMGYM_MASK_LAUNCH(PT_, 8); } while (0)
        if (small) VMGYM_MASK_WORDS(uint8_t); else VMGYM_MASK_WORDS(uint16_t);
#undef VMGYM_MASK_WORDS
#undef VMGYM_MASK_LAUNCH
        cudaError_t merr = cudaGetLastError();
        if (merr != cudaSuccess) return pfail(VMGYM_ECUDA, cudaGetErrorString(merr));
        return VMGYM_OK;
    }
    if (hp.action_in && hp.logits && !hp.state && !hp.mask_out && L.A <= 128 && (!hp.masked || hp.mask_in)) {
        // evaluate mode on a packed mask (the PPO update): 8 lanes per row
#define VMGYM_EVAL_LAUNCH(PT_, BW_) do { if (L.A <= 104) heads_eval_kernel<PT_, BW_, 13><<<(unsigned)hp.n_envs, 256, 0, st>>>(hp); \
                                         else heads_eval_kernel<PT_, BW_, 16><<<(unsigned)hp.n_envs, 256, 0, st>>>(hp); } while (0)
        if (backward) { if (small) VMGYM_EVAL_LAUNCH(uint8_t, true); else VMGYM_EVAL_LAUNCH(uint16_t, true); }
        else { if (small) VMGYM_EVAL_LAUNCH(uint8_t, false); else VMGYM_EVAL_LAUNCH(uint16_t, false); }
#undef VMGYM_EVAL_LAUNCH
        cudaError_t eerr = cudaGetLastError();
        if (eerr != cudaSuccess) return pfail(VMGYM_ECUDA, cudaGetErrorString(eerr));
        return VMGYM_OK;
    }
    const int W = (L.A + 31) / 32;
#define VMGYM_HEADS_LAUNCH(PT_, BW_, NIC_) heads_kernel<PT_, BW_, NIC_><<<(unsigned)hp.n_envs, threads, smem, st>>>(hp)
#define VMGYM_HEADS_PICK(PT_, BW_) do { if (W == 4) VMGYM_HEADS_LAUNCH(PT_, BW_, 4); else if (W == 1) VMGYM_HEADS_LAUNCH(PT_, BW_, 1); \
                                        else VMGYM_HEADS_LAUNCH(PT_, BW_, 0); } while (0)
    if (backward) {
        if (small) VMGYM_HEADS_PICK(uint8_t, true); else VMGYM_HEADS_PICK(uint16_t, true);
    } else {
        if (small) VMGYM_HEADS_PICK(uint8_t, false); else VMGYM_HEADS_PICK(uint16_t, false);
    }
#undef VMGYM_HEADS_PICK
#undef VMGYM_HEADS_LAUNCH
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return pfail(VMGYM_ECUDA, cudaGetErrorString(err));
    return VMGYM_OK;
}

int vmgym_policy_heads(const vmgym_config* cfg, const void* d_state, const uint32_t* d_mask_in, int masked,
                       const float* d_logits, int64_t n_envs, const void* d_action_in, int action_dtype,
                       float migration_ratio, uint64_t seed, uint64_t counter, void* d_action_out, float* d_logprob,
                       float* d_entropy, uint32_t* d_mask_out, void* stream)
{
    HeadParams hp = {};
    hp.state = (const unsigned char*)d_state; hp.mask_in = d_mask_in; hp.masked = masked; hp.logits = d_logits;
    hp.n_envs = n_envs; hp.action_in = d_action_in; hp.action_dtype = action_dtype; hp.migration_ratio = migration_ratio;
    hp.seed = seed; hp.counter = counter; hp.action_out = d_action_out; hp.logprob = d_logprob; hp.entropy = d_entropy;
    hp.mask_out = d_mask_out;
    if (d_logits && !d_action_in && !d_action_out) return pfail(VMGYM_EINVAL, "need action_in (evaluate) or action_out (sample)");
    return heads_launch(hp, cfg, false, stream);
}

int vmgym_policy_heads_backward(const vmgym_config* cfg, const uint32_t* d_mask_in, int masked, const float* d_logits,
                                int64_t n_envs, const void* d_action_in, int action_dtype, const float* d_g_logprob,
                                const float* d_g_entropy, float* d_g_logits, void* stream)
{
    HeadParams hp = {};
    hp.mask_in = d_mask_in; hp.masked = masked; hp.logits = d_logits; hp.n_envs = n_envs; hp.action_in = d_action_in;
    hp.action_dtype = action_dtype; hp.migration_ratio = -1.f; hp.g_logprob = d_g_logprob; hp.g_entropy = d_g_entropy;
    hp.g_logits = d_g_logits;
    if (!d_action_in || !d_g_logprob || !d_g_entropy || !d_g_logits) return pfail(VMGYM_EINVAL, "null backward operand");
    return heads_launch(hp, cfg, true, stream);
}

int vmgym_gae(const float* d_rewards, const float* d_values, const float* d_next_values, const uint8_t* d_dones, int32_t T,
              int64_t n_envs, float gamma, float lambda, float* d_advantages, float* d_returns, void* stream)
{
    if (!d_rewards || !d_values || !d_next_values || !d_dones || !d_advantages || !d_returns || T < 0 || n_envs < 0)
        return pfail(VMGYM_EINVAL, "null GAE operand");
    if (T == 0 || n_envs == 0) return VMGYM_OK;
    const int threads = 256;
    const long long blocks = (n_envs * 32 + threads - 1) / threads;
    gae_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(d_rewards, d_values, d_next_values, d_dones, T, n_envs,
                                                                     gamma, lambda, d_advantages, d_returns);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return pfail(VMGYM_ECUDA, cudaGetErrorString(err));
    return VMGYM_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// DRL-VMP placement heuristics (src/agents/drlvmp.py:549-617): for one waiting VM of each env and a choice in
// {0 worst-fit, 1 min dot-product, 2 min L2 distance, 3 best-fit}, pick a PM from the float32 observation.  The
// reference never updates cpu/memory in the observation copy between VMs (drlvmp.py:557-565), so every call sees
// the original PM loads.  Tie rules: worst-fit / dot / L2 -> lowest PM index (first minimum, torch.argmin; ascending
// stable order), best-fit -> highest PM index among equal keys (torch.flip of a stable ascending argsort).
// ---------------------------------------------------------------------------------------------------
namespace vmgym {

__global__ void drlvmp_choice_kernel(int P, int V, int D, const float* obs, const int* vm_index, const int* choice,
                                     long long n_envs, int* pm_out)
{
    const int lane = threadIdx.x & 31;
    const long long env = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (env >= n_envs) return;
    const int v = vm_index[env];
    if (v < 0 || v >= V) { if (lane == 0) pm_out[env] = -1; return; }
    const float* o = obs + env * (long long)D;
    const float vc = o[V + v], vm = o[2 * V + v];
    const float* cpu = o + 3 * V;
    const float* mem = o + 3 * V + P;
    const int ch = choice[env];
    // lexicographic arg-min of (key, tie) over PMs; `tie` encodes the index order wanted among equal keys
    float bestk = INFINITY;
    int bestt = 0x7fffffff, bestp = -1;
    for (int p = lane; p < P; p += 32) {
        const float c = cpu[p], m = mem[p];
        float key;
        int tie = p;
        bool ok = true;
        if (ch == 0) { key = c + m; ok = (c + vc <= 1.0f) && (m + vm <= 1.0f); }                    // worst-fit: smallest load that fits
        else if (ch == 3) { key = -(c + m); tie = -p; ok = (c + vc <= 1.0f) && (m + vm <= 1.0f); }  // best-fit: largest load that fits
        else if (ch == 1) key = c * vc + m * vm;                                                     // torch.dot([c,m],[vc,vm])
        else { const float dc = c - vc, dm = m - vm; key = sqrtf(dc * dc + dm * dm); }              // torch.norm
        if (ok && (key < bestk || (key == bestk && tie < bestt))) { bestk = key; bestt = tie; bestp = p; }
    }
    for (int o2 = 16; o2 > 0; o2 >>= 1) {
        const float ok_ = __shfl_xor_sync(FULL, bestk, o2);
        const int ot = __shfl_xor_sync(FULL, bestt, o2), op = __shfl_xor_sync(FULL, bestp, o2);
        if (op >= 0 && (bestp < 0 || ok_ < bestk || (ok_ == bestk && ot < bestt))) { bestk = ok_; bestt = ot; bestp = op; }
    }
    if (lane == 0) pm_out[env] = bestp;
}

}  // namespace vmgym

extern "C" int vmgym_drlvmp_choice(const vmgym_config* cfg, const float* d_obs, const int32_t* d_vm_index,
                                   const int32_t* d_choice, int64_t n_envs, int32_t* d_pm_out, void* stream)
{
    if (!cfg || !d_obs || !d_vm_index || !d_choice || !d_pm_out || n_envs < 0) return pfail(VMGYM_EINVAL, "null operand");
    if (n_envs == 0) return VMGYM_OK;
    const int P = cfg->pms, V = cfg->vms, D = 3 * V + 2 * P;
    const int threads = 256;
    const long long blocks = (n_envs * 32 + threads - 1) / threads;
    drlvmp_choice_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(P, V, D, d_obs, d_vm_index, d_choice, n_envs, d_pm_out);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return pfail(VMGYM_ECUDA, cudaGetErrorString(err));
    return VMGYM_OK;
}

// ---------------------------------------------------------------------------------------------------
// One iteration of DRLVMPAgent.act (src/agents/drlvmp.py:504-530) for every env of a batch, everything after the
// head GEMMs in ONE kernel: dueling combination + softmax over the atoms + clamp + expectation (Network.dist/forward,
// :355-372) -> argmax action (:514-515) -> heuristic PM choice on the working observation (:549-617, same arithmetic
// as drlvmp_choice_kernel) -> placement written into the working observation -> rank-1 correction of the feature
// layer's pre-activation for the changed input (the VM's placement entry) and its ReLU for the next iteration's GEMMs.
// The waiting-VM counter k lives on the device (incremented by the caller between launches) so that a captured CUDA
// graph of one iteration can be replayed for every waiting VM.  One warp per env.
// ---------------------------------------------------------------------------------------------------
namespace vmgym {

__global__ void drlvmp_iter_kernel(int P, int V, int D, int H, int n_actions, int atoms, const float* __restrict__ heads, int heads_ld,
                                   const float* __restrict__ support, float* obs, const long long* __restrict__ order,
                                   const long long* __restrict__ n_wait, const long long* __restrict__ kdev, int k_offset,
                                   const float* __restrict__ w_cols, float* pre, float* feat, __nv_bfloat16* feat_split, long long n_envs)
{
    const int lane = threadIdx.x & 31;
    const long long env = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (env >= n_envs) return;
    const long long k = *kdev + k_offset;
    if (!(n_wait[env] > k)) return;                                        // no k-th waiting VM in this env
    const int v = (int)order[env * V + (k < V ? k : V - 1)];
    if (v < 0 || v >= V) return;

    // ---- q-values: softmax_j(val_j + adv_aj - mean_a adv_aj), clamped at 1e-3, dotted with the support ----
    const float* hd = heads + env * (long long)heads_ld;                   // [n_actions][atoms] advantages, then [atoms] values
    // all (4 + 1) x atoms head outputs of this env are loaded up front (<= 4 atoms per lane), the mean over actions once
    constexpr int NA = 4;
    float adv[NA][4], base[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int j = lane + 32 * i;
        const bool in = j < atoms;
#pragma unroll
        for (int b = 0; b < NA; b++) adv[b][i] = in ? hd[b * atoms + j] : 0.f;
        base[i] = in ? hd[NA * atoms + j] : 0.f;
    }
    float sup[4];
#pragma unroll
    for (int i = 0; i < 4; i++) sup[i] = lane + 32 * i < atoms ? support[lane + 32 * i] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float mean = 0.f;
#pragma unroll
        for (int b = 0; b < NA; b++) mean += adv[b][i];
        base[i] -= mean / (float)NA;                                       // val_j - mean_a adv_aj
    }
    int best_a = 0;
    float best_q = -INFINITY;
#pragma unroll
    for (int a = 0; a < NA; a++) {
        float z[4];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            z[i] = -INFINITY;
            if (lane + 32 * i < atoms) { z[i] = base[i] + adv[a][i]; mx = fmaxf(mx, z[i]); }
        }
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (lane + 32 * i < atoms) { z[i] = expf(z[i] - mx); sum += z[i]; }
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (lane + 32 * i < atoms) q += fmaxf(z[i] / sum, 1e-3f) * sup[i];
        }
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(FULL, q, o);
        if (q > best_q) { best_q = q; best_a = a; }                        // first maximum, like torch.argmax
    }
    // heuristic of the chosen action only exists for the reference's four (worst-fit, dot, L2, best-fit)
    const int ch = best_a;

    // ---- heuristic PM choice on the working observation (see drlvmp_choice_kernel) ----
    float* o = obs + env * (long long)D;
    const float vc = o[V + v], vm = o[2 * V + v];
    const float* cpu = o + 3 * V;
    const float* mem = o + 3 * V + P;
    float bestk = INFINITY;
    int bestt = 0x7fffffff, bestp = -1;
    // 8 PMs per lane and pass, all 16 loads issued before the first compare (the scan is latency-, not bandwidth-bound)
    for (int p0 = 0; p0 < P; p0 += 256) {
        float cl[8], ml[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int p = p0 + lane + 32 * i;
            cl[i] = p < P ? cpu[p] : 0.f;
            ml[i] = p < P ? mem[p] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int p = p0 + lane + 32 * i;
            const float c = cl[i], m = ml[i];
            float key;
            int tie = p;
            bool ok = p < P;
            if (ch == 0) { key = c + m; ok = ok && (c + vc <= 1.0f) && (m + vm <= 1.0f); }
            else if (ch == 3) { key = -(c + m); tie = -p; ok = ok && (c + vc <= 1.0f) && (m + vm <= 1.0f); }
            else if (ch == 1) key = c * vc + m * vm;
            else { const float dc = c - vc, dm = m - vm; key = sqrtf(dc * dc + dm * dm); }
            if (ok && (key < bestk || (key == bestk && tie < bestt))) { bestk = key; bestt = tie; bestp = p; }
        }
    }
    for (int o2 = 16; o2 > 0; o2 >>= 1) {
        const float ok_ = __shfl_xor_sync(FULL, bestk, o2);
        const int ot = __shfl_xor_sync(FULL, bestt, o2), op = __shfl_xor_sync(FULL, bestp, o2);
        if (op >= 0 && (bestp < 0 || ok_ < bestk || (ok_ == bestk && ot < bestt))) { bestk = ok_; bestt = ot; bestp = op; }
    }
    if (bestp < 0) return;                                                 // nothing fits: the VM stays WAIT (:527-528)

    // ---- write the placement into the observation copy, correct the pre-activation, refresh its ReLU ----
    const float oldv = o[v], newv = (float)bestp;
    __syncwarp();
    if (lane == 0) o[v] = newv;
    const float delta = newv - oldv;
    const float* __restrict__ wc = w_cols + (long long)v * H;
    float* __restrict__ pr = pre + env * (long long)H;
    float* __restrict__ ft = feat + env * (long long)H;
    for (int h0 = 0; h0 < H; h0 += 256) {
        float a[8], w[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int h = h0 + lane + 32 * i;
            a[i] = h < H ? pr[h] : 0.f;
            w[i] = h < H ? wc[h] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int h = h0 + lane + 32 * i;
            if (h < H) {
                const float x = a[i] + w[i] * delta;
                pr[h] = x;
                const float fx = fmaxf(x, 0.f);
                ft[h] = fx;
                if (feat_split) {
                    // the tensor-core head GEMMs' operand: [hi | lo | hi] bf16 split of relu(pre) (vmgym_cast_split_bf16 layout)
                    __nv_bfloat16* fs = feat_split + env * 3ll * H;
                    const __nv_bfloat16 hi = __float2bfloat16(fx);
                    fs[h] = hi; fs[H + h] = __float2bfloat16(fx - __bfloat162float(hi)); fs[2 * H + h] = hi;
                }
            }
        }
    }
}

}  // namespace vmgym

extern "C" int vmgym_drlvmp_iter(const vmgym_config* cfg, int32_t hidden, int32_t n_actions, int32_t atoms, const float* d_heads,
                                 int32_t heads_ld, const float* d_support, float* d_obs, const int64_t* d_order, const int64_t* d_n_wait,
                                 const int64_t* d_k, int32_t k_offset, const float* d_w_cols, float* d_pre, float* d_feat, void* d_feat_split_bf16,
                                 int64_t n_envs,
                                 void* stream)
{
    if (!cfg || !d_heads || !d_support || !d_obs || !d_order || !d_n_wait || !d_k || !d_w_cols || !d_pre || !d_feat || n_envs < 0)
        return pfail(VMGYM_EINVAL, "null operand");
    if (n_actions != 4) return pfail(VMGYM_EINVAL, "DRL-VMP has exactly four heuristic actions (drlvmp.py:517-530)");
    if (atoms < 1 || atoms > 128 || hidden < 1) return pfail(VMGYM_EINVAL, "atoms must be 1..128");
    if (heads_ld < (n_actions + 1) * atoms) return pfail(VMGYM_EINVAL, "heads_ld smaller than (n_actions + 1) * atoms");
    if (n_envs == 0) return VMGYM_OK;
    const int P = cfg->pms, V = cfg->vms, D = 3 * V + 2 * P;
    const int threads = 256;
    const long long blocks = (n_envs * 32 + threads - 1) / threads;
    vmgym::drlvmp_iter_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        P, V, D, hidden, n_actions, atoms, d_heads, heads_ld, d_support, d_obs, (const long long*)d_order, (const long long*)d_n_wait,
        (const long long*)d_k, k_offset, d_w_cols, d_pre, d_feat, (__nv_bfloat16*)d_feat_split_bf16, n_envs);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return pfail(VMGYM_ECUDA, cudaGetErrorString(err));
    return VMGYM_OK;
}

// ---------------------------------------------------------------------------------------------------
// Prioritized-replay segment trees (src/segment_tree.py:8-142, used by src/agents/drlvmp.py:157-241): array-backed
// binary trees over `capacity` (power of two) leaves holding fp64 values, node i = op(node 2i, node 2i+1) with the
// reference's association, so sums / minima / prefix-sum descents are bit-identical to the Python floats.
// ---------------------------------------------------------------------------------------------------
namespace vmgym {

// __setitem__ for a batch (segment_tree.py:63-71): leaves first, then one tree level per barrier; threads that share
// a parent write the same value.  Duplicated indices resolve to the LAST value of the batch, like sequential stores.
__global__ void segtree_update_kernel(double* sum_tree, double* min_tree, long long capacity, const long long* idx,
                                      const double* val, int n)
{
    const int t = threadIdx.x;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + t;
        long long node = 0;
        bool owner = false;
        if (i < n) {
            node = idx[i] + capacity;
            owner = true;
            for (int j = i + 1; j < n; j++) if (idx[j] == idx[i]) { owner = false; break; }   // a later write wins
            if (owner) { if (sum_tree) sum_tree[node] = val[i]; if (min_tree) min_tree[node] = val[i]; }
        }
        __syncthreads();
        for (node >>= 1; ; node >>= 1) {
            if (i < n && node >= 1) {
                if (sum_tree) sum_tree[node] = sum_tree[2 * node] + sum_tree[2 * node + 1];
                if (min_tree) min_tree[node] = fmin(min_tree[2 * node], min_tree[2 * node + 1]);
            }
            __syncthreads();
            if (__syncthreads_and(node <= 1)) break;
        }
    }
}

// SumSegmentTree.retrieve (segment_tree.py:103-118) for a batch of upper bounds
__global__ void segtree_retrieve_kernel(const double* sum_tree, long long capacity, const double* upper, int n, long long* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ub = upper[i];
    long long node = 1;
    while (node < capacity) {
        const long long left = 2 * node;
        const double l = sum_tree[left];
        if (l > ub) node = left;
        else { ub -= l; node = left + 1; }
    }
    out[i] = node - capacity;
}

}  // namespace vmgym

extern "C" int vmgym_segtree_update(double* d_sum_tree, double* d_min_tree, int64_t capacity, const int64_t* d_idx,
                                    const double* d_val, int32_t n, void* stream)
{
    if ((!d_sum_tree && !d_min_tree) || !d_idx || !d_val || n < 0 || capacity < 1 || (capacity & (capacity - 1)))
        return pfail(VMGYM_EINVAL, "segment tree: capacity must be a positive power of two (segment_tree.py:30-32)");
    if (n == 0) return VMGYM_OK;
    segtree_update_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_sum_tree, d_min_tree, capacity, (const long long*)d_idx, d_val, n);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return pfail(VMGYM_ECUDA, cudaGetErrorString(err));
    return VMGYM_OK;
}

extern "C" int vmgym_segtree_retrieve(const double* d_sum_tree, int64_t capacity, const double* d_upper, int32_t n,
                                      int64_t* d_out, void* stream)
{
    if (!d_sum_tree || !d_upper || !d_out || n < 0 || capacity < 1 || (capacity & (capacity - 1)))
        return pfail(VMGYM_EINVAL, "segment tree: bad operand");
    if (n == 0) return VMGYM_OK;
    segtree_retrieve_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_sum_tree, capacity, d_upper, n, (long long*)d_out);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return pfail(VMGYM_ECUDA, cudaGetErrorString(err));
    return VMGYM_OK;
}

namespace vmgym {

// SegmentTree.operate(start, end) (segment_tree.py:35-62): the recursive range query with its own association order
// (left part `op` right part), so that a range sum is the same fp64 value as the reference's.  `end` is INCLUSIVE here
// (operate() has already done `end -= 1`).  Depth <= log2(capacity); one thread.
__device__ double segtree_range(const double* tree, long long start, long long end, long long node, long long node_start,
                                long long node_end, int is_min)
{
    if (start == node_start && end == node_end) return tree[node];
    const long long mid = (node_start + node_end) / 2;
    if (end <= mid) return segtree_range(tree, start, end, 2 * node, node_start, mid, is_min);
    if (mid + 1 <= start) return segtree_range(tree, start, end, 2 * node + 1, mid + 1, node_end, is_min);
    const double a = segtree_range(tree, start, mid, 2 * node, node_start, mid, is_min);
    const double b = segtree_range(tree, mid + 1, end, 2 * node + 1, mid + 1, node_end, is_min);
    return is_min ? fmin(a, b) : a + b;
}

// PrioritizedReplayBuffer.sample_batch (drlvmp.py:178-241): stratified proportional sampling + importance weights.
//   p_total = sum_tree.sum(0, len - 1)   -> operate() turns that into the INCLUSIVE range [0, len - 2] (segment_tree.py:55-59):
//                                           the newest element is left out of the total, as in the reference
//   upperbound_i = a + (b - a) * u_i,  a = segment * i,  b = segment * (i + 1),  segment = p_total / batch      (:216-224)
//   weight_i = (leaf_i / sum * len)^-beta / (min / sum * len)^-beta                                             (:229-241)
__global__ void per_sample_kernel(const double* sum_tree, const double* min_tree, long long capacity, long long len, int batch,
                                  const double* u, double beta, long long* idx_out, double* w_out)
{
    __shared__ double s_total;
    if (threadIdx.x == 0) {
        long long end = len - 1;
        if (end <= 0) end += capacity;
        end -= 1;
        s_total = segtree_range(sum_tree, 0, end, 1, 0, capacity - 1, 0);
    }
    __syncthreads();
    const double segment = s_total / (double)batch;
    const double sum_all = sum_tree[1], min_all = min_tree[1];
    const double max_weight = pow(min_all / sum_all * (double)len, -beta);
    for (int i = threadIdx.x; i < batch; i += blockDim.x) {
        const double a = segment * (double)i, b = segment * (double)(i + 1);
        double ub = a + (b - a) * u[i];
        long long node = 1;
        while (node < capacity) {                                        // retrieve (segment_tree.py:103-118)
            const long long left = 2 * node;
            const double l = sum_tree[left];
            if (l > ub) node = left;
            else { ub -= l; node = left + 1; }
        }
        idx_out[i] = node - capacity;
        w_out[i] = pow(sum_tree[node] / sum_all * (double)len, -beta) / max_weight;
    }
}

// Categorical-DQN projection (drlvmp.py:676-699), float32 like the reference: for sample i and atom j
//   t_z = clamp(r + (1 - done) * gamma * z_j, v_min, v_max);  b = (t_z - v_min) / delta_z;  l = floor(b), u = ceil(b)
//   proj[l] += p_j * (u - b)   (first index_add_, j ascending)   then   proj[u] += p_j * (b - l)   (second index_add_)
// When b is integral l == u and both weights are 0 — that probability mass is dropped, as in the reference.
// One warp per sample; each output atom gathers its contributions in j order = the order of the reference's sequential
// index_add_ on the CPU (deterministic, no atomics).
__global__ void c51_project_kernel(const float* next_dist, const float* reward, const int* done, const float* support,
                                   float gamma, float v_min, float v_max, float delta_z, int atoms, long long n, float* proj)
{
    extern __shared__ float c51s[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const long long i = (long long)blockIdx.x * wpc + warp;
    if (i >= n) return;
    float* lo_w = c51s + (size_t)warp * 4 * atoms;        // p_j * (u - b)
    float* up_w = lo_w + atoms;                            // p_j * (b - l)
    int* lo_i = reinterpret_cast<int*>(up_w + atoms);
    int* up_i = lo_i + atoms;
    const float r = reward[i];
    const float ndg = (float)(1 - done[i]) * gamma;        // (1 - done) * gamma
    const float* pd = next_dist + i * atoms;
    for (int j = lane; j < atoms; j += 32) {
        float tz = r + ndg * support[j];
        tz = fminf(fmaxf(tz, v_min), v_max);
        const float b = (tz - v_min) / delta_z;
        const float l = floorf(b), u = ceilf(b);
        lo_i[j] = (int)l; up_i[j] = (int)u;
        lo_w[j] = pd[j] * (u - b);
        up_w[j] = pd[j] * (b - l);
    }
    __syncwarp();
    for (int k = lane; k < atoms; k += 32) {
        float acc = 0.f;
        for (int j = 0; j < atoms; j++) if (lo_i[j] == k) acc += lo_w[j];
        for (int j = 0; j < atoms; j++) if (up_i[j] == k) acc += up_w[j];
        proj[i * atoms + k] = acc;
    }
}

}  // namespace vmgym

extern "C" int vmgym_per_sample(const double* d_sum_tree, const double* d_min_tree, int64_t capacity, int64_t len, int32_t batch,
                                const double* d_u, double beta, int64_t* d_idx_out, double* d_weight_out, void* stream)
{
    if (!d_sum_tree || !d_min_tree || !d_u || !d_idx_out || !d_weight_out || batch < 0 || len < 1 || capacity < 1 ||
        (capacity & (capacity - 1)) || len > capacity)
        return pfail(VMGYM_EINVAL, "prioritized replay: bad operand");
    if (batch == 0) return VMGYM_OK;
    per_sample_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(d_sum_tree, d_min_tree, capacity, len, batch, d_u, beta,
                                                         (long long*)d_idx_out, d_weight_out);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return pfail(VMGYM_ECUDA, cudaGetErrorString(err));
    return VMGYM_OK;
}

extern "C" int vmgym_c51_project(const float* d_next_dist, const float* d_reward, const int32_t* d_done, const float* d_support,
                                 float gamma, float v_min, float v_max, int32_t atoms, int64_t n, float* d_proj, void* stream)
{
    if (!d_next_dist || !d_reward || !d_done || !d_support || !d_proj || atoms < 2 || atoms > 1024 || n < 0 || !(v_max > v_min))
        return pfail(VMGYM_EINVAL, "c51 projection: bad operand");
    if (n == 0) return VMGYM_OK;
    const int wpc = 4;
    const size_t smem = (size_t)wpc * 4 * atoms * sizeof(float);
    const float delta_z = (float)((double)(v_max - v_min) / (double)(atoms - 1));        // drlvmp.py:672, python float -> float32
    c51_project_kernel<<<(unsigned)((n + wpc - 1) / wpc), wpc * 32, smem, (cudaStream_t)stream>>>(
        d_next_dist, d_reward, d_done, d_support, gamma, v_min, v_max, delta_z, atoms, n, d_proj);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return pfail(VMGYM_ECUDA, cudaGetErrorString(err));
    return VMGYM_OK;
}
