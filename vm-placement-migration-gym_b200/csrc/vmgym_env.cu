// vmgym_env.cu — the batched env hot path as hand-written sm_100a CUDA: reset, fused step,
// fused heuristic-agent step (multi-step resident in shared memory), agent.act on observations,
// observation and invalid-action-mask kernels, and their C ABI (include/vmgym.h).
//
// Semantics follow the reference's vmenv/envs/env.py and src/agents/{firstfit,bestfit}.py line by line
// (cited at each phase); the mapping onto the GPU is new: one warp owns one env, the env record
// (vmgym_layout) is staged HBM -> shared memory by a single bulk-async copy, mutated there, and written
// back by a single bulk-async store.  fp64 PM accumulators are updated in the reference's VM-index order so
// capacity decisions are bit-identical.  Compiled with -fmad=false: no FMA contraction.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "vmgym_device.cuh"

namespace vmgym {

// ---------------------------------------------------------------------------------------------------
// host-side layout
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int g_warps_per_cta = 0;
static int g_use_bulk = 1;

static int fail(int code, const char* fmt, const char* detail = "")
{
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}

__host__ __device__ static inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

static int make_layout(const vmgym_config* c, DevLayout* L, vmgym_layout* pub)
{
    if (!c) return fail(VMGYM_EINVAL, "null config");
    if (c->pms < 1 || c->pms > 65000 || c->vms < 1 || c->vms > 65000) return fail(VMGYM_EINVAL, "pms/vms out of range");
    if (c->reward_function < VMGYM_REWARD_WR || c->reward_function > VMGYM_REWARD_KL)
        return fail(VMGYM_EINVAL, "unknown reward function");   // env.py:155-156 asserts
    DevLayout l;
    l.P = c->pms; l.V = c->vms; l.A = c->allow_null_action ? c->pms + 2 : c->pms + 1;
    l.Pp = align_up(l.P, 2); l.Vp = align_up(l.V, 16); l.D = 3 * l.V + 2 * l.P;
    const int pb = (l.P <= 253) ? 1 : 2;
    l.off_mem = 8 * l.Pp;
    l.off_rem = 16 * l.Pp;
    l.off_place = l.off_rem + 2 * l.Vp;
    l.off_cpuc = l.off_place + pb * l.Vp;
    l.off_memc = l.off_cpuc + l.Vp;
    l.off_scal = l.off_memc + l.Vp;
    l.rec_bytes = align_up(l.off_scal + (int)sizeof(vmgym_env_scalars), 128);
    l.sm_cpu32 = l.rec_bytes;
    l.sm_mem32 = l.sm_cpu32 + align_up(4 * l.P, 16);
    l.sm_act = l.sm_mem32 + align_up(4 * l.P, 16);
    l.sm_tmp = l.sm_act + 2 * l.Vp;
    const int tmp_bytes = align_up((2 * l.Vp > 6 * l.Pp) ? 2 * l.Vp : 6 * l.Pp, 16);
    l.sm_fit = l.sm_tmp + tmp_bytes;                       // u32 fitm[128] | u16 cap[Pp]
    l.sm_bar = l.sm_fit + 512 + align_up(2 * l.Pp, 16);
    l.sm_stride = align_up(l.sm_bar + 16, 128);
    l.sm_tables = SIZE_TABLE * 8 + SIZE_TABLE * 4 + ARR_CDF_SMEM * 8;
    if (L) *L = l;
    if (pub) {
        pub->record_bytes = l.rec_bytes; pub->pms_padded = l.Pp; pub->vms_padded = l.Vp; pub->place_bytes = pb;
        pub->off_cpu = 0; pub->off_memory = l.off_mem; pub->off_remaining = l.off_rem; pub->off_placement = l.off_place;
        pub->off_cpu_code = l.off_cpuc; pub->off_mem_code = l.off_memc; pub->off_scalars = l.off_scal;
        pub->obs_dim = l.D; pub->action_dim = l.A; pub->smem_bytes_per_env = l.sm_stride;
    }
    return VMGYM_OK;
}

struct StepParams {
    DevLayout L;
    int reward_fn, cap_target, step_limit;
    double beta;
    unsigned char* state;
    long long n_envs;
    vmgym_trace tr;
    const void* action;       // [n_envs, V] of action_dtype (external-action mode)
    int action_dtype;
    vmgym_outputs out;
    int agent, tiebreak, n_steps;
    int use_bulk;             // stage records with cp.async.bulk (1) or 128-bit loads/stores (0)
};

// ---------------------------------------------------------------------------------------------------
// per-warp env context (pointers into the warp's shared-memory region)
// ---------------------------------------------------------------------------------------------------
template <typename PT>
struct Env {
    double* cpu; double* mem;          // fp64 PM accumulators (env.py:190-191)
    uint16_t* rem;                     // vm_remaining_runtime
    PT* place;                         // vm_placement
    uint8_t* cpuc; uint8_t* memc;      // size codes (bit 7 of cpuc = suspended)
    vmgym_env_scalars* sc;
    float* cpu32; float* mem32;        // the agents' float32 view (env.py:296) with local accumulation
    uint16_t* act;                     // this step's action vector (external or chosen by a fused agent)
    uint8_t* tmp;                      // compaction / sort scratch
    uint16_t* cap;                     // per-PM capacity codes of the float32 view: cpu | mem << 8
    unsigned* fitm;                    // fit table over cpu codes (see rebuild_fit_table)
    const double* sz64; const float* sz32;   // code -> k/100.0 and (float)(k/100.0)
    const uint64_t* arr_cdf;           // arrival inverse-CDF thresholds (shared-memory copy when small)
    int P, V, lane;
};

__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}

// ---- numpy's scalar argsort replayed by one lane (best-fit compat tie mode) ------------
__device__ __noinline__ void introsort_argsort(const float* v, uint16_t* t, int num)
{
    // third-party algorithm: numpy npysort aquicksort_<float> (median-of-3 quicksort, insertion sort for
    // partitions of <= 16, larger side pushed); the heapsort fallback (depth limit) is kept for completeness.
    int pl = 0, pr = num - 1;
    int st_l[64], st_r[64], st_d[64];
    int sp = 0, cdepth = 0;
    for (int i = 0; i < num; i++) t[i] = (uint16_t)i;
    for (int n = num; n >>= 1;) cdepth++;
    cdepth *= 2;
    for (;;) {
        bool heap = false;
        if (cdepth < 0) heap = true;
        if (!heap) {
            while (pr - pl > 15) {
                int pm = pl + ((pr - pl) >> 1);
                uint16_t x;
                if (v[t[pm]] < v[t[pl]]) { x = t[pm]; t[pm] = t[pl]; t[pl] = x; }
                if (v[t[pr]] < v[t[pm]]) { x = t[pr]; t[pr] = t[pm]; t[pm] = x; }
                if (v[t[pm]] < v[t[pl]]) { x = t[pm]; t[pm] = t[pl]; t[pl] = x; }
                const float vp = v[t[pm]];
                int pi = pl, pj = pr - 1;
                x = t[pm]; t[pm] = t[pj]; t[pj] = x;
                for (;;) {
                    do { ++pi; } while (v[t[pi]] < vp);
                    do { --pj; } while (vp < v[t[pj]]);
                    if (pi >= pj) break;
                    x = t[pi]; t[pi] = t[pj]; t[pj] = x;
                }
                x = t[pi]; t[pi] = t[pr - 1]; t[pr - 1] = x;
                if (pi - pl < pr - pi) { st_l[sp] = pi + 1; st_r[sp] = pr; pr = pi - 1; }
                else { st_l[sp] = pl; st_r[sp] = pi - 1; pl = pi + 1; }
                st_d[sp] = --cdepth;
                sp++;
                if (cdepth < 0) { heap = true; break; }
            }
        }
        if (heap) {
            // heapsort of t[pl..pr] (1-based sift-down on a = t + pl - 1)
            uint16_t* a = t + pl - 1;
            int n = pr - pl + 1, i, j, l;
            uint16_t tmp;
            for (l = n >> 1; l > 0; --l) {
                tmp = a[l];
                for (i = l, j = l << 1; j <= n;) {
                    if (j < n && v[a[j]] < v[a[j + 1]]) j += 1;
                    if (v[tmp] < v[a[j]]) { a[i] = a[j]; i = j; j += j; } else break;
                }
                a[i] = tmp;
            }
            for (; n > 1;) {
                tmp = a[n]; a[n] = a[1]; n -= 1;
                for (i = 1, j = 2; j <= n;) {
                    if (j < n && v[a[j]] < v[a[j + 1]]) j++;
                    if (v[tmp] < v[a[j]]) { a[i] = a[j]; i = j; j += j; } else break;
                }
                a[i] = tmp;
            }
        } else {
            for (int pi = pl + 1; pi <= pr; ++pi) {
                const uint16_t vi = t[pi];
                const float vp = v[vi];
                int pj = pi;
                while (pj > pl && vp < v[t[pj - 1]]) { t[pj] = t[pj - 1]; pj--; }
                t[pj] = vi;
            }
        }
        if (sp == 0) break;
        sp--;
        pl = st_l[sp]; pr = st_r[sp]; cdepth = st_d[sp];
    }
}

// ---------------------------------------------------------------------------------------------------
// Heuristic agents on the float32 view (firstfit.py:21-38, bestfit.py:21-40).  lanes own PMs p = lane+32i.
// `c32(v)`, `m32(v)` give the VM sizes as the agent sees them; `waiting(v)` tells whether slot v is a
// waiting VM in the observation.  Writes e.act[v] for waiting VMs it places (others keep their placement).
// A VM that fits nowhere makes every later VM with component-wise >= sizes fit nowhere too (PM loads only
// grow inside act() and fp32 rounding is monotone), so such VMs are skipped without a scan.
// ---------------------------------------------------------------------------------------------------
// largest size code k in [0,100] with x + sz32[k] <= 1.0f (monotone in k because fp32 rounding is monotone)
__device__ __forceinline__ int max_code(const float* sz32, float x)
{
    int k = min(100, max(0, (int)((1.0f - x) * 100.0f)));      // estimate, then exact correction (usually 1-2 probes)
    while (k < 100 && x + sz32[k + 1] <= 1.0f) k++;
    while (k > 0 && x + sz32[k] > 1.0f) k--;
    return k;
}

// fitm[c] = 1 + max{ mem-capacity code of PM p : cpu-capacity code of p >= c }, 0 if no PM takes cpu code c.
// A VM with size codes (c, m) fits on SOME PM iff m + 1 <= fitm[c] — an exact O(1) test that removes the hopeless
// waiting VMs (the large majority at saturation) from the sequential scan.
template <typename PT>
__device__ __forceinline__ void rebuild_fit_table(Env<PT>& e)
{
    const int lane = e.lane;
    for (int c = lane; c < 128; c += 32) e.fitm[c] = 0u;
    __syncwarp();
    for (int p = lane; p < e.P; p += 32) {
        const unsigned w = e.cap[p];
        atomicMax(&e.fitm[w & 0xffu], (w >> 8) + 1u);
    }
    __syncwarp();
    uint4 q = reinterpret_cast<uint4*>(e.fitm)[lane];          // lane owns codes 4*lane .. 4*lane+3
    q.z = max(q.z, q.w); q.y = max(q.y, q.z); q.x = max(q.x, q.y);
    unsigned s = q.x;                                          // suffix max over lanes >= lane
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_down_sync(FULL, s, o);
        if (lane + o < 32) s = max(s, t);
    }
    unsigned ex = __shfl_down_sync(FULL, s, 1);
    if (lane == 31) ex = 0u;
    q.x = max(q.x, ex); q.y = max(q.y, ex); q.z = max(q.z, ex); q.w = max(q.w, ex);
    reinterpret_cast<uint4*>(e.fitm)[lane] = q;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
// Heuristic agents on the float32 view (firstfit.py:21-38, bestfit.py:21-40).  lanes own PMs p = lane+32i.
// `waiting(v)`: slot v is a waiting VM in the observation; `ccode/mcode(v)`: its size codes (hundredths) or -1
// when the observed size is not an exact hundredth (then the VM is always scanned); `c32of/m32of(v)`: the sizes
// as the agent sees them.  Writes e.act[v] for the waiting VMs it places (others keep their placement).
// ---------------------------------------------------------------------------------------------------
template <typename PT, class FW, class FK, class FL, class FC, class FM>
__device__ __forceinline__ int agent_act(Env<PT>& e, int agent, int tiebreak, FW waiting, FK ccode, FL mcode, FC c32of,
                                         FM m32of)
{
    const int P = e.P, V = e.V, lane = e.lane;
    int n_found = 0;
    for (int p = lane; p < P; p += 32)
        e.cap[p] = (uint16_t)(max_code(e.sz32, e.cpu32[p]) | (max_code(e.sz32, e.mem32[p]) << 8));
    rebuild_fit_table(e);
    for (int c0 = 0; c0 < V; c0 += 32) {
        const int v = c0 + lane;
        bool cand = v < V && waiting(v);
        if (cand) {
            const int cc = ccode(v), mc = mcode(v);
            if (cc >= 0 && mc >= 0) cand = (unsigned)(mc + 1) <= e.fitm[cc];
        }
        unsigned m = __ballot_sync(FULL, cand);
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int vv = c0 + b;
            const float c32 = c32of(vv), m32 = m32of(vv);
            int found = -1;
            if (agent == VMGYM_AGENT_FIRSTFIT) {
                for (int i0 = 0; i0 < P; i0 += 32) {
                    const int p = i0 + lane;
                    const bool fit = p < P && (e.cpu32[p] + c32 <= 1.0f) && (e.mem32[p] + m32 <= 1.0f);
                    const unsigned bb = __ballot_sync(FULL, fit);
                    if (bb) { found = i0 + __ffs(bb) - 1; break; }
                }
                if (found >= 0 && lane == (found & 31)) {
                    const float nc = e.cpu32[found] + c32;       // firstfit.py:36 — only the local cpu is updated
                    e.cpu32[found] = nc;
                    e.cap[found] = (uint16_t)((e.cap[found] & 0xff00u) | (unsigned)max_code(e.sz32, nc));
                }
            } else {
                // best-fit: first fitting PM in descending (cpu+memory) order (bestfit.py:33-39)
                unsigned bestk = 0;
                int bestp = -1;
                for (int p = lane; p < P; p += 32) {
                    const bool fit = (e.cpu32[p] + c32 <= 1.0f) && (e.mem32[p] + m32 <= 1.0f);
                    const unsigned kb = __float_as_uint(e.cpu32[p] + e.mem32[p]) + 1u;   // keys >= 0: bits order like values
                    if (fit && kb >= bestk) { bestk = kb; bestp = p; }
                }
                const unsigned gk = __reduce_max_sync(FULL, bestk);
                if (gk != 0) {
                    found = (int)__reduce_max_sync(FULL, (unsigned)((bestk == gk ? bestp : -1) + 1)) - 1;  // ties -> highest index
                    if (tiebreak == VMGYM_TIE_NUMPY_INTROSORT) {
                        int cnt = 0;
                        for (int p = lane; p < P; p += 32) {
                            const bool fit = (e.cpu32[p] + c32 <= 1.0f) && (e.mem32[p] + m32 <= 1.0f);
                            cnt += (fit && __float_as_uint(e.cpu32[p] + e.mem32[p]) + 1u == gk);
                        }
                        cnt = __reduce_add_sync(FULL, cnt);
                        if (cnt >= 2) {
                            // several fitting PMs share the maximal key: numpy's unstable default argsort decides
                            float* keys = reinterpret_cast<float*>(e.tmp);
                            uint16_t* perm = reinterpret_cast<uint16_t*>(e.tmp + 4 * ((P + 1) & ~1));
                            for (int p = lane; p < P; p += 32) keys[p] = e.cpu32[p] + e.mem32[p];
                            __syncwarp();
                            int pick = -1;
                            if (lane == 0) {
                                introsort_argsort(keys, perm, P);
                                for (int i = P - 1; i >= 0; i--) {
                                    const int p = perm[i];
                                    if ((e.cpu32[p] + c32 <= 1.0f) && (e.mem32[p] + m32 <= 1.0f)) { pick = p; break; }
                                }
                            }
                            found = __shfl_sync(FULL, pick, 0);
                            __syncwarp();
                        }
                    }
                    if (lane == (found & 31)) {
                        const float nc = e.cpu32[found] + c32, nm = e.mem32[found] + m32;   // bestfit.py:37-38
                        e.cpu32[found] = nc;
                        e.mem32[found] = nm;
                        e.cap[found] = (uint16_t)(max_code(e.sz32, nc) | (max_code(e.sz32, nm) << 8));
                    }
                }
            }
            if (found >= 0) {
                n_found++;
                if (lane == 0) e.act[vv] = (uint16_t)found;
                __syncwarp();
                rebuild_fit_table(e);              // capacities shrank: later candidates of this chunk are re-tested
                const int cc = v < V ? ccode(v) : -1, mc = v < V ? mcode(v) : -1;
                const bool still = !(cc >= 0 && mc >= 0) || (unsigned)(mc + 1) <= e.fitm[cc];
                m &= __ballot_sync(FULL, still);
            }
        }
    }
    __syncwarp();
    return n_found;
}

// ---------------------------------------------------------------------------------------------------
// Rewards `ut` (env.py:151-152) and `kl` (env.py:125-150, kl_divergence :8-17).  All reductions use numpy's
// summation order so the fp64 values (and the exact-zero variance tests) are those of the reference.
// ---------------------------------------------------------------------------------------------------
__device__ __noinline__ double reward_ut(const double* cpu, const double* mem, int P, double beta)
{
    const SumSrc sc{cpu, nullptr, nullptr, 0.0, 0}, sm{mem, nullptr, nullptr, 0.0, 0};
    return beta * np_sum(sc, P) + (1 - beta) * np_sum(sm, P);
}

__device__ __noinline__ double reward_kl(const double* cpu, const double* mem, int P, const uint8_t* ex_cc,
                                         const uint8_t* ex_mc, int arrived, const double* sz, int cap_target)
{
    const double dP = (double)P, dn = (double)arrived;
    const double ex_sum_c = np_sum(SumSrc{nullptr, ex_cc, sz, 0.0, 0}, arrived);
    const double ex_sum_m = np_sum(SumSrc{nullptr, ex_mc, sz, 0.0, 0}, arrived);
    double t_cpu = ex_sum_c / dP, t_mem = ex_sum_m / dP;                         // env.py:116,119
    if (cap_target && t_cpu > 1) t_cpu = 1.0;
    if (cap_target && t_mem > 1) t_mem = 1.0;
    const double cur_cpu = np_sum(SumSrc{cpu, nullptr, nullptr, 0.0, 0}, P) / dP;  // np.mean(self.cpu)
    const double cur_mem = np_sum(SumSrc{mem, nullptr, nullptr, 0.0, 0}, P) / dP;
    double cpu_var = np_sum(SumSrc{cpu, nullptr, nullptr, cur_cpu, 1}, P) / dP;    // np.var(self.cpu)
    double mem_var = np_sum(SumSrc{mem, nullptr, nullptr, cur_mem, 1}, P) / dP;
    if (cpu_var == 0) cpu_var = 1e-6;
    if (mem_var == 0) mem_var = 1e-6;
    // np.var(vm_cpu[existing]): deviations from the compacted array's own mean (sum / n)
    const double xm_c = ex_sum_c / dn, xm_m = ex_sum_m / dn;
    double t_cpu_var = np_sum(SumSrc{nullptr, ex_cc, sz, xm_c, 1}, arrived) / dn;
    double t_mem_var = np_sum(SumSrc{nullptr, ex_mc, sz, xm_m, 1}, arrived) / dn;
    if (t_cpu_var == 0) t_cpu_var = 1e-6;
    if (t_mem_var == 0) t_mem_var = 1e-6;
    if (t_cpu == 0 || t_mem == 0) return 0.0;
    // diagonal 2x2 covariances: det = product, inverse = reciprocals; evaluation order of env.py:17 kept
    const double det_p = t_cpu_var * t_mem_var, det_q = cpu_var * mem_var;
    const double qi0 = 1.0 / cpu_var, qi1 = 1.0 / mem_var;
    const double trace_term = qi0 * t_cpu_var + qi1 * t_mem_var;
    const double d0 = t_cpu - cur_cpu, d1 = t_mem - cur_mem;
    const double m1 = (d0 * qi0) * d0 + (d1 * qi1) * d1;
    return -(0.5 * (log(det_q / det_p) - 2 + trace_term + m1 - trace_term));
}

// ---------------------------------------------------------------------------------------------------
// One env.step on the shared-memory record.  Returns the reward (uniform across lanes).
// ---------------------------------------------------------------------------------------------------
struct StepResult { double reward; int terminated; int rejected; int waiting, arrived; int changed; };

constexpr uint32_t STATUS_EXHAUSTED = 1u;   // pre-sampled trace ran out (the reference would raise, env.py:282)
constexpr uint32_t STATUS_QUIET = 2u;       // a fused agent's act()+apply would change nothing (see step_kernel)
constexpr uint32_t STATUS_KEY_SHIFT = 8;    // bits 8..15: which (agent, tiebreak) established QUIET; 0 = any agent
constexpr uint32_t STATUS_KEY_MASK = 0xff00u;

// ---------------------------------------------------------------------------------------------------
// One env.step on the shared-memory record.  `have_actions` == false means "every action equals the current
// placement" (a fused agent that proposed nothing): phase 1 is skipped and every action is valid.
// The per-env counters n_waiting / n_empty are maintained incrementally so the common quiet step (nothing
// placed, nothing departs, nothing admitted) costs only the service countdown, one arrival draw and the outputs.
// ---------------------------------------------------------------------------------------------------
template <typename PT>
__device__ __forceinline__ StepResult env_step(Env<PT>& e, const StepParams& p, long long env_id, uint8_t* valid_g,
                                               bool have_actions)
{
    const int P = e.P, V = e.V, lane = e.lane;
    vmgym_env_scalars* sc = e.sc;
    int n_place = 0, n_susp = 0, rejected = 0;

    // ---- 1. apply actions in VM-index order, each seeing earlier updates (env.py:69-87, validate :35-42) ----
    if (have_actions) {
        for (int c0 = 0; c0 < V; c0 += 32) {
            const int v = c0 + lane;
            const int a = v < V ? (int)e.act[v] : 0;
            const int cur = v < V ? (int)e.place[v] : 0;
            const bool diff = v < V && a != cur;
            unsigned m = __ballot_sync(FULL, diff);
            unsigned okbits = 0;
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const int av = __shfl_sync(FULL, a, b), cv = __shfl_sync(FULL, cur, b), vv = c0 + b;
                bool ok = false;
                if (cv == P) {                                   // waiting VM: place iff it fits in fp64 (:38-39,55-56)
                    if ((unsigned)av < (unsigned)P) {
                        const double nc = e.cpu[av] + e.sz64[e.cpuc[vv] & 0x7f];
                        const double nm = e.mem[av] + e.sz64[e.memc[vv]];
                        if (nc <= 1.0 && nm <= 1.0) {
                            ok = true;
                            n_place++;
                            __syncwarp();
                            if (lane == 0) { e.cpu[av] = nc; e.mem[av] = nm; e.place[vv] = (PT)av; e.cpuc[vv] &= 0x7f; }  // :82-85
                        }
                    }
                } else if (cv < P) {                             // running VM: only suspend is legal (:40-41,78-81)
                    if (av == P) {
                        ok = true;
                        n_susp++;
                        const double nc = e.cpu[cv] - e.sz64[e.cpuc[vv] & 0x7f];
                        const double nm = e.mem[cv] - e.sz64[e.memc[vv]];
                        __syncwarp();
                        if (lane == 0) { e.cpu[cv] = nc; e.mem[cv] = nm; e.place[vv] = (PT)P; e.cpuc[vv] |= 0x80; }
                    }
                }
                __syncwarp();
                okbits |= ok ? (1u << b) : 0u;
            }
            const bool okv = !diff || ((okbits >> lane) & 1u);
            rejected += __popc(__ballot_sync(FULL, v < V && !okv));
            if (valid_g && v < V) valid_g[v] = okv ? 1 : 0;
        }
    } else if (valid_g) {
        for (int v = lane; v < V; v += 32) valid_g[v] = 1;
    }

    // ---- 2+3. service countdown and departures in VM-index order (_run_vms, env.py:244-265) ----
    int served = 0;
    if (sizeof(PT) == 1) {
        // 4 slots per lane: placement bytes as one u32, remaining runtimes as 4 x u16 (padding slots are empty)
        const uint32_t P4 = (uint32_t)P * 0x01010101u;
        const uint32_t* pl4 = reinterpret_cast<const uint32_t*>(e.place);
        uint2* rem4 = reinterpret_cast<uint2*>(e.rem);
        const int groups = (V + 3) >> 2;
        for (int g0 = 0; g0 < groups; g0 += 32) {
            const int g = g0 + lane;
            unsigned term4 = 0;
            if (g < groups) {
                const uint32_t run = __vcmpltu4(pl4[g], P4);            // 0xff per running slot
                if (run) {
                    uint2 r = rem4[g];
                    const uint32_t dlo = (run & 1u) | ((run & 0x100u) << 8), dhi = ((run >> 16) & 1u) | ((run >> 8) & 0x10000u);
                    r.x = __vsubus2(r.x, dlo);                           // if remaining > 0: remaining -= 1 (:245-247)
                    r.y = __vsubus2(r.y, dhi);
                    rem4[g] = r;
                    const uint32_t zlo = __vcmpeq2(r.x, 0u), zhi = __vcmpeq2(r.y, 0u);
                    term4 = ((zlo & 1u) | ((zlo >> 15) & 2u) | ((zhi & 1u) << 2) | ((zhi >> 13) & 8u)) &
                            ((run & 1u) | ((run >> 7) & 2u) | ((run >> 14) & 4u) | ((run >> 21) & 8u));
                }
            }
            unsigned m = __ballot_sync(FULL, term4 != 0);
            if (m) {                                                      // rare: some VM finished (:248-265)
                __syncwarp();
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    unsigned t4 = __shfl_sync(FULL, term4, b);
                    served += __popc(t4);
                    if (lane == 0) {
                        while (t4) {
                            const int j = __ffs(t4) - 1;
                            t4 &= t4 - 1;
                            const int vv = 4 * (g0 + b) + j, pm = (int)e.place[vv];
                            e.cpu[pm] -= e.sz64[e.cpuc[vv] & 0x7f];
                            e.mem[pm] -= e.sz64[e.memc[vv]];
                            e.place[vv] = (PT)(P + 1); e.cpuc[vv] = 0; e.memc[vv] = 0; e.rem[vv] = 0;
                        }
                    }
                }
                __syncwarp();
            }
        }
    } else {
        for (int c0 = 0; c0 < V; c0 += 32) {
            const int v = c0 + lane;
            const int pl = v < V ? (int)e.place[v] : P + 1;
            int r = v < V ? (int)e.rem[v] : 0;
            const bool running = pl < P;
            if (running && r > 0) { r -= 1; e.rem[v] = (uint16_t)r; }
            const bool term = running && r == 0;
            unsigned m = __ballot_sync(FULL, term);
            served += __popc(m);
            if (m) {
                if (lane == 0) {
                    unsigned mm = m;
                    while (mm) {
                        const int b = __ffs(mm) - 1;
                        mm &= mm - 1;
                        const int vv = c0 + b, pm = (int)e.place[vv];
                        e.cpu[pm] -= e.sz64[e.cpuc[vv] & 0x7f];
                        e.mem[pm] -= e.sz64[e.memc[vv]];
                    }
                }
                __syncwarp();
                if (term) { e.place[v] = (PT)(P + 1); e.cpuc[v] = 0; e.memc[v] = 0; e.rem[v] = 0; }
            }
        }
    }
    __syncwarp();
    // ---- 4. clamp (env.py:267-268): values only shrink when something was subtracted this step ----
    if (served > 0 || n_susp > 0) {
        for (int q = lane; q < P; q += 32) {
            if (e.cpu[q] < 1e-7) e.cpu[q] = 0.0;
            if (e.mem[q] < 1e-7) e.mem[q] = 0.0;
        }
    }

    // ---- 5. arrivals (_accept_vm_requests, env.py:271-293) ----
    int n_arr = 0;
    const vmgym_trace& tr = p.tr;
    if (tr.mode == VMGYM_TRACE_PRESAMPLED) {
        if ((long long)sc->arrival_pos < tr.arrivals_len) n_arr = tr.d_arrivals[env_id * tr.arrivals_len + sc->arrival_pos];
    } else {
        const Philox4 r = philox4x32_10(sc->arrival_pos, 0u, 1u, 0u, (uint32_t)sc->seed, (uint32_t)(sc->seed >> 32));
        const uint64_t u = ((uint64_t)r.x << 32) | r.y;
        const uint64_t* cdf = e.arr_cdf;            // shared-memory copy when it fits, else the global table
        int lo = 0, hi = tr.arrival_cdf_len;       // first i with cdf[i] > u
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
        n_arr = tr.arrival_kmin + min(lo, tr.arrival_cdf_len - 1);
    }
    int exhausted = (tr.mode == VMGYM_TRACE_PRESAMPLED && (long long)sc->arrival_pos >= tr.arrivals_len) ? 1 : 0;
    int quota = n_arr;                                         // admissions still allowed this step
    if (tr.mode == VMGYM_TRACE_PRESAMPLED) {
        const long long left = tr.admissions_len - (long long)sc->admission_pos;
        if ((long long)quota > left) { quota = (int)(left > 0 ? left : 0); exhausted = 1; }
    }
    const int n_empty0 = (int)sc->n_empty + served;            // empty slots before admission
    int admitted = 0;
    long long csum = 0, msum = 0;
    if (quota > 0 && n_empty0 > 0) {
        for (int c0 = 0; c0 < V && admitted < quota; c0 += 32) {
            const int v = c0 + lane;
            const bool empty = v < V && (int)e.place[v] == P + 1;
            const unsigned m = __ballot_sync(FULL, empty);
            const int rank = admitted + __popc(m & ((1u << lane) - 1u));
            if (empty && rank < quota) {                       // lowest-index empty slots (:275-277)
                const uint32_t j = sc->admission_pos + (uint32_t)rank;
                uint32_t cc, mc, svc;
                if (tr.mode == VMGYM_TRACE_PRESAMPLED) {
                    const uint32_t w = tr.d_admissions[env_id * tr.admissions_len + j];
                    cc = w & 0xff; mc = (w >> 8) & 0xff; svc = w >> 16;
                } else {
                    const Philox4 r = philox4x32_10(j, 0u, 2u, 0u, (uint32_t)sc->seed, (uint32_t)(sc->seed >> 32));
                    const uint32_t span = 2u * (uint32_t)(tr.size_hi_code - tr.size_lo_code);
                    cc = (uint32_t)tr.size_lo_code + ((mulhi32(r.x, span) + 1u) >> 1);
                    mc = (uint32_t)tr.size_lo_code + ((mulhi32(r.y, span) + 1u) >> 1);
                    const uint64_t u = ((uint64_t)r.z << 32) | r.w;
                    int lo = 0, hi = tr.service_cdf_len;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (tr.d_service_cdf[mid] <= u) lo = mid + 1; else hi = mid; }
                    svc = (uint32_t)(tr.service_kmin + min(lo, tr.service_cdf_len - 1)) + 1u;     // Poisson + 1 (:289)
                }
                e.place[v] = (PT)P;
                e.cpuc[v] = (uint8_t)cc; e.memc[v] = (uint8_t)mc; e.rem[v] = (uint16_t)svc;
                csum += cc; msum += mc;
            }
            admitted += __popc(m);
        }
        admitted = min(admitted, quota);
        csum = (long long)__reduce_add_sync(FULL, (unsigned)csum);
        msum = (long long)__reduce_add_sync(FULL, (unsigned)msum);
        __syncwarp();
    }

    // ---- 6. metrics (env.py:112-121) from the incrementally maintained slot counters ----
    const int n_empty = n_empty0 - admitted;
    const int waiting = (int)sc->n_waiting - n_place + n_susp + admitted;
    const int arrived = V - n_empty;

    // ---- 7. reward (env.py:123-156) ----
    double reward = 0.0;
    if (arrived > 0) {
        if (p.reward_fn == VMGYM_REWARD_WR) {
            reward = -((double)waiting / (double)arrived);
        } else if (p.reward_fn == VMGYM_REWARD_UT) {
            reward = reward_ut(e.cpu, e.mem, P, p.beta);
        } else {
            // compacted vm_cpu[existing], vm_memory[existing] in slot order (the reference's boolean indexing)
            uint8_t* ex_cc = e.tmp;
            uint8_t* ex_mc = e.tmp + ((V + 15) & ~15);
            int pos0 = 0;
            for (int c0 = 0; c0 < V; c0 += 32) {
                const int v = c0 + lane;
                const bool ex = v < V && (int)e.place[v] <= P;
                const unsigned mx = __ballot_sync(FULL, ex);
                if (ex) {
                    const int pos = pos0 + __popc(mx & ((1u << lane) - 1u));
                    ex_cc[pos] = e.cpuc[v] & 0x7f;
                    ex_mc[pos] = e.memc[v];
                }
                pos0 += __popc(mx);
            }
            __syncwarp();
            reward = reward_kl(e.cpu, e.mem, P, ex_cc, ex_mc, arrived, e.sz64, p.cap_target);
        }
    }

    // ---- 9. termination flag, counters, clock (env.py:160-163,101) ----
    const int terminated = sc->timestep >= p.step_limit;
    __syncwarp();
    if (lane == 0) {
        sc->total_requests += n_arr;
        sc->served_requests += served;
        sc->dropped_requests += n_arr - admitted;
        sc->suspend_actions += n_susp;
        sc->place_actions += n_place;
        sc->arrival_pos += 1;
        sc->admission_pos += (uint32_t)admitted;
        sc->status |= (uint32_t)exhausted;
        sc->n_waiting = (uint16_t)waiting;
        sc->n_empty = (uint16_t)n_empty;
        sc->cpu_code_sum += csum;
        sc->mem_code_sum += msum;
        sc->episode_return += reward;
        sc->last_reward = reward;
        sc->timestep += 1;
    }
    __syncwarp();
    StepResult res;
    res.reward = reward; res.terminated = terminated; res.rejected = rejected; res.waiting = waiting; res.arrived = arrived;
    res.changed = n_place + n_susp + served + admitted;      // anything that can change which waiting VMs fit
    return res;
}

// observation row (env.py:295-296): f32[ placement | vm_cpu | vm_memory | cpu | memory ]
template <typename PT>
__device__ __forceinline__ void write_obs(const Env<PT>& e, float* __restrict__ o)
{
    const int P = e.P, V = e.V;
    if (sizeof(PT) == 1 && (V & 3) == 0 && (P & 3) == 0) {
        // 128-bit stores: the row and its five segments are 16-byte aligned when V and P are multiples of 4
        float4* o4 = reinterpret_cast<float4*>(o);
        const uint32_t* pl4 = reinterpret_cast<const uint32_t*>(e.place);
        const uint32_t* cc4 = reinterpret_cast<const uint32_t*>(e.cpuc);
        const uint32_t* mc4 = reinterpret_cast<const uint32_t*>(e.memc);
        const int vg = V >> 2, pg = P >> 2;
        for (int g = e.lane; g < vg; g += 32) {
            const uint32_t a = pl4[g], c = cc4[g] & 0x7f7f7f7fu, m = mc4[g];
            o4[g] = make_float4((float)(a & 0xff), (float)((a >> 8) & 0xff), (float)((a >> 16) & 0xff), (float)(a >> 24));
            o4[vg + g] = make_float4(e.sz32[c & 0xff], e.sz32[(c >> 8) & 0xff], e.sz32[(c >> 16) & 0xff], e.sz32[c >> 24]);
            o4[2 * vg + g] = make_float4(e.sz32[m & 0xff], e.sz32[(m >> 8) & 0xff], e.sz32[(m >> 16) & 0xff], e.sz32[m >> 24]);
        }
        const double2* c2 = reinterpret_cast<const double2*>(e.cpu);
        const double2* m2 = reinterpret_cast<const double2*>(e.mem);
        for (int g = e.lane; g < pg; g += 32) {
            const double2 a = c2[2 * g], b = c2[2 * g + 1], c = m2[2 * g], d = m2[2 * g + 1];
            o4[3 * vg + g] = make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
            o4[3 * vg + pg + g] = make_float4((float)c.x, (float)c.y, (float)d.x, (float)d.y);
        }
        return;
    }
    for (int v = e.lane; v < V; v += 32) o[v] = (float)e.place[v];
    for (int v = e.lane; v < V; v += 32) o[V + v] = e.sz32[e.cpuc[v] & 0x7f];
    for (int v = e.lane; v < V; v += 32) o[2 * V + v] = e.sz32[e.memc[v]];
    for (int q = e.lane; q < P; q += 32) o[3 * V + q] = (float)e.cpu[q];
    for (int q = e.lane; q < P; q += 32) o[3 * V + P + q] = (float)e.mem[q];
}

template <typename PT>
__device__ __forceinline__ void bind_env(Env<PT>& e, unsigned char* base, const DevLayout& L, const double* sz64,
                                         const float* sz32, int lane)
{
    e.cpu = reinterpret_cast<double*>(base);
    e.mem = reinterpret_cast<double*>(base + L.off_mem);
    e.rem = reinterpret_cast<uint16_t*>(base + L.off_rem);
    e.place = reinterpret_cast<PT*>(base + L.off_place);
    e.cpuc = base + L.off_cpuc;
    e.memc = base + L.off_memc;
    e.sc = reinterpret_cast<vmgym_env_scalars*>(base + L.off_scal);
    e.cpu32 = reinterpret_cast<float*>(base + L.sm_cpu32);
    e.mem32 = reinterpret_cast<float*>(base + L.sm_mem32);
    e.act = reinterpret_cast<uint16_t*>(base + L.sm_act);
    e.tmp = base + L.sm_tmp;
    e.fitm = reinterpret_cast<unsigned*>(base + L.sm_fit);
    e.cap = reinterpret_cast<uint16_t*>(base + L.sm_fit + 512);
    e.sz64 = sz64; e.sz32 = sz32; e.P = L.P; e.V = L.V; e.lane = lane;
}

__device__ __forceinline__ void fill_tables(double* sz64, float* sz32)
{
    for (int k = threadIdx.x; k < SIZE_TABLE; k += blockDim.x) {
        const double x = (double)k / 100.0;      // == np.around(u, 2) for the code k (env.py:212-219)
        sz64[k] = x;
        sz32[k] = (float)x;                        // env.py:296 float32 cast
    }
}

// action element -> int; anything outside [0, 65534] becomes 0xFFFF, which matches no placement value and
// therefore fails every branch of validate() (env.py:35-42) exactly like an out-of-range action does.
__device__ __forceinline__ int load_action(const void* row, int dtype, int v)
{
    if (dtype == VMGYM_U8) return (int)reinterpret_cast<const uint8_t*>(row)[v];
    if (dtype == VMGYM_I16) {
        const int x = reinterpret_cast<const int16_t*>(row)[v];
        return x < 0 ? 0xFFFF : x;
    }
    const long long x = reinterpret_cast<const long long*>(row)[v];
    return (x < 0 || x > 65534) ? 0xFFFF : (int)x;
}
__host__ __device__ __forceinline__ int dtype_bytes(int dtype) { return dtype == VMGYM_U8 ? 1 : (dtype == VMGYM_I16 ? 2 : 8); }

// ---------------------------------------------------------------------------------------------------
// The step kernel: external actions (agent == NONE) or fused heuristic agent, n_steps per launch.
// grid-stride over envs, one warp per env.
// ---------------------------------------------------------------------------------------------------
template <typename PT>
__global__ void __launch_bounds__(256, 4) step_kernel(const StepParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const DevLayout& L = p.L;
    double* sz64 = reinterpret_cast<double*>(smem);
    float* sz32 = reinterpret_cast<float*>(smem + SIZE_TABLE * 8);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    unsigned char* base = smem + L.sm_tables + (size_t)warp * L.sm_stride;
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + L.sm_bar);
    fill_tables(sz64, sz32);
    uint64_t* arr_cdf_s = reinterpret_cast<uint64_t*>(smem + SIZE_TABLE * 12);
    const bool cdf_in_smem = p.tr.mode == VMGYM_TRACE_PHILOX && p.tr.arrival_cdf_len <= ARR_CDF_SMEM;
    if (cdf_in_smem)
        for (int k = threadIdx.x; k < p.tr.arrival_cdf_len; k += blockDim.x) arr_cdf_s[k] = p.tr.d_arrival_cdf[k];
    const bool BULK = p.use_bulk != 0;
    if (BULK && lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncthreads();

    Env<PT> e;
    bind_env(e, base, L, sz64, sz32, lane);
    e.arr_cdf = cdf_in_smem ? arr_cdf_s : p.tr.d_arrival_cdf;
    uint32_t phase = 0;
    const long long stride = (long long)gridDim.x * wpc;
    for (long long env = (long long)blockIdx.x * wpc + warp; env < p.n_envs; env += stride) {
        unsigned char* grec = p.state + env * (long long)L.rec_bytes;
        // ---- stage the record into shared memory ----
        if (BULK) {
            if (lane == 0) {
                mbar_arrive_expect_tx(bar, (uint32_t)L.rec_bytes);
                bulk_g2s(base, grec, (uint32_t)L.rec_bytes, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
        } else {
            const uint4* src = reinterpret_cast<const uint4*>(grec);
            uint4* dst = reinterpret_cast<uint4*>(base);
            for (int i = lane; i < L.rec_bytes / 16; i += 32) dst[i] = __ldg(src + i);
            __syncwarp();
        }

        uint8_t* valid_g = p.out.d_valid ? p.out.d_valid + env * (long long)L.V : nullptr;
        StepResult res;
        res.reward = 0.0; res.terminated = 0; res.rejected = 0; res.waiting = 0; res.arrived = 0; res.changed = 0;
        double st_drop = 0, st_wr = 0, st_mc = 0, st_vc = 0, st_mm = 0, st_vm = 0, st_rej = 0, st_n = 0;
        // STATUS_QUIET: "a fused agent's act() followed by the env's apply loop would change nothing".  It is
        // established by a full evaluation after which the step changed no placement (the agent proposed nothing, or
        // every proposal was rejected by the fp64 capacity check — SURVEY App. B-2) and nothing departed or was
        // admitted: the next act() then sees the same float32 PM loads and the same waiting VMs, proposes the same
        // actions and the env rejects them again.  While it holds, act() and the apply loop are skipped.  The key
        // records which agent established it (0 = "no waiting VM fits anywhere", which holds for every agent).
        const bool need_vectors = p.out.d_action != nullptr || p.out.d_valid != nullptr;
        const uint32_t my_key = (uint32_t)(p.agent | (p.tiebreak << 4));
        uint32_t status = e.sc->status;
        bool quiet = (status & STATUS_QUIET) != 0;
        uint32_t quiet_key = (status & STATUS_KEY_MASK) >> STATUS_KEY_SHIFT;
        int quiet_rejected = (int)(status >> 16);
        for (int s = 0; s < p.n_steps; s++) {
            bool have_actions;
            bool evaluated = false;
            int n_found = 0;
            if (p.agent != VMGYM_AGENT_NONE) {
                if (!(quiet && (quiet_key == 0 || quiet_key == my_key)) || need_vectors) {
                    // the agent sees the float32 observation of the current state (env.py:296)
                    for (int q = lane; q < L.P; q += 32) { e.cpu32[q] = (float)e.cpu[q]; e.mem32[q] = (float)e.mem[q]; }
                    for (int v = lane; v < L.V; v += 32) e.act[v] = (uint16_t)e.place[v];
                    __syncwarp();
                    const PT* place = e.place; const uint8_t* cpuc = e.cpuc; const uint8_t* memc = e.memc;
                    const float* t32 = sz32;
                    const int P = L.P;
                    n_found = agent_act(e, p.agent, p.tiebreak, [=](int v) { return (int)place[v] == P; },
                                        [=](int v) { return (int)(cpuc[v] & 0x7f); }, [=](int v) { return (int)memc[v]; },
                                        [=](int v) { return t32[cpuc[v] & 0x7f]; }, [=](int v) { return t32[memc[v]]; });
                    evaluated = true;
                }
                have_actions = n_found > 0;
            } else {
                const int adt = p.action_dtype;
                const unsigned char* arow = reinterpret_cast<const unsigned char*>(p.action) + env * (long long)L.V * dtype_bytes(adt);
                for (int v = lane; v < L.V; v += 32) e.act[v] = (uint16_t)load_action(arow, adt, v);
                __syncwarp();
                have_actions = true;
            }
            res = env_step(e, p, env, valid_g, have_actions);
            if (res.changed) {
                quiet = false;
            } else if (evaluated) {
                quiet = true;
                quiet_key = n_found == 0 ? 0u : my_key;
                quiet_rejected = res.rejected;
            } else if (quiet && p.agent != VMGYM_AGENT_NONE) {
                res.rejected = quiet_rejected;        // the skipped proposals would have been rejected again
            }
            if (p.out.d_stats) {
                // running sums for the eval summary (record.py:98-134, exp_performance.py:104-113)
                double sc_ = 0, sm_ = 0;
                for (int q = lane; q < L.P; q += 32) { sc_ += e.cpu[q]; sm_ += e.mem[q]; }
                const double mc = warp_sum(sc_) / L.P, mm = warp_sum(sm_) / L.P;
                double vc = 0, vm = 0;
                for (int q = lane; q < L.P; q += 32) {
                    const double dc = e.cpu[q] - mc, dm = e.mem[q] - mm;
                    vc += dc * dc; vm += dm * dm;
                }
                vc = warp_sum(vc) / L.P; vm = warp_sum(vm) / L.P;
                const int tot = e.sc->total_requests;
                st_drop += tot ? (double)e.sc->dropped_requests / (double)tot : 0.0;
                st_wr += res.arrived ? (double)res.waiting / (double)res.arrived : 0.0;
                st_mc += mc; st_vc += vc; st_mm += mm; st_vm += vm; st_rej += res.rejected; st_n += 1;
            }
            if (res.terminated) break;
        }

        // ---- outputs ----
        if (p.out.d_obs) write_obs(e, p.out.d_obs + env * (long long)L.D);
        if (p.out.d_action && p.agent != VMGYM_AGENT_NONE) {
            PT* ao = reinterpret_cast<PT*>(p.out.d_action) + env * (long long)L.V;
            for (int v = lane; v < L.V; v += 32) ao[v] = (PT)e.act[v];    // need_vectors forced the evaluation
        }
        if (lane == 0) {
            e.sc->status = (e.sc->status & STATUS_EXHAUSTED) |
                           (quiet ? (STATUS_QUIET | (quiet_key << STATUS_KEY_SHIFT) | ((uint32_t)quiet_rejected << 16)) : 0u);
            if (p.out.d_reward) p.out.d_reward[env] = res.reward;
            if (p.out.d_terminated) p.out.d_terminated[env] = (uint8_t)res.terminated;
            if (p.out.d_stats) {
                double* st = p.out.d_stats + env * 8;
                st[0] += st_drop; st[1] += st_wr; st[2] += st_mc; st[3] += st_vc; st[4] += st_mm; st[5] += st_vm;
                st[6] += st_rej; st[7] += st_n;
            }
        }

        // ---- write the record back ----
        if (BULK) {
            fence_proxy_async();          // generic-proxy writes to smem -> visible to the async proxy
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(grec, base, (uint32_t)L.rec_bytes);
                bulk_commit();
                bulk_wait_read0();        // smem may be overwritten by the next bulk load after this
            }
            __syncwarp();
        } else {
            __syncwarp();
            const uint4* src = reinterpret_cast<const uint4*>(base);
            uint4* dst = reinterpret_cast<uint4*>(grec);
            for (int i = lane; i < L.rec_bytes / 16; i += 32) dst[i] = src[i];
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// reset / observe / invalid-action mask / agent.act(obs)
// ---------------------------------------------------------------------------------------------------
template <typename PT>
__global__ void reset_kernel(DevLayout L, unsigned char* state, long long n_envs, const uint8_t* env_mask,
                             const uint64_t* seeds, int rewind, float* obs)
{
    const int lane = threadIdx.x & 31;
    const long long env = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (env >= n_envs) return;
    if (env_mask && !env_mask[env]) return;
    unsigned char* rec = state + env * (long long)L.rec_bytes;
    vmgym_env_scalars* sc = reinterpret_cast<vmgym_env_scalars*>(rec + L.off_scal);
    vmgym_env_scalars keep = *sc;
    __syncwarp();
    uint4* r4 = reinterpret_cast<uint4*>(rec);
    for (int i = lane; i < L.rec_bytes / 16; i += 32) r4[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    PT* place = reinterpret_cast<PT*>(rec + L.off_place);
    for (int v = lane; v < L.Vp; v += 32) place[v] = (PT)(L.P + 1);      // env.py:187 (padding slots stay empty forever)
    if (lane == 0) {
        sc->timestep = 1;                                                 // env.py:197
        sc->n_waiting = 0;
        sc->n_empty = (uint16_t)L.V;
        sc->seed = seeds ? seeds[env] : keep.seed;
        sc->arrival_pos = rewind ? 0u : keep.arrival_pos;
        sc->admission_pos = rewind ? 0u : keep.admission_pos;
        sc->status = rewind ? 0u : (keep.status & STATUS_EXHAUSTED);
    }
    if (obs) {
        float* o = obs + env * (long long)L.D;
        for (int i = lane; i < L.D; i += 32) o[i] = i < L.V ? (float)(L.P + 1) : 0.0f;
    }
}

template <typename PT>
__global__ void observe_kernel(DevLayout L, const unsigned char* state, long long n_envs, float* obs)
{
    const int lane = threadIdx.x & 31;
    const long long env = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (env >= n_envs) return;
    const unsigned char* rec = state + env * (long long)L.rec_bytes;
    const double* cpu = reinterpret_cast<const double*>(rec);
    const double* mem = reinterpret_cast<const double*>(rec + L.off_mem);
    const PT* place = reinterpret_cast<const PT*>(rec + L.off_place);
    const uint8_t* cpuc = rec + L.off_cpuc;
    const uint8_t* memc = rec + L.off_memc;
    float* o = obs + env * (long long)L.D;
    const int V = L.V, P = L.P;
    for (int v = lane; v < V; v += 32) {
        o[v] = (float)place[v];
        o[V + v] = (float)((double)(cpuc[v] & 0x7f) / 100.0);
        o[2 * V + v] = (float)((double)memc[v] / 100.0);
    }
    for (int q = lane; q < P; q += 32) { o[3 * V + q] = (float)cpu[q]; o[3 * V + P + q] = (float)mem[q]; }
}

// get_invalid_action_mask (env.py:45-53), evaluated against the current state (not sequentially).
template <typename PT>
__global__ void mask_kernel(DevLayout L, const unsigned char* state, long long n_envs, uint8_t* mask)
{
    const int lane = threadIdx.x & 31;
    const long long env = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (env >= n_envs) return;
    const unsigned char* rec = state + env * (long long)L.rec_bytes;
    const double* cpu = reinterpret_cast<const double*>(rec);
    const double* mem = reinterpret_cast<const double*>(rec + L.off_mem);
    const PT* place = reinterpret_cast<const PT*>(rec + L.off_place);
    const uint8_t* cpuc = rec + L.off_cpuc;
    const uint8_t* memc = rec + L.off_memc;
    const int V = L.V, P = L.P, A = L.A;
    uint8_t* out = mask + env * (long long)V * A;
    const long long total = (long long)V * A;
    for (long long i = lane; i < total; i += 32) {
        const int v = (int)(i / A), a = (int)(i - (long long)v * A);
        const int cur = (int)place[v];
        bool valid;
        if (a == cur) valid = true;
        else if (cur == P) {
            valid = false;
            if (a < P) {
                const double vc = (double)(cpuc[v] & 0x7f) / 100.0, vm = (double)memc[v] / 100.0;
                valid = (cpu[a] + vc <= 1.0) && (mem[a] + vm <= 1.0);
            }
        } else if (cur < P) valid = (a == P);
        else valid = false;
        out[i] = valid ? 0 : 1;
    }
}

// agent.act(observation) on float32 observations [n_envs, D] (firstfit.py:21-38, bestfit.py:21-40).
// per-warp shared memory: obs row f32[D] | act PT[Vp] | tmp (sort scratch) | fitm u32[128] | cap u16[Pp]
__host__ __device__ inline int act_row_bytes(const DevLayout& L) { return align_up(4 * L.D, 16); }
__host__ __device__ inline int act_smem_per_warp(const DevLayout& L)
{
    return act_row_bytes(L) + 2 * L.Vp + align_up(6 * L.Pp, 16) + 512 + align_up(2 * L.Pp, 16);
}

template <typename PT>
__global__ void act_kernel(DevLayout L, int agent, int tiebreak, const float* obs, long long n_envs, void* action, int adt)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    float* sz32 = reinterpret_cast<float*>(smem);
    for (int k = threadIdx.x; k < SIZE_TABLE; k += blockDim.x) sz32[k] = (float)((double)k / 100.0);
    __syncthreads();
    const int row_bytes = act_row_bytes(L);
    unsigned char* base = smem + SIZE_TABLE * 4 + (size_t)warp * act_smem_per_warp(L);
    float* row = reinterpret_cast<float*>(base);
    const long long env = (long long)blockIdx.x * wpc + warp;
    if (env >= n_envs) return;
    const float* o = obs + env * (long long)L.D;
    for (int i = lane; i < L.D; i += 32) row[i] = o[i];
    __syncwarp();
    Env<PT> e;
    e.P = L.P; e.V = L.V; e.lane = lane;
    e.sz32 = sz32; e.sz64 = nullptr;
    e.cpu32 = row + 3 * L.V; e.mem32 = row + 3 * L.V + L.P;
    e.act = reinterpret_cast<uint16_t*>(base + row_bytes);
    e.tmp = base + row_bytes + 2 * L.Vp;
    e.fitm = reinterpret_cast<unsigned*>(e.tmp + align_up(6 * L.Pp, 16));
    e.cap = reinterpret_cast<uint16_t*>(e.tmp + align_up(6 * L.Pp, 16) + 512);
    const int V = L.V, P = L.P;
    for (int v = lane; v < V; v += 32) e.act[v] = (uint16_t)(int)row[v];      // utils.py:41 astype(int)
    __syncwarp();
    // size code of an observed size: the hundredth whose float32 image equals it, else -1 (never filtered)
    auto code_of = [=](float x) {
        const int k = __float2int_rn(x * 100.0f);
        return (k >= 0 && k <= 100 && sz32[k] == x) ? k : -1;
    };
    agent_act(e, agent, tiebreak, [=](int v) { return (int)row[v] == P; }, [=](int v) { return code_of(row[V + v]); },
              [=](int v) { return code_of(row[2 * V + v]); }, [=](int v) { return row[V + v]; },
              [=](int v) { return row[2 * V + v]; });
    unsigned char* ao = reinterpret_cast<unsigned char*>(action) + env * (long long)V * dtype_bytes(adt);
    for (int v = lane; v < V; v += 32) {
        const int a = (int)e.act[v];
        if (adt == VMGYM_U8) ao[v] = (uint8_t)a;
        else if (adt == VMGYM_I16) reinterpret_cast<int16_t*>(ao)[v] = (int16_t)a;
        else reinterpret_cast<long long*>(ao)[v] = a;
    }
}

}  // namespace vmgym

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
using namespace vmgym;

static int check_cuda(cudaError_t err, const char* what)
{
    if (err == cudaSuccess) return VMGYM_OK;
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(err));
    return VMGYM_ECUDA;
}

static int g_sm_count = 0;
static int sm_count()
{
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

// warps per CTA: keep >= ~6 CTAs of work per SM when the batch is small (balance across the 148 SMs), up to 8
// warps when it is large.
static int pick_warps(long long n_envs, int smem_per_warp, int smem_fixed)
{
    if (g_warps_per_cta > 0) return g_warps_per_cta;
    int w = 8;
    while (w > 1 && n_envs < (long long)sm_count() * w * 6) w >>= 1;
    while (w > 1 && smem_fixed + w * smem_per_warp > 200 * 1024) w >>= 1;
    return w;
}

template <typename PT>
static int launch_step(StepParams& sp, cudaStream_t st)
{
    sp.use_bulk = g_use_bulk;
    const DevLayout& L = sp.L;
    const int w = pick_warps(sp.n_envs, L.sm_stride, L.sm_tables);
    const size_t smem = (size_t)L.sm_tables + (size_t)w * L.sm_stride;
    if (smem > 227 * 1024) return fail(VMGYM_EUNSUPPORTED, "env record does not fit in shared memory (pms/vms too large)");
    auto kern = step_kernel<PT>;
    // per (kernel, smem, warps) launch plan, computed once (also keeps these calls out of CUDA-graph capture)
    static thread_local size_t plan_smem = 0;
    static thread_local int plan_w = 0, plan_occ = 1;
    if (plan_smem != smem || plan_w != w) {
        int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute");
        if (rc) return rc;
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, w * 32, smem);
        plan_occ = occ < 1 ? 1 : occ;
        plan_smem = smem; plan_w = w;
    }
    const int occ = plan_occ;
    long long blocks = (sp.n_envs + w - 1) / w;
    const long long cap = (long long)sm_count() * occ;
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, w * 32, smem, st>>>(sp);
    return check_cuda(cudaGetLastError(), "step_kernel launch");
}

static int fill_params(StepParams* sp, const vmgym_config* cfg, void* d_state, int64_t n_envs, const vmgym_trace* trace,
                       const vmgym_outputs* out)
{
    int rc = make_layout(cfg, &sp->L, nullptr);
    if (rc) return rc;
    if (!d_state || n_envs < 0) return fail(VMGYM_EINVAL, "null state / negative n_envs");
    if (!trace) return fail(VMGYM_EINVAL, "null trace");
    if (trace->mode == VMGYM_TRACE_PRESAMPLED) {
        if (!trace->d_arrivals || !trace->d_admissions) return fail(VMGYM_EINVAL, "pre-sampled trace arrays missing");
    } else if (trace->mode == VMGYM_TRACE_PHILOX) {
        if (!trace->d_arrival_cdf || !trace->d_service_cdf || trace->arrival_cdf_len < 1 || trace->service_cdf_len < 1)
            return fail(VMGYM_EINVAL, "philox CDF tables missing");
        if (trace->size_lo_code < 0 || trace->size_hi_code > 100 || trace->size_hi_code <= trace->size_lo_code)
            return fail(VMGYM_EINVAL, "bad size code range");
    } else return fail(VMGYM_EINVAL, "unknown trace mode");
    sp->reward_fn = cfg->reward_function; sp->cap_target = cfg->cap_target_util; sp->step_limit = cfg->step_limit;
    sp->beta = cfg->beta;
    sp->state = (unsigned char*)d_state; sp->n_envs = n_envs; sp->tr = *trace;
    if (out) sp->out = *out; else memset(&sp->out, 0, sizeof(sp->out));
    sp->action = nullptr; sp->action_dtype = VMGYM_U8; sp->use_bulk = 1; sp->agent = VMGYM_AGENT_NONE; sp->tiebreak = 0; sp->n_steps = 1;
    return VMGYM_OK;
}

extern "C" {

const char* vmgym_last_error(void) { return g_err; }
int vmgym_abi_version(void) { return VMGYM_ABI_VERSION; }

int vmgym_set_tuning(int warps_per_cta, int use_bulk_copy)
{
    if (warps_per_cta < 0 || warps_per_cta > 16) return fail(VMGYM_EINVAL, "warps_per_cta must be 0..16");
    g_warps_per_cta = warps_per_cta;
    g_use_bulk = use_bulk_copy ? 1 : 0;
    return VMGYM_OK;
}

int vmgym_get_layout(const vmgym_config* cfg, vmgym_layout* out)
{
    if (!out) return fail(VMGYM_EINVAL, "null layout");
    return make_layout(cfg, nullptr, out);
}

int vmgym_reset(const vmgym_config* cfg, void* d_state, int64_t n_envs, const uint8_t* d_env_mask, const uint64_t* d_seeds,
                int rewind_streams, float* d_obs, void* stream)
{
    DevLayout L;
    int rc = make_layout(cfg, &L, nullptr);
    if (rc) return rc;
    if (!d_state || n_envs < 0) return fail(VMGYM_EINVAL, "null state / negative n_envs");
    if (n_envs == 0) return VMGYM_OK;
    const int threads = 256;
    const long long blocks = (n_envs * 32 + threads - 1) / threads;
    if (L.P <= 253)
        reset_kernel<uint8_t><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(L, (unsigned char*)d_state, n_envs, d_env_mask, d_seeds, rewind_streams, d_obs);
    else
        reset_kernel<uint16_t><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(L, (unsigned char*)d_state, n_envs, d_env_mask, d_seeds, rewind_streams, d_obs);
    return check_cuda(cudaGetLastError(), "reset_kernel launch");
}

int vmgym_step(const vmgym_config* cfg, void* d_state, int64_t n_envs, const vmgym_trace* trace, const void* d_action,
               int action_dtype, const vmgym_outputs* out, void* stream)
{
    StepParams sp;
    int rc = fill_params(&sp, cfg, d_state, n_envs, trace, out);
    if (rc) return rc;
    if (!d_action) return fail(VMGYM_EINVAL, "null action");
    if (n_envs == 0) return VMGYM_OK;
    sp.action = d_action;
    sp.action_dtype = action_dtype;
    cudaStream_t st = (cudaStream_t)stream;
    const bool small = sp.L.P <= 253;
    if (action_dtype != VMGYM_U8 && action_dtype != VMGYM_I16 && action_dtype != VMGYM_I64)
        return fail(VMGYM_EINVAL, "unknown action dtype");
    if (action_dtype == VMGYM_U8 && !small) return fail(VMGYM_EINVAL, "u8 actions need pms <= 253");
    return small ? launch_step<uint8_t>(sp, st) : launch_step<uint16_t>(sp, st);
}

int vmgym_agent_step(const vmgym_config* cfg, void* d_state, int64_t n_envs, const vmgym_trace* trace, int agent,
                     int tiebreak, int n_steps, const vmgym_outputs* out, void* stream)
{
    StepParams sp;
    int rc = fill_params(&sp, cfg, d_state, n_envs, trace, out);
    if (rc) return rc;
    if (agent != VMGYM_AGENT_FIRSTFIT && agent != VMGYM_AGENT_BESTFIT)
        return fail(VMGYM_EUNSUPPORTED, "fused agent must be firstfit or bestfit");
    if (tiebreak != VMGYM_TIE_STABLE && tiebreak != VMGYM_TIE_NUMPY_INTROSORT) return fail(VMGYM_EINVAL, "unknown tiebreak");
    if (n_steps < 1) return fail(VMGYM_EINVAL, "n_steps must be >= 1");
    if (n_envs == 0) return VMGYM_OK;
    sp.agent = agent; sp.tiebreak = tiebreak; sp.n_steps = n_steps;
    cudaStream_t st = (cudaStream_t)stream;
    return sp.L.P <= 253 ? launch_step<uint8_t>(sp, st) : launch_step<uint16_t>(sp, st);
}

int vmgym_agent_act(const vmgym_config* cfg, int agent, int tiebreak, const float* d_obs, int64_t n_envs, void* d_action,
                    int action_dtype, void* stream)
{
    DevLayout L;
    int rc = make_layout(cfg, &L, nullptr);
    if (rc) return rc;
    if (!d_obs || !d_action || n_envs < 0) return fail(VMGYM_EINVAL, "null obs/action");
    if (agent != VMGYM_AGENT_FIRSTFIT && agent != VMGYM_AGENT_BESTFIT) return fail(VMGYM_EUNSUPPORTED, "agent must be firstfit or bestfit");
    if (n_envs == 0) return VMGYM_OK;
    const int per_warp = act_smem_per_warp(L);
    int w = 4;
    while (w > 1 && (size_t)w * per_warp > 200 * 1024) w >>= 1;
    const size_t smem = (size_t)SIZE_TABLE * 4 + (size_t)w * per_warp;
    if (smem > 227 * 1024) return fail(VMGYM_EUNSUPPORTED, "observation row does not fit in shared memory");
    const long long blocks = (n_envs + w - 1) / w;
    cudaStream_t st = (cudaStream_t)stream;
    const bool small = L.P <= 253;
#define VMGYM_LAUNCH_ACT(PT)                                                                                      \
    do {                                                                                                          \
        rc = check_cuda(cudaFuncSetAttribute(act_kernel<PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                        "cudaFuncSetAttribute");                                                                  \
        if (rc) return rc;                                                                                        \
        act_kernel<PT><<<(unsigned)blocks, w * 32, smem, st>>>(L, agent, tiebreak, d_obs, n_envs, d_action, action_dtype); \
    } while (0)
    if (action_dtype != VMGYM_U8 && action_dtype != VMGYM_I16 && action_dtype != VMGYM_I64)
        return fail(VMGYM_EINVAL, "unknown action dtype");
    if (action_dtype == VMGYM_U8 && !small) return fail(VMGYM_EINVAL, "u8 actions need pms <= 253");
    if (small) VMGYM_LAUNCH_ACT(uint8_t); else VMGYM_LAUNCH_ACT(uint16_t);
#undef VMGYM_LAUNCH_ACT
    return check_cuda(cudaGetLastError(), "act_kernel launch");
}

int vmgym_observe(const vmgym_config* cfg, const void* d_state, int64_t n_envs, float* d_obs, void* stream)
{
    DevLayout L;
    int rc = make_layout(cfg, &L, nullptr);
    if (rc) return rc;
    if (!d_state || !d_obs || n_envs < 0) return fail(VMGYM_EINVAL, "null state/obs");
    if (n_envs == 0) return VMGYM_OK;
    const int threads = 256;
    const long long blocks = (n_envs * 32 + threads - 1) / threads;
    if (L.P <= 253) observe_kernel<uint8_t><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(L, (const unsigned char*)d_state, n_envs, d_obs);
    else observe_kernel<uint16_t><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(L, (const unsigned char*)d_state, n_envs, d_obs);
    return check_cuda(cudaGetLastError(), "observe_kernel launch");
}

int vmgym_invalid_action_mask(const vmgym_config* cfg, const void* d_state, int64_t n_envs, uint8_t* d_mask, void* stream)
{
    DevLayout L;
    int rc = make_layout(cfg, &L, nullptr);
    if (rc) return rc;
    if (!d_state || !d_mask || n_envs < 0) return fail(VMGYM_EINVAL, "null state/mask");
    if (n_envs == 0) return VMGYM_OK;
    const int threads = 256;
    const long long blocks = (n_envs * 32 + threads - 1) / threads;
    if (L.P <= 253) mask_kernel<uint8_t><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(L, (const unsigned char*)d_state, n_envs, d_mask);
    else mask_kernel<uint16_t><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(L, (const unsigned char*)d_state, n_envs, d_mask);
    return check_cuda(cudaGetLastError(), "mask_kernel launch");
}

}  // extern "C"
