// vmgym_env_kernels.cuh — device code of the batched env hot path (sm_100a).
//
// Semantics follow the reference's vmenv/envs/env.py and src/agents/{firstfit,bestfit}.py line by line (cited at
// each phase); the mapping onto the GPU is new: one warp owns one env, the env record (vmgym_layout) is staged
// HBM -> shared memory by one bulk-async copy (cp.async.bulk + mbarrier), mutated there for n_steps steps, and
// written back by one bulk-async store.  fp64 PM accumulators are updated in the reference's VM-index order so
// capacity decisions are bit-identical.  Compiled with -fmad=false: no FMA contraction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vmgym_device.cuh"
#include "vmgym_sort.cuh"

namespace vmgym {

// Section timing of the step kernel's decision warp (development builds only: -DVMGYM_PROF): cycles per section summed over all envs
// into g_prof[], read back with vmgym_debug_prof().  Compiles to nothing otherwise.
#ifdef VMGYM_PROF
__device__ unsigned long long g_prof[16];
#define PROF_T0() long long _pt = clock64()
#define PROF_ADD(i)                                                                                          \
    do {                                                                                                      \
        const long long _n = clock64();                                                                       \
        if ((threadIdx.x & 31) == 0 && threadIdx.x < 32) atomicAdd(&g_prof[i], (unsigned long long)(_n - _pt)); \
        _pt = _n;                                                                                             \
    } while (0)
#define PROF_SPAN_BEGIN() const long long _ps = clock64()
#define PROF_SPAN_END(i, cnt)                                                                                 \
    do {                                                                                                      \
        if ((threadIdx.x & 31) == 0 && threadIdx.x < 32) {                                                    \
            atomicAdd(&g_prof[i], (unsigned long long)(clock64() - _ps));                                     \
            if ((cnt) >= 0) atomicAdd(&g_prof[cnt], 1ull);                                                    \
        }                                                                                                     \
    } while (0)
#else
#define PROF_T0() do { } while (0)
#define PROF_ADD(i) do { } while (0)
#define PROF_SPAN_BEGIN() do { } while (0)
#define PROF_SPAN_END(i, cnt) do { } while (0)
#endif

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
    int use_bulk;             // bit 0: load records with cp.async.bulk, bit 1: store them with cp.async.bulk (else 128-bit ld/st),
                              // bit 2: programmatic dependent launch, bit 3: team mode builds the fit table with the main warp alone (A/B)
    // Rotation (rot_batches > 0): the state holds rot_batches batches of rot_envs envs each, back to back, and ONE launch executes
    // rot_steps consecutive "batch steps": batch step k advances every env of batch (rot_first + k) % rot_batches by n_steps.
    // A warp owns env index i of EVERY batch, so the steps of one record are always taken by the same warp, in order, and no
    // warp ever waits for another: the grid stays resident across what would otherwise be rot_steps dependent launches.
    int rot_batches, rot_steps, rot_first;
    long long rot_envs;
};

constexpr uint32_t STATUS_EXHAUSTED = 1u;   // pre-sampled trace ran out (the reference would raise, env.py:282)
constexpr uint32_t STATUS_QUIET = 2u;       // a fused agent's act()+apply would change nothing (see step_kernel)
constexpr uint32_t STATUS_OBS_STALE = 4u;   // the state changed since the env's row of a persistent observation buffer was stored
constexpr uint32_t STATUS_KEY_SHIFT = 8;    // bits 8..15: which (agent, tiebreak) established QUIET; 0 = any agent
constexpr uint32_t STATUS_KEY_MASK = 0xff00u;

__host__ __device__ static inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------------
// Per-warp env context: one base pointer into shared memory + the layout (kernel-parameter constant bank), so the
// dozen array pointers are recomputed from constants instead of living in (spilled) registers.
// ---------------------------------------------------------------------------------------------------
// every per-env array lives in shared memory: telling the compiler lets it emit LDS/STS instead of generic LD/ST
#define VMGYM_SMEM(p) (__builtin_assume(__isShared(p)), (p))

template <typename PT>
struct Env {
    unsigned char* base;               // the warp's shared-memory region: scratch arrays (L->sm_*)
    unsigned char* rec;                // the staged record (L->off_*): == base, or the second record buffer of a double-buffered kernel
    const DevLayout* L;
    const double* sz64;                // code -> k/100.0   (== np.around(u, 2), env.py:212-219)
    const float* sz32;                 // code -> (float)(k/100.0)   (env.py:296)
    const uint64_t* arr_cdf;           // arrival / service inverse-CDF thresholds (global tables)
    const uint32_t* arr_cdf32;         // top 32 bits of the arrival thresholds in shared memory (nullptr if the table is too long)
    const uint64_t* svc_cdf;
    const uint16_t* svc_bracket;       // 65-entry search brackets of the service table (shared memory) or nullptr
    int P, V, lane;
    int tune;                          // StepParams::use_bulk (team-mode experiment bits); only read by the team-mode kernel

    __device__ __forceinline__ double* cpu() const { return reinterpret_cast<double*>(VMGYM_SMEM(rec)); }           // env.py:190
    __device__ __forceinline__ double* mem() const { return reinterpret_cast<double*>(VMGYM_SMEM(rec) + L->off_mem); }
    __device__ __forceinline__ uint16_t* rem() const { return reinterpret_cast<uint16_t*>(VMGYM_SMEM(rec) + L->off_rem); }
    __device__ __forceinline__ PT* place() const { return reinterpret_cast<PT*>(VMGYM_SMEM(rec) + L->off_place); }
    __device__ __forceinline__ uint8_t* cpuc() const { return VMGYM_SMEM(rec) + L->off_cpuc; }                      // bit 7 = suspended
    __device__ __forceinline__ uint8_t* memc() const { return VMGYM_SMEM(rec) + L->off_memc; }
    // capacity codes of every PM in the agents' float32 view, kept consistent with cpu()/mem() by every update:
    // rcap[p] = kc | km << 8 with kc = max{k : (float)cpu[p] + sz32[k] <= 1.0f} (likewise km for memory)
    __device__ __forceinline__ uint16_t* rcap() const { return reinterpret_cast<uint16_t*>(VMGYM_SMEM(rec) + L->off_cap); }
    __device__ __forceinline__ vmgym_env_scalars* sc() const { return reinterpret_cast<vmgym_env_scalars*>(VMGYM_SMEM(rec) + L->off_scal); }
    // earliest step at which a running VM departs (absolute; 0xffffffff = no running VM; may be stale-early after a suspension):
    // lives behind the scalars and the parked Philox words.  While timestep < next_dep the service countdown has nothing to do.
    __device__ __forceinline__ uint32_t* next_dep() const
    {
        return reinterpret_cast<uint32_t*>(VMGYM_SMEM(rec) + L->off_scal + sizeof(vmgym_env_scalars) + 16);
    }
    // scratch (not part of the record)
    __device__ __forceinline__ float* cpu32() const { return reinterpret_cast<float*>(VMGYM_SMEM(base) + L->sm_cpu32); }  // agents' fp32 view
    __device__ __forceinline__ float* mem32() const { return reinterpret_cast<float*>(VMGYM_SMEM(base) + L->sm_mem32); }
    __device__ __forceinline__ uint16_t* act() const { return reinterpret_cast<uint16_t*>(VMGYM_SMEM(base) + L->sm_act); }
    __device__ __forceinline__ uint8_t* tmp() const { return VMGYM_SMEM(base) + L->sm_tmp; }
    __device__ __forceinline__ unsigned* fitm() const { return reinterpret_cast<unsigned*>(VMGYM_SMEM(base) + L->sm_fit); }
    __device__ __forceinline__ uint16_t* cap() const { return reinterpret_cast<uint16_t*>(VMGYM_SMEM(base) + L->sm_fit + 512); }
    __device__ __forceinline__ unsigned* prop() const { return reinterpret_cast<unsigned*>(VMGYM_SMEM(base) + L->sm_prop); }
    // team mode scratch (large shapes only, see "Team mode" below): command words, departure / candidate bitmaps per 32 slots
    __device__ __forceinline__ volatile int* ctl() const { return reinterpret_cast<volatile int*>(VMGYM_SMEM(base) + L->sm_team); }
    __device__ __forceinline__ unsigned* tmask() const { return reinterpret_cast<unsigned*>(VMGYM_SMEM(base) + L->sm_team + 16); }
    __device__ __forceinline__ unsigned* cmask() const
    {
        return reinterpret_cast<unsigned*>(VMGYM_SMEM(base) + L->sm_team + 16 + align_up(4 * ((L->Vp + 31) / 32), 16));
    }
    __device__ __forceinline__ unsigned* emask() const
    {
        return reinterpret_cast<unsigned*>(VMGYM_SMEM(base) + L->sm_team + 16 + 2 * align_up(4 * ((L->Vp + 31) / 32), 16));
    }
};

// Where the agent reads the slots it decides on: arrays of placements and size codes (+ float sizes as the agent
// sees them).  In the fused kernel these are the record's own arrays; in act_kernel they are built from the
// observation row.
template <typename PT>
struct AgentView {
    const PT* place;
    const uint8_t* cc;       // cpu size code (bit 7 may carry the suspended flag; masked on use)
    const uint8_t* mc;
    const float* c32;        // nullptr -> sz32[code]
    const float* m32;
    const float* sz32;
    __device__ __forceinline__ float cpu_size(int v) const { return c32 ? c32[v] : sz32[cc[v] & 0x7f]; }
    __device__ __forceinline__ float mem_size(int v) const { return m32 ? m32[v] : sz32[mc[v]]; }
};

__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}

// largest size code k in [0,100] with x + sz32[k] <= 1.0f (monotone in k because fp32 rounding is monotone).
// k0 = trunc((1 - x) * 100) is within one of the answer: 1 - x and the product carry < 1.3e-5 of absolute error in
// units of k, and the rounding of x + sz32[k] moves the threshold by < 1e-5, so the two can only disagree next to an
// integer — one probe up and one probe down settle it without a loop (checked against the definition over 2 M loads
// incl. every boundary neighbour, and by the capacity-cache parity checks in tests/test_env_cuda.py).
__device__ __forceinline__ int max_code(const float* sz32, float x)
{
    const int k0 = min(100, max(0, (int)((1.0f - x) * 100.0f)));
    const int up = (k0 < 100 && x + sz32[min(k0 + 1, 100)] <= 1.0f) ? 1 : 0;
    const int dn = (k0 > 0 && x + sz32[k0] > 1.0f) ? 1 : 0;
    return k0 + up - dn;
}

template <typename PT>
__device__ __forceinline__ void refresh_cap(const Env<PT>& e, int p)
{
    e.rcap()[p] = (uint16_t)(max_code(e.sz32, (float)e.cpu()[p]) | (max_code(e.sz32, (float)e.mem()[p]) << 8));
}

// fitm[c] = 1 + max{ mem-capacity code of PM p : cpu-capacity code of p >= c }, 0 if no PM takes cpu code c.
// A VM with size codes (c, m) fits on SOME PM iff m + 1 <= fitm[c] — an exact O(1) test that removes the hopeless
// waiting VMs (the large majority at saturation) from the sequential scan.
// Returns (max cpu capacity code) | (max mem capacity code) << 8 over all PMs, for the cheap byte pre-filter.
template <typename PT>
__device__ __forceinline__ unsigned rebuild_fit_table(const Env<PT>& e)
{
    const int lane = e.lane;
    unsigned* fitm = e.fitm();
    const uint16_t* cap = e.cap();
    reinterpret_cast<uint4*>(fitm)[lane] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    unsigned kcmax = 0, kmmax = 0;
    for (int p = lane; p < e.P; p += 32) {
        const unsigned w = cap[p];
        atomicMax(&fitm[w & 0xffu], (w >> 8) + 1u);
        kcmax = max(kcmax, w & 0xffu); kmmax = max(kmmax, w >> 8);
    }
    kcmax = __reduce_max_sync(FULL, kcmax); kmmax = __reduce_max_sync(FULL, kmmax);
    __syncwarp();
    uint4 q = reinterpret_cast<uint4*>(fitm)[lane];            // lane owns codes 4*lane .. 4*lane+3
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
    reinterpret_cast<uint4*>(fitm)[lane] = q;
    __syncwarp();
    return kcmax | (kmmax << 8);
}

// candidate bits of the 4 slots 4g..4g+3 (u8 placements): waiting VMs that fit on some PM
__device__ __forceinline__ unsigned cand_bits4(uint32_t pl4, uint32_t cc4, uint32_t mc4, uint32_t P4, const unsigned* fitm,
                                               unsigned kmax)
{
    cc4 &= 0x7f7f7f7fu;
    // waiting, and not larger than the largest free cpu / memory capacity of any PM (cheap necessary condition)
    const uint32_t w4 = __vcmpeq4(pl4, P4) & __vcmpleu4(cc4, (kmax & 0xffu) * 0x01010101u) & __vcmpleu4(mc4, (kmax >> 8) * 0x01010101u);
    unsigned c = 0;
    if (w4) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if ((w4 >> (8 * j)) & 1u) c |= (((mc4 >> (8 * j)) & 0xffu) + 1u <= fitm[(cc4 >> (8 * j)) & 0xffu]) ? (1u << j) : 0u;
    }
    return c;
}

// ---------------------------------------------------------------------------------------------------
// Team mode (large shapes: u16 placements, e.g. the synthetic 1000-PM / 3000-slot shape).  A 40 KB record leaves room
// for three envs per SM, so a warp per env runs at one warp per scheduler with nothing to hide its latencies behind.
// Here a CTA owns one env: warp 0 (the "main" warp) runs the same sequential step logic as the warp-per-env kernel and
// the other warps join it for the phases that are plain loops over all slots / PMs (observation row, float32 view +
// capacity copy, candidate pre-filter, service countdown).  Hand-off is two named barriers around a command word in
// shared memory; the results of a parallel phase that feed sequential logic (which slots finished, which 32-slot chunks
// hold a candidate) are bitmaps, consumed by the main warp in slot order, so every decision is taken in the reference's order.
// ---------------------------------------------------------------------------------------------------
enum { TEAM_END = 0, TEAM_OBS = 1, TEAM_PREP = 2, TEAM_FILTER = 3, TEAM_COUNTDOWN = 4, TEAM_FIT = 5, TEAM_SCAN = 6 };

__device__ __forceinline__ void team_bar(int id, int nth) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nth) : "memory"); }

template <typename PT>
__device__ __forceinline__ void team_phase(const Env<PT>& e, int cmd, int tid, int nth)
{
    const int P = e.P, V = e.V;
    if (cmd == TEAM_OBS) {
        // observation row (env.py:295-296); the destination pointer travels in ctl[2..3]
        const volatile int* ctl = e.ctl();
        float* o = reinterpret_cast<float*>(((unsigned long long)(unsigned)ctl[3] << 32) | (unsigned long long)(unsigned)ctl[2]);
        const PT* place = e.place();
        const uint8_t* cpuc = e.cpuc();
        const uint8_t* memc = e.memc();
        const double* cpu = e.cpu();
        const double* mem = e.mem();
        if (sizeof(PT) == 2 && (V & 3) == 0 && (P & 3) == 0) {
            // four slots / PMs per thread and 128-bit stores: the row and its five segments are 16-byte aligned
            float4* o4 = reinterpret_cast<float4*>(o);
            const uint2* pl4 = reinterpret_cast<const uint2*>(place);          // 4 x u16 placements
            const uint32_t* cc4 = reinterpret_cast<const uint32_t*>(cpuc);
            const uint32_t* mc4 = reinterpret_cast<const uint32_t*>(memc);
            const int vg = V >> 2, pg = P >> 2;
            for (int g = tid; g < vg; g += nth) {
                const uint2 a = pl4[g];
                const uint32_t c = cc4[g] & 0x7f7f7f7fu, m = mc4[g];
                o4[g] = make_float4((float)(a.x & 0xffffu), (float)(a.x >> 16), (float)(a.y & 0xffffu), (float)(a.y >> 16));
                o4[vg + g] = make_float4(e.sz32[c & 0xff], e.sz32[(c >> 8) & 0xff], e.sz32[(c >> 16) & 0xff], e.sz32[c >> 24]);
                o4[2 * vg + g] = make_float4(e.sz32[m & 0xff], e.sz32[(m >> 8) & 0xff], e.sz32[(m >> 16) & 0xff], e.sz32[m >> 24]);
            }
            const double2* c2 = reinterpret_cast<const double2*>(cpu);
            const double2* m2 = reinterpret_cast<const double2*>(mem);
            for (int g = tid; g < pg; g += nth) {
                const double2 a = c2[2 * g], b = c2[2 * g + 1], c = m2[2 * g], d = m2[2 * g + 1];
                o4[3 * vg + g] = make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
                o4[3 * vg + pg + g] = make_float4((float)c.x, (float)c.y, (float)d.x, (float)d.y);
            }
        } else {
            for (int v = tid; v < V; v += nth) { o[v] = (float)place[v]; o[V + v] = e.sz32[cpuc[v] & 0x7f]; o[2 * V + v] = e.sz32[memc[v]]; }
            for (int q = tid; q < P; q += nth) { o[3 * V + q] = (float)cpu[q]; o[3 * V + P + q] = (float)mem[q]; }
        }
    } else if (cmd == TEAM_PREP) {
        // the agent's float32 view of the PM loads (env.py:296), its local copy of the capacity codes, no proposals yet
        const double* cpu = e.cpu();
        const double* mem = e.mem();
        for (int q = tid; q < P; q += nth) { e.cpu32()[q] = (float)cpu[q]; e.mem32()[q] = (float)mem[q]; e.cap()[q] = e.rcap()[q]; }
        for (int c = tid; c < (V + 31) / 32; c += nth) e.prop()[c] = 0u;
    } else if (cmd == TEAM_FILTER) {
        // which 32-slot chunks hold a waiting VM that fits on some PM under the capacities at the start of act()
        // (capacities only shrink while the agent proposes, so this is a superset of the chunks worth visiting)
        const PT* place = e.place();
        const uint8_t* cc = e.cpuc();
        const uint8_t* mc = e.memc();
        const unsigned* fitm = e.fitm();
        unsigned* cm = e.cmask();
        if constexpr (sizeof(PT) == 2) {
            // four slots per thread (the arrays are padded to 16 slots with empty ones); a 32-slot chunk = 8 neighbouring threads
            const uint2* pl4 = reinterpret_cast<const uint2*>(place);
            const uint32_t* cc4 = reinterpret_cast<const uint32_t*>(cc);
            const uint32_t* mc4 = reinterpret_cast<const uint32_t*>(mc);
            const int nq = e.L->Vp >> 2;
            for (int q0 = (tid & ~31); q0 < nq; q0 += nth) {
                const int q = q0 + (tid & 31);
                unsigned w = 0;
                if (q < nq) {
                    const uint2 a = pl4[q];
                    const uint32_t c4 = cc4[q] & 0x7f7f7f7fu, m4 = mc4[q];
                    const unsigned pl[4] = {a.x & 0xffffu, a.x >> 16, a.y & 0xffffu, a.y >> 16};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const bool wait = pl[j] == (unsigned)P;
                        const unsigned f = fitm[wait ? ((c4 >> (8 * j)) & 0xffu) : 0u];
                        w |= (wait && ((m4 >> (8 * j)) & 0xffu) + 1u <= f) ? (1u << j) : 0u;
                    }
                    w <<= 4 * (q & 7);
                }
                w |= __shfl_xor_sync(FULL, w, 1); w |= __shfl_xor_sync(FULL, w, 2); w |= __shfl_xor_sync(FULL, w, 4);
                if ((q & 7) == 0 && q < nq) cm[q >> 3] = w;
            }
        } else {
            for (int v0 = (tid & ~31); v0 < V; v0 += nth) {
                const int v = v0 + (tid & 31);
                const bool cand = v < V && (int)place[v] == P && (unsigned)mc[v] + 1u <= fitm[cc[v] & 0x7f];
                const unsigned m = __ballot_sync(FULL, cand);
                if ((tid & 31) == 0) cm[v0 >> 5] = m;
            }
        }
    } else if (cmd == TEAM_COUNTDOWN) {
        // running VMs whose departure step is this step (ctl[1], env.py:245-249) are reported as a bitmap, the empty slots as a
        // second bitmap (for the admission of arrivals), and the earliest later departure as a distance in ctl[2] (atomicMin)
        const PT* place = e.place();
        const uint16_t* rem = e.rem();
        unsigned* tmk = e.tmask();
        unsigned* emk = e.emask();
        const uint32_t now16 = (uint32_t)e.ctl()[1] & 0xffffu;
        unsigned dmin = 0xffffffffu;
        if constexpr (sizeof(PT) == 2) {
            // four slots per thread, a 32-slot chunk = 8 neighbouring threads (see TEAM_FILTER); padding slots read as waiting here:
            // neither running nor empty (they must not receive admissions)
            const uint2* pl4 = reinterpret_cast<const uint2*>(place);
            const uint2* rm4 = reinterpret_cast<const uint2*>(rem);
            const int nq = e.L->Vp >> 2;
            for (int q0 = (tid & ~31); q0 < nq; q0 += nth) {
                const int q = q0 + (tid & 31);
                unsigned wt = 0, we = 0;
                if (q < nq) {
                    const uint2 a = pl4[q], r = rm4[q];
                    const unsigned pl[4] = {a.x & 0xffffu, a.x >> 16, a.y & 0xffffu, a.y >> 16};
                    const unsigned rr[4] = {r.x & 0xffffu, r.x >> 16, r.y & 0xffffu, r.y >> 16};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const bool running = pl[j] < (unsigned)P;
                        const uint32_t d = (rr[j] - now16) & 0xffffu;              // steps until this VM departs
                        const bool term = running && d == 0u;
                        if (running && !term) dmin = min(dmin, d);
                        wt |= term ? (1u << j) : 0u;
                        we |= (pl[j] == (unsigned)(P + 1) && 4 * q + j < V) ? (1u << j) : 0u;
                    }
                    wt <<= 4 * (q & 7); we <<= 4 * (q & 7);
                }
                wt |= __shfl_xor_sync(FULL, wt, 1); wt |= __shfl_xor_sync(FULL, wt, 2); wt |= __shfl_xor_sync(FULL, wt, 4);
                we |= __shfl_xor_sync(FULL, we, 1); we |= __shfl_xor_sync(FULL, we, 2); we |= __shfl_xor_sync(FULL, we, 4);
                if ((q & 7) == 0 && q < nq) { tmk[q >> 3] = wt; emk[q >> 3] = we; }
            }
        } else {
            for (int v0 = (tid & ~31); v0 < V; v0 += nth) {
                const int v = v0 + (tid & 31);
                const int pl = v < V ? (int)place[v] : P;
                bool term = false;
                if (pl < P) {
                    const uint32_t d = ((uint32_t)rem[v] - now16) & 0xffffu;       // steps until this VM departs
                    term = d == 0u;
                    if (!term) dmin = min(dmin, d);
                }
                const unsigned m = __ballot_sync(FULL, term), em = __ballot_sync(FULL, pl == P + 1);
                if ((tid & 31) == 0) { tmk[v0 >> 5] = m; emk[v0 >> 5] = em; }
            }
        }
        dmin = __reduce_min_sync(FULL, dmin);
        if ((tid & 31) == 0) atomicMin(reinterpret_cast<unsigned*>(const_cast<int*>(e.ctl())) + 2, dmin);
    } else if (cmd == TEAM_FIT) {
        // scatter pass of the fit table (see rebuild_fit_table): fitm[kc] = max(km + 1), plus the two global maxima in
        // ctl[1] / ctl[2] (zeroed by the main warp)
        unsigned* fitm = e.fitm();
        const uint16_t* cap = e.cap();
        if (tid < 128) fitm[tid] = 0u;
        team_bar(3, nth);
        unsigned kcmax = 0, kmmax = 0;
        for (int q = tid; q < P; q += nth) {
            const unsigned w = cap[q];
            atomicMax(&fitm[w & 0xffu], (w >> 8) + 1u);
            kcmax = max(kcmax, w & 0xffu); kmmax = max(kmmax, w >> 8);
        }
        kcmax = __reduce_max_sync(FULL, kcmax); kmmax = __reduce_max_sync(FULL, kmmax);
        if ((tid & 31) == 0) {
            atomicMax(reinterpret_cast<unsigned*>(const_cast<int*>(e.ctl())) + 1, kcmax);
            atomicMax(reinterpret_cast<unsigned*>(const_cast<int*>(e.ctl())) + 2, kmmax);
        }
    } else if (cmd == TEAM_SCAN) {
        // the PM scan of one waiting VM (sizes in ctl[2] / ctl[3] as float bits, agent in ctl[1]) on the agent's float32
        // view: first-fit -> lowest fitting index; best-fit -> largest cpu+memory among the fitting PMs, ties -> highest
        // index.  Every warp leaves (key, index + 1) in tmp[]; the main warp combines them.
        const volatile int* ctl = e.ctl();
        const int agent = ctl[1];
        const float c32 = __int_as_float(ctl[2]), m32 = __int_as_float(ctl[3]);
        const float* cpu32 = e.cpu32();
        const float* mem32 = e.mem32();
        unsigned* res = reinterpret_cast<unsigned*>(e.tmp());
        unsigned bestk = 0, bestp = 0;
        if (agent == VMGYM_AGENT_FIRSTFIT) {
            // key = ~index so that the maximum key is the lowest index
            for (int q = tid; q < P; q += nth) {
                const bool fit = (cpu32[q] + c32 <= 1.0f) && (mem32[q] + m32 <= 1.0f);
                if (fit && bestk == 0u) { bestk = ~(unsigned)q; bestp = (unsigned)q + 1u; }
            }
        } else {
            for (int q = tid; q < P; q += nth) {
                const bool fit = (cpu32[q] + c32 <= 1.0f) && (mem32[q] + m32 <= 1.0f);
                const unsigned kb = __float_as_uint(cpu32[q] + mem32[q]) + 1u;   // keys >= 0: bits order like values
                if (fit && kb >= bestk) { bestk = kb; bestp = (unsigned)q + 1u; }
            }
        }
        const unsigned gk = __reduce_max_sync(FULL, bestk);
        const unsigned gi = __reduce_max_sync(FULL, bestk == gk ? bestp : 0u);
        if ((tid & 31) == 0) { res[2 * (tid >> 5)] = gk; res[2 * (tid >> 5) + 1] = gi; }
    }
}

// main warp: run one phase with the whole team (all 32 lanes call this, converged)
template <typename PT>
__device__ __forceinline__ void team_run(const Env<PT>& e, int cmd, int nth)
{
    __syncwarp();
    if (e.lane == 0) e.ctl()[0] = cmd;
    team_bar(1, nth);
    if (cmd != TEAM_END) team_phase(e, cmd, e.lane, nth);
    team_bar(2, nth);
}

// helper warps: serve the main warp's phases of one env
template <typename PT>
__device__ __forceinline__ void team_serve(const Env<PT>& e, int tid, int nth)
{
    for (;;) {
        team_bar(1, nth);
        const int cmd = e.ctl()[0];
        if (cmd != TEAM_END) {
            team_phase(e, cmd, tid, nth);
            fence_proxy_async();          // this thread's shared-memory writes vs the record's bulk-async store
        }
        team_bar(2, nth);
        if (cmd == TEAM_END) break;
    }
}

// the fit table (rebuild_fit_table) built by the whole team: scatter pass in parallel, suffix maximum by the main warp
template <typename PT>
__device__ __forceinline__ unsigned rebuild_fit_table_team(const Env<PT>& e, int nth)
{
    const int lane = e.lane;
    unsigned* fitm = e.fitm();
    __syncwarp();
    if (lane == 0) { e.ctl()[1] = 0; e.ctl()[2] = 0; }
    team_run(e, TEAM_FIT, nth);
    const unsigned kcmax = (unsigned)e.ctl()[1], kmmax = (unsigned)e.ctl()[2];
    uint4 q = reinterpret_cast<uint4*>(fitm)[lane];            // lane owns codes 4*lane .. 4*lane+3
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
    reinterpret_cast<uint4*>(fitm)[lane] = q;
    __syncwarp();
    return kcmax | (kmmax << 8);
}

// one waiting VM's PM scan by the whole team (see TEAM_SCAN); returns the chosen PM or -1
template <typename PT>
__device__ __forceinline__ int team_scan(const Env<PT>& e, int agent, float c32, float m32, int nth)
{
    __syncwarp();
    if (e.lane == 0) { e.ctl()[1] = agent; e.ctl()[2] = __float_as_int(c32); e.ctl()[3] = __float_as_int(m32); }
    team_run(e, TEAM_SCAN, nth);
    const unsigned* res = reinterpret_cast<const unsigned*>(e.tmp());
    const int nw = nth >> 5;
    const unsigned k = e.lane < nw ? res[2 * e.lane] : 0u, ix = e.lane < nw ? res[2 * e.lane + 1] : 0u;
    const unsigned gk = __reduce_max_sync(FULL, k);
    const int found = (int)__reduce_max_sync(FULL, k == gk ? ix : 0u) - 1;      // gk == 0: every ix is 0 -> -1
    __syncwarp();
    return found;
}

// ---------------------------------------------------------------------------------------------------
// Heuristic agents on the float32 view (firstfit.py:21-38, bestfit.py:21-40).  Lanes own PMs p = lane + 32 i.
// For every waiting VM, in slot order: first-fit takes the lowest-index PM that fits, best-fit the fitting PM with
// the largest cpu+memory (ties: see `tiebreak`); the local float32 loads are updated like the reference does
// (first-fit: cpu only, firstfit.py:36).  Proposals are recorded as act[v] = pm plus a bit in prop[v / 32].
// Returns the number of proposals.
// ---------------------------------------------------------------------------------------------------
template <typename PT, bool TM = false>
__device__ __forceinline__ int agent_act(const Env<PT>& e, const AgentView<PT>& av, int agent, int tiebreak, bool have_rcap, int nth = 32)
{
    const int P = e.P, V = e.V, lane = e.lane;
    float* cpu32 = e.cpu32();
    float* mem32 = e.mem32();
    uint16_t* cap = e.cap();
    const unsigned* fitm = e.fitm();
    int n_found = 0;
    // local capacity codes: the record's (fused kernel) or computed from the observed loads (act_kernel); the agent's
    // own proposals then shrink this local copy only, like the reference's local cpu/memory arrays
    if constexpr (TM) {
        team_run(e, TEAM_PREP, nth);      // float32 view, cap[] = rcap[] (the fused kernel always has the record's codes), prop[] = 0
    } else {
        if (have_rcap) { for (int p = lane; p < P; p += 32) cap[p] = e.rcap()[p]; }
        else { for (int p = lane; p < P; p += 32) cap[p] = (uint16_t)(max_code(e.sz32, cpu32[p]) | (max_code(e.sz32, mem32[p]) << 8)); }
        for (int c = lane; c < (V + 31) / 32; c += 32) e.prop()[c] = 0u;
    }
    const bool team_fit = TM && (e.tune & 8) == 0;       // the team builds the fit table (use_bulk bit 3: main warp alone, for A/B runs)
    unsigned kmax;
    {
        PROF_SPAN_BEGIN();
        kmax = team_fit ? rebuild_fit_table_team(e, nth) : rebuild_fit_table(e);
        PROF_SPAN_END(13, -1);
    }
    bool dirty = false;                                  // proposals since the table was built (it over-states the free capacity)
    bool refiltered = false;                             // team mode: the chunk marks were refreshed inside visit()

    // slots are visited in groups: 4 per lane (128 per pass) for byte placements, 1 per lane otherwise
    constexpr int SPL = sizeof(PT) == 1 ? 4 : 1;
    const uint32_t P4 = (uint32_t)P * 0x01010101u;
    const int n_units = (V + SPL - 1) / SPL;
    auto visit = [&](const int u0) {
        const int u = u0 + lane;
        unsigned cb = 0;                      // candidate bits of this lane's SPL slots
        if (u < n_units) {
            if (SPL == 4) {
                cb = cand_bits4(reinterpret_cast<const uint32_t*>(av.place)[u], reinterpret_cast<const uint32_t*>(av.cc)[u],
                                reinterpret_cast<const uint32_t*>(av.mc)[u], P4, fitm, kmax);
            } else {
                cb = ((int)av.place[u] == P && (unsigned)av.mc[u] + 1u <= fitm[av.cc[u] & 0x7f]) ? 1u : 0u;
            }
        }
        unsigned m = __ballot_sync(FULL, cb != 0);
        while (m) {
            const int b = __ffs(m) - 1;
            unsigned bits = __shfl_sync(FULL, cb, b);
            bool placed_any = false;
            while (bits) {
                const int j = __ffs(bits) - 1;
                bits &= bits - 1;
                const int vv = SPL * (u0 + b) + j;
                const float c32 = av.cpu_size(vv), m32 = av.mem_size(vv);
                int found = -1;
                if (TM && tiebreak != VMGYM_TIE_NUMPY_INTROSORT) {
                    // team mode: all warps scan (the numpy-introsort tie rule stays on the main warp, below)
                    {
                        PROF_SPAN_BEGIN();
                        found = team_scan(e, agent, c32, m32, nth);
                        PROF_SPAN_END(11, -1);
                    }
                    if (found >= 0 && lane == 0) {
                        const float nc = cpu32[found] + c32;
                        cpu32[found] = nc;                              // firstfit.py:36 / bestfit.py:37-38
                        unsigned cw = (cap[found] & 0xff00u) | (unsigned)max_code(e.sz32, nc);
                        if (agent != VMGYM_AGENT_FIRSTFIT) {
                            const float nm = mem32[found] + m32;
                            mem32[found] = nm;
                            cw = (cw & 0xffu) | ((unsigned)max_code(e.sz32, nm) << 8);
                        }
                        cap[found] = (uint16_t)cw;
                    }
                } else if (agent == VMGYM_AGENT_FIRSTFIT) {
                    for (int i0 = 0; i0 < P; i0 += 32) {
                        const int p = i0 + lane;
                        const bool fit = p < P && (cpu32[p] + c32 <= 1.0f) && (mem32[p] + m32 <= 1.0f);
                        const unsigned bb = __ballot_sync(FULL, fit);
                        if (bb) { found = i0 + __ffs(bb) - 1; break; }
                    }
                    if (found >= 0 && lane == (found & 31)) {
                        const float nc = cpu32[found] + c32;        // firstfit.py:36 — only the local cpu is updated
                        cpu32[found] = nc;
                        cap[found] = (uint16_t)((cap[found] & 0xff00u) | (unsigned)max_code(e.sz32, nc));
                    }
                } else {
                    // best-fit: first fitting PM in descending (cpu+memory) order (bestfit.py:33-39)
                    unsigned bestk = 0;
                    int bestp = -1;
                    for (int p = lane; p < P; p += 32) {
                        const bool fit = (cpu32[p] + c32 <= 1.0f) && (mem32[p] + m32 <= 1.0f);
                        const unsigned kb = __float_as_uint(cpu32[p] + mem32[p]) + 1u;   // keys >= 0: bits order like values
                        if (fit && kb >= bestk) { bestk = kb; bestp = p; }
                    }
                    const unsigned gk = __reduce_max_sync(FULL, bestk);
                    if (gk != 0) {
                        found = (int)__reduce_max_sync(FULL, (unsigned)((bestk == gk ? bestp : -1) + 1)) - 1;  // ties -> highest index
                        if (tiebreak == VMGYM_TIE_NUMPY_INTROSORT) {
                            int cnt = 0;
                            for (int p = lane; p < P; p += 32) {
                                const bool fit = (cpu32[p] + c32 <= 1.0f) && (mem32[p] + m32 <= 1.0f);
                                cnt += (fit && __float_as_uint(cpu32[p] + mem32[p]) + 1u == gk);
                            }
                            cnt = __reduce_add_sync(FULL, cnt);
                            if (cnt >= 2) {
                                // several fitting PMs share the maximal key: numpy's unstable default argsort decides
                                float* keys = reinterpret_cast<float*>(e.tmp());
                                uint16_t* perm = reinterpret_cast<uint16_t*>(e.tmp() + 4 * ((P + 1) & ~1));
                                for (int p = lane; p < P; p += 32) keys[p] = cpu32[p] + mem32[p];
                                __syncwarp();
                                int pick = -1;
                                if (lane == 0) {
                                    introsort_argsort(keys, perm, P);
                                    for (int i = P - 1; i >= 0; i--) {
                                        const int p = perm[i];
                                        if ((cpu32[p] + c32 <= 1.0f) && (mem32[p] + m32 <= 1.0f)) { pick = p; break; }
                                    }
                                }
                                found = __shfl_sync(FULL, pick, 0);
                                __syncwarp();
                            }
                        }
                        if (lane == (found & 31)) {
                            const float nc = cpu32[found] + c32, nm = mem32[found] + m32;   // bestfit.py:37-38
                            cpu32[found] = nc;
                            mem32[found] = nm;
                            cap[found] = (uint16_t)(max_code(e.sz32, nc) | (max_code(e.sz32, nm) << 8));
                        }
                    }
                }
                if (found >= 0) {
                    n_found++;
                    if (lane == 0) { e.act()[vv] = (uint16_t)found; e.prop()[vv >> 5] |= 1u << (vv & 31); }
                    __syncwarp();
                    // capacities shrank, so the fit table now over-states what fits: it stays a valid pre-filter (a superset of the
                    // candidates — the PM scan itself is exact) and is only rebuilt once a scan comes back empty-handed
                    dirty = true;
                } else if (dirty) {
                    PROF_SPAN_BEGIN();
                    kmax = team_fit ? rebuild_fit_table_team(e, nth) : rebuild_fit_table(e);
                    PROF_SPAN_END(13, -1);
                    dirty = false;
                    placed_any = true;                 // the not-yet-visited lanes' candidates are re-tested against the fresh table
                    if constexpr (TM && SPL == 1) {
                        // The chunk marks date from the start of act(): every marked chunk would still get a (fruitless) visit although
                        // the capacity that made its VMs candidates is gone (at saturation ~20 visits per step for the one or two PMs a
                        // departure freed).  The fresh table is exact, so the team marks the chunks again (four slots per thread: cheaper
                        // than the visits it saves).
                        team_run(e, TEAM_FILTER, nth);
                        refiltered = true;
                    }
                    if (bits) {
                        unsigned nb;
                        if (SPL == 4) {
                            const int ub = u0 + b;
                            nb = cand_bits4(reinterpret_cast<const uint32_t*>(av.place)[ub], reinterpret_cast<const uint32_t*>(av.cc)[ub],
                                            reinterpret_cast<const uint32_t*>(av.mc)[ub], P4, fitm, kmax);
                        } else nb = 0;
                        bits &= nb;
                    }
                }
            }
            m &= m - 1;
            if (placed_any && m) {
                // re-test the not-yet-visited lanes' candidates against the shrunken capacities
                if (u < n_units && lane > b) {
                    if (SPL == 4)
                        cb = cand_bits4(reinterpret_cast<const uint32_t*>(av.place)[u], reinterpret_cast<const uint32_t*>(av.cc)[u],
                                        reinterpret_cast<const uint32_t*>(av.mc)[u], P4, fitm, kmax);
                    else
                        cb = ((int)av.place[u] == P && (unsigned)av.mc[u] + 1u <= fitm[av.cc[u] & 0x7f]) ? 1u : 0u;
                }
                m &= __ballot_sync(FULL, cb != 0);
            }
        }
    };
    if constexpr (TM && SPL == 1) {
        // the team marks the 32-slot chunks that hold a candidate; the main warp visits only those, in slot order, and
        // re-tests their slots against the current capacities (visit() evaluates the chunk afresh)
        {
            PROF_SPAN_BEGIN();
            team_run(e, TEAM_FILTER, nth);
            PROF_SPAN_END(14, -1);
        }
        const unsigned* cm = e.cmask();
        const int n_chunks = (V + 31) / 32;
        for (int cb0 = 0; cb0 < n_chunks; cb0 += 32) {
            unsigned nz = __ballot_sync(FULL, cb0 + lane < n_chunks && cm[cb0 + lane] != 0u);
            while (nz) {
                const int c = cb0 + __ffs(nz) - 1;
                nz &= nz - 1;
                PROF_SPAN_BEGIN();
                visit(32 * c);
                PROF_SPAN_END(15, 12);
                if (refiltered) {
                    // marks only ever lose bits (capacities shrink during act()); later groups of 32 chunks read the fresh marks anyway
                    nz &= __ballot_sync(FULL, cb0 + lane < n_chunks && cm[cb0 + lane] != 0u);
                    refiltered = false;
                }
            }
        }
    } else {
        for (int u0 = 0; u0 < n_units; u0 += 32) visit(u0);
    }
    __syncwarp();
    return n_found;
}

// ---------------------------------------------------------------------------------------------------
// Per-VM episode statistics (src/record.py:34-96).  Record rebuilds, for every VM that ever occupied a slot, the list
// of its post-step placement samples from arrival to departure and derives
//   pending rate  = around((first running sample index + 1) / len, 3)        (1.0 if never placed)
//   slowdown rate = around(#WAIT samples after the first placement / (len - index - 1), 3)   (placed VMs only)
//   lifetime      = len - index - 1                                            (0 if never placed)
// With t_a / t_p / t_d the steps of arrival, first placement and departure: index = t_p - t_a (arrivals follow the
// apply loop, so index >= 1), len = t_d - t_a, and the WAIT samples after the first placement are the lengths of the
// suspension intervals.  Four clocks per slot therefore replace the V x T sample matrix; rates are binned by
// rint(1000 * rate), which is exactly what np.around(., 3) keeps.
// ---------------------------------------------------------------------------------------------------
struct VmStat {
    uint32_t* slots;      // this env's [V][4]
    uint32_t* hist;       // this env's [2][VMGYM_VMSTAT_BINS]
    unsigned long long* totals;   // this env's [4]
};

// the VM in `s` ends its sample list after `len` samples; `wait_extra`: WAIT samples of a still-open suspension
__device__ __forceinline__ void vmstat_close(const uint32_t* s, uint32_t len, uint32_t wait_extra, uint32_t* hist,
                                             unsigned long long* totals)
{
    totals[0] += 1ull;
    if (s[1] != 0u) {
        const uint32_t idx = s[1] - s[0];                                     // allocated_at (record.py:59)
        const uint32_t life = len - idx - 1u;
        const int kp = (int)rint(((double)idx + 1.0) / (double)len * 1000.0);
        const uint32_t wait = s[2] + wait_extra;
        const int ks = life == 0u ? 0 : (int)rint((double)wait / (double)life * 1000.0);
        hist[min(max(kp, 0), VMGYM_VMSTAT_BINS - 1)] += 1u;
        hist[VMGYM_VMSTAT_BINS + min(max(ks, 0), VMGYM_VMSTAT_BINS - 1)] += 1u;
        totals[1] += 1ull;
        totals[2] += (unsigned long long)life;
    } else {
        hist[1000] += 1u;                                                      // never placed: pending rate 1.0
    }
}

// ---------------------------------------------------------------------------------------------------
// Rewards `ut` (env.py:151-152) and `kl` (env.py:125-150, kl_divergence :8-17).  All reductions use numpy's
// summation order so the fp64 values (and the exact-zero variance tests) are those of the reference.
// ---------------------------------------------------------------------------------------------------
static __device__ __noinline__ double reward_ut(const double* cpu, const double* mem, int P, double beta)
{
    const SumSrc sc{cpu, nullptr, nullptr, 0.0, 0}, sm{mem, nullptr, nullptr, 0.0, 0};
    return beta * np_sum(sc, P) + (1 - beta) * np_sum(sm, P);
}

static __device__ __noinline__ double reward_kl(const double* cpu, const double* mem, int P, const uint8_t* ex_cc,
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

__device__ __forceinline__ Philox4 philox_dev(uint32_t c0, uint32_t c2, uint32_t k0, uint32_t k1)
{
    return philox4x32_10(c0, 0u, c2, 0u, k0, k1);
}

__device__ __forceinline__ int cdf_search(const uint64_t* cdf, int len, uint64_t u)
{
    int lo = 0, hi = len;            // first i with cdf[i] > u
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
    return min(lo, len - 1);
}

// `kl` reward of one env: compact vm_cpu[existing], vm_memory[existing] in slot order (the reference's boolean
// indexing, env.py:113,129-130) and evaluate reward_kl.
template <typename PT>
__device__ __noinline__ double reward_kl_env(const Env<PT> e, int arrived, int cap_target)
{
    const int P = e.P, V = e.V, lane = e.lane;
    const PT* place = e.place();
    uint8_t* ex_cc = VMGYM_SMEM(e.base) + e.L->sm_kl;
    uint8_t* ex_mc = ex_cc + ((V + 15) & ~15);
    int pos0 = 0;
    for (int c0 = 0; c0 < V; c0 += 32) {
        const int v = c0 + lane;
        const bool ex = v < V && (int)place[v] <= P;
        const unsigned mx = __ballot_sync(FULL, ex);
        if (ex) {
            const int pos = pos0 + __popc(mx & ((1u << lane) - 1u));
            ex_cc[pos] = e.cpuc()[v] & 0x7f;
            ex_mc[pos] = e.memc()[v];
        }
        pos0 += __popc(mx);
    }
    __syncwarp();
    return reward_kl(e.cpu(), e.mem(), P, ex_cc, ex_mc, arrived, e.sz64, cap_target);
}

// same result as cdf_search, started from the bracket of the top 6 bits of u: bracket[b] = #{i : cdf[i] <= b << 58}
__device__ __forceinline__ int cdf_search_bracketed(const uint64_t* cdf, int len, const uint16_t* bracket, uint64_t u)
{
    const int b = (int)(u >> 58);
    int lo = bracket[b], hi = bracket[b + 1];
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
    return min(lo, len - 1);
}

struct StepResult { double reward; int terminated; int rejected; int waiting, arrived; int changed; };

// running sums for the eval summary (record.py:98-134, exp_performance.py:104-113): drop rate, waiting ratio,
// per-step mean / population variance of PM cpu and memory, rejected actions, step count
template <typename PT>
__device__ __noinline__ void stats_update(const Env<PT> e, const StepResult res, double* st_acc, int cap_target)
{
    const int P = e.P, lane = e.lane;
    const double* cpu = e.cpu();
    const double* mem = e.mem();
    double sc_ = 0, sm_ = 0;
    for (int q = lane; q < P; q += 32) { sc_ += cpu[q]; sm_ += mem[q]; }
    const double mc = warp_sum(sc_) / P, mm = warp_sum(sm_) / P;
    double vc = 0, vm = 0;
    for (int q = lane; q < P; q += 32) {
        const double dc = cpu[q] - mc, dm = mem[q] - mm;
        vc += dc * dc; vm += dm * dm;
    }
    vc = warp_sum(vc) / P; vm = warp_sum(vm) / P;
    // what Record additionally averages over the episode (record.py:127-133): target means (env.py:116-121: size sums of
    // the existing VMs over P, capped), and the rank of the placement matrix = number of PMs hosting a VM (env.py:319-325)
    int used = 0;
    for (int q = lane; q < P; q += 32) used += cpu[q] != 0.0 ? 1 : 0;
    used = __reduce_add_sync(FULL, used);
    unsigned csum = 0, msum = 0;
    const PT* place = e.place();
    const uint8_t* cc = e.cpuc();
    const uint8_t* mcode = e.memc();
    for (int v = lane; v < e.V; v += 32)
        if ((int)place[v] <= P) { csum += cc[v] & 0x7f; msum += mcode[v]; }
    csum = __reduce_add_sync(FULL, csum); msum = __reduce_add_sync(FULL, msum);
    if (lane == 0) {
        const int tot = e.sc()->total_requests;
        st_acc[0] += tot ? (double)e.sc()->dropped_requests / (double)tot : 0.0;
        st_acc[1] += res.arrived ? (double)res.waiting / (double)res.arrived : 0.0;
        st_acc[2] += mc; st_acc[3] += vc; st_acc[4] += mm; st_acc[5] += vm;
        st_acc[6] += res.rejected; st_acc[7] += 1;
        st_acc[8] += mc * mc; st_acc[9] += mm * mm;                 // with [3] / [5]: E[x^2] for the global std (np.std over T x P)
        double tc = (double)csum / 100.0 / P, tm = (double)msum / 100.0 / P;
        if (cap_target && tc > 1.0) tc = 1.0;
        if (cap_target && tm > 1.0) tm = 1.0;
        st_acc[10] += tc; st_acc[11] += tm;
        st_acc[12] += used;
    }
}

// ---------------------------------------------------------------------------------------------------
// One env.step on the shared-memory record.  The action vector is given as act[v] for the slots whose bit is
// set in prop[] (slots whose action differs from their placement); every other action is a valid no-op
// (validate(), env.py:36-37).  The per-env counters n_waiting / n_empty are maintained incrementally so the
// common quiet step (nothing placed, nothing departs, nothing admitted) costs only the service countdown, one
// arrival draw and the outputs.
// ---------------------------------------------------------------------------------------------------
template <typename PT, int REWARD_CT, int MODE_CT, bool VMSTAT, bool TM = false>
__device__ __forceinline__ StepResult env_step(const Env<PT>& e, const StepParams& p, long long env_id, uint8_t* valid_g,
                                               bool have_actions, int nth = 32)
{
    // per-VM statistics are compiled into the generic instantiations only (the specialised throughput kernels skip them)
    const bool vmstat = VMSTAT && p.out.d_vm_slots != nullptr;
    VmStat vs;
    vs.slots = vmstat ? p.out.d_vm_slots + env_id * (long long)e.V * 4 : nullptr;
    vs.hist = vmstat ? p.out.d_vm_hist + env_id * 2ll * VMGYM_VMSTAT_BINS : nullptr;
    vs.totals = vmstat ? reinterpret_cast<unsigned long long*>(p.out.d_vm_totals) + env_id * 4 : nullptr;
    const uint32_t tnow = (uint32_t)e.sc()->timestep;                  // the step being executed (1-based, env.py:101)
    const int reward_fn = REWARD_CT ? REWARD_CT : p.reward_fn;            // compile-time in the specialised kernels
    const int trace_mode = MODE_CT >= 0 ? MODE_CT : p.tr.mode;
    const int P = e.P, V = e.V, lane = e.lane;
    vmgym_env_scalars* sc = e.sc();
    double* cpu = e.cpu();
    double* mem = e.mem();
    PT* place = e.place();
    uint8_t* cpuc = e.cpuc();
    uint8_t* memc = e.memc();
    uint16_t* rem = e.rem();
    int n_place = 0, n_susp = 0, rejected = 0;

    PROF_T0();
    // ---- 1. apply actions in VM-index order, each seeing earlier updates (env.py:69-87, validate :35-42) ----
    if (have_actions) {
        const uint16_t* act = e.act();
        const unsigned* prop = e.prop();
        auto apply_chunk = [&](const int c0) {
            const unsigned pm = prop[c0 >> 5];
            unsigned m = pm, okbits = 0;
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const int vv = c0 + b, a = (int)act[vv], cv = (int)place[vv];
                bool ok = false;
                if (cv == P) {                                   // waiting VM: place iff it fits in fp64 (:38-39,55-56)
                    if ((unsigned)a < (unsigned)P) {
                        const double nc = cpu[a] + e.sz64[cpuc[vv] & 0x7f];
                        const double nm = mem[a] + e.sz64[memc[vv]];
                        if (nc <= 1.0 && nm <= 1.0) {
                            ok = true;
                            n_place++;
                            __syncwarp();
                            if (lane == 0) {                                                                   // :82-85
                                cpu[a] = nc; mem[a] = nm; place[vv] = (PT)a; cpuc[vv] &= 0x7f;
                                // a VM placed in step t with r steps left ticks in step t already (:245-247 run after the apply
                                // loop) and departs in step t + r - 1: running slots hold that step (mod 2^16) instead of a counter
                                const uint32_t fin = tnow + (uint32_t)rem[vv] - 1u;
                                rem[vv] = (uint16_t)fin;
                                if (fin < *e.next_dep()) *e.next_dep() = fin;
                                e.rcap()[a] = (uint16_t)(max_code(e.sz32, (float)nc) | (max_code(e.sz32, (float)nm) << 8));
                                if (vmstat) {
                                    uint32_t* s = vs.slots + vv * 4;
                                    if (s[1] == 0u) s[1] = tnow;                               // first placement
                                    else { s[2] += tnow - s[3]; s[3] = 0u; }                   // end of a suspension
                                }
                            }
                        }
                    }
                } else if (cv < P) {                             // running VM: only suspend is legal (:40-41,78-81)
                    if (a == P) {
                        ok = true;
                        n_susp++;
                        const double nc = cpu[cv] - e.sz64[cpuc[vv] & 0x7f];
                        const double nm = mem[cv] - e.sz64[memc[vv]];
                        __syncwarp();
                        if (lane == 0) {
                            cpu[cv] = nc; mem[cv] = nm; place[vv] = (PT)P; cpuc[vv] |= 0x80;
                            rem[vv] = (uint16_t)((((uint32_t)rem[vv] - tnow) & 0xffffu) + 1u);     // back to "steps left" (next_dep may go stale-early)
                            e.rcap()[cv] = (uint16_t)(max_code(e.sz32, (float)nc) | (max_code(e.sz32, (float)nm) << 8));
                            if (vmstat) vs.slots[vv * 4 + 3] = tnow;                           // WAIT samples start at this step
                        }
                    }
                }
                __syncwarp();
                okbits |= ok ? (1u << b) : 0u;
            }
            rejected += __popc(pm & ~okbits);
            if (valid_g && c0 + lane < V) valid_g[c0 + lane] = ((pm & ~okbits) >> lane) & 1u ? 0 : 1;
        };
        if (TM && !valid_g) {
            // large shapes: only the 32-slot chunks that carry a proposal, in slot order
            const int n_chunks = (V + 31) / 32;
            for (int cb0 = 0; cb0 < n_chunks; cb0 += 32) {
                unsigned nz = __ballot_sync(FULL, cb0 + lane < n_chunks && prop[cb0 + lane] != 0u);
                while (nz) {
                    const int c = cb0 + __ffs(nz) - 1;
                    nz &= nz - 1;
                    apply_chunk(32 * c);
                }
            }
        } else {
            for (int c0 = 0; c0 < V; c0 += 32) apply_chunk(c0);
        }
    } else if (valid_g) {
        for (int v = lane; v < V; v += 32) valid_g[v] = 1;
    }

    PROF_ADD(6);                                               // 6: apply
    // ---- 5a. the arrival draw (_accept_vm_requests, env.py:272) does not depend on the state: taken first so that a step in which
    // nothing departs and nothing arrives skips the scans over the slots altogether ----
    int n_arr = 0;
    const vmgym_trace& tr = p.tr;
    const uint32_t arrival_pos = sc->arrival_pos, admission_pos = sc->admission_pos;
    const uint32_t k0 = (uint32_t)sc->seed, k1 = (uint32_t)(sc->seed >> 32);
    int exhausted = 0;
    if (trace_mode == VMGYM_TRACE_PRESAMPLED) {
        if ((long long)arrival_pos < tr.arrivals_len) n_arr = tr.d_arrivals[env_id * tr.arrivals_len + arrival_pos];
        else exhausted = 1;
    } else {
        // one Philox call serves four consecutive steps: step t takes 32-bit word (t & 3) of block (t >> 2) and inverts
        // the Poisson CDF on the top 32 bits of the thresholds (resolution 2^-32).  Words 1..3 of the block are parked in
        // the record's padding behind the scalars (tag = block index + 1) so three steps out of four skip the 10 rounds.
        const uint32_t sub = arrival_pos & 3u, blk = arrival_pos >> 2;
        uint32_t* park = reinterpret_cast<uint32_t*>(e.sc() + 1);             // 16 B: tag, y, z, w
        uint32_t u32;
        if (sub != 0u && park[0] == blk + 1u) {
            u32 = park[sub];
        } else {
            const Philox4 r = philox_dev(blk, 1u, k0, k1);
            u32 = sub == 0 ? r.x : (sub == 1 ? r.y : (sub == 2 ? r.z : r.w));
            __syncwarp();
            if (lane == 0) { park[0] = blk + 1u; park[1] = r.y; park[2] = r.z; park[3] = r.w; }
        }
        int lo;
        if (e.arr_cdf32 && tr.arrival_cdf_len <= 64) {
            // #{i : cdf[i] <= u} with one table entry per lane (the table is sorted, so this is the search result)
            lo = __popc(__ballot_sync(FULL, lane < tr.arrival_cdf_len && e.arr_cdf32[lane] <= u32));
            if (tr.arrival_cdf_len > 32)
                lo += __popc(__ballot_sync(FULL, lane + 32 < tr.arrival_cdf_len && e.arr_cdf32[lane + 32] <= u32));
        } else {
            int hi = tr.arrival_cdf_len;
            lo = 0;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if ((uint32_t)(e.arr_cdf[mid] >> 32) <= u32) lo = mid + 1; else hi = mid; }
        }
        n_arr = tr.arrival_kmin + min(lo, tr.arrival_cdf_len - 1);
    }

    PROF_ADD(7);                                               // 7: arrival draw
    // ---- 2+3. service countdown and departures in VM-index order (_run_vms, env.py:244-265) ----
    // Running slots hold their departure step (see the apply loop), so "remaining -= 1; terminate at 0" is "terminate the VMs whose
    // departure step is this step" — and the earliest departure step of the env is kept in next_dep: the scan over the slots runs
    // only in steps in which something (possibly) departs.  The scan also renews next_dep.
    int served = 0;
    bool need_full_refresh = false;
    int freed0 = -1, freed1 = -1, freed2 = -1, freed3 = -1;    // first slots freed by this step's departures (slot order)
    const uint32_t now16 = tnow & 0xffffu;
    const bool scan = tnow >= *e.next_dep();
    // team mode also takes its bitmap of empty slots from the scan: needed when arrivals may be admitted into slots that were
    // already empty before this step
    const bool team_scan_needed = TM && sizeof(PT) != 1 && (scan || (n_arr > 0 && sc->n_empty > 0));
    unsigned dmin = 0xffffffffu;                               // distance to the earliest departure after this step
    if (sizeof(PT) == 1) {
        if (scan) {
        // 4 slots per lane: placement bytes as one u32, departure steps as 4 x u16 (padding slots are empty)
        const uint32_t P4 = (uint32_t)P * 0x01010101u;
        const uint32_t now2 = now16 * 0x00010001u;
        const uint32_t* pl4 = reinterpret_cast<const uint32_t*>(place);
        const uint2* rem4 = reinterpret_cast<const uint2*>(rem);
        const int groups = (V + 3) >> 2;
        for (int g0 = 0; g0 < groups; g0 += 32) {
            const int g = g0 + lane;
            unsigned term4 = 0;
            if (g < groups) {
                const uint32_t run = __vcmpltu4(pl4[g], P4);            // 0xff per running slot
                if (run) {
                    const uint2 r = rem4[g];
                    const uint32_t dlo = __vsub2(r.x, now2), dhi = __vsub2(r.y, now2);    // steps until departure, per u16 lane
                    const uint32_t zlo = __vcmpeq2(dlo, 0u), zhi = __vcmpeq2(dhi, 0u);
                    const unsigned runbits = (run & 1u) | ((run >> 7) & 2u) | ((run >> 14) & 4u) | ((run >> 21) & 8u);
                    term4 = ((zlo & 1u) | ((zlo >> 15) & 2u) | ((zhi & 1u) << 2) | ((zhi >> 13) & 8u)) & runbits;
                    const unsigned stay = runbits & ~term4;
                    if (stay & 1u) dmin = min(dmin, dlo & 0xffffu);
                    if (stay & 2u) dmin = min(dmin, dlo >> 16);
                    if (stay & 4u) dmin = min(dmin, dhi & 0xffffu);
                    if (stay & 8u) dmin = min(dmin, dhi >> 16);
                }
            }
            unsigned m = __ballot_sync(FULL, term4 != 0);
            if (m) {                                                      // some VM finished (:248-265)
                __syncwarp();
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    unsigned t4 = __shfl_sync(FULL, term4, b);
                    for (unsigned tt = t4; tt; tt &= tt - 1) {
                        const int vf = 4 * (g0 + b) + __ffs(tt) - 1;
                        if (served == 0) freed0 = vf; else if (served == 1) freed1 = vf; else if (served == 2) freed2 = vf;
                        else if (served == 3) freed3 = vf;
                        served++;
                    }
                    if (lane == 0) {
                        while (t4) {
                            const int j = __ffs(t4) - 1;
                            t4 &= t4 - 1;
                            const int vv = 4 * (g0 + b) + j, pm = (int)place[vv];
                            cpu[pm] -= e.sz64[cpuc[vv] & 0x7f];
                            mem[pm] -= e.sz64[memc[vv]];
                            if (cpu[pm] < 1e-7) cpu[pm] = 0.0;           // :267-268 clamp, applied here for this PM
                            if (mem[pm] < 1e-7) mem[pm] = 0.0;
                            refresh_cap(e, pm);
                            place[vv] = (PT)(P + 1); cpuc[vv] = 0; memc[vv] = 0; rem[vv] = 0;
                            if (vmstat) vmstat_close(vs.slots + vv * 4, tnow - vs.slots[vv * 4], 0u, vs.hist, vs.totals);
                        }
                    }
                }
                __syncwarp();
            }
        }
        }
    } else if (TM) {
        if (team_scan_needed) {
        // team scan -> bitmap of the slots that finished; the main warp retires them in slot order exactly like the
        // byte-placement path above (per-PM subtraction, 1e-7 clamp and capacity-code refresh)
        __syncwarp();
        if (lane == 0) { e.ctl()[1] = (int)tnow; e.ctl()[2] = (int)0xffffffffu; }
        team_run(e, TEAM_COUNTDOWN, nth);
        dmin = (unsigned)e.ctl()[2];
        const unsigned* tmk = e.tmask();
        const int n_chunks = (V + 31) / 32;
        for (int cb0 = 0; cb0 < n_chunks; cb0 += 32) {
            unsigned nz = __ballot_sync(FULL, cb0 + lane < n_chunks && tmk[cb0 + lane] != 0u);
            while (nz) {
                const int c = cb0 + __ffs(nz) - 1;
                nz &= nz - 1;
                unsigned mm = tmk[c];
                served += __popc(mm);
                if (lane == 0) {
                    while (mm) {
                        const int vv = 32 * c + __ffs(mm) - 1, pm = (int)place[vv];
                        mm &= mm - 1;
                        cpu[pm] -= e.sz64[cpuc[vv] & 0x7f];
                        mem[pm] -= e.sz64[memc[vv]];
                        if (cpu[pm] < 1e-7) cpu[pm] = 0.0;           // :267-268 clamp, applied here for this PM
                        if (mem[pm] < 1e-7) mem[pm] = 0.0;
                        refresh_cap(e, pm);
                        place[vv] = (PT)(P + 1); cpuc[vv] = 0; memc[vv] = 0; rem[vv] = 0;
                        e.emask()[vv >> 5] |= 1u << (vv & 31);         // empty from now on (admissions below)
                        if (vmstat) vmstat_close(vs.slots + vv * 4, tnow - vs.slots[vv * 4], 0u, vs.hist, vs.totals);
                    }
                }
                __syncwarp();
            }
        }
        }
    } else {
        if (scan) {
        for (int c0 = 0; c0 < V; c0 += 32) {
            const int v = c0 + lane;
            const int pl = v < V ? (int)place[v] : P + 1;
            const bool running = pl < P;
            const uint32_t d = running ? (((uint32_t)rem[v] - now16) & 0xffffu) : 1u;
            const bool term = running && d == 0u;
            if (running && !term) dmin = min(dmin, d);
            unsigned m = __ballot_sync(FULL, term);
            served += __popc(m);
            if (m) {
                if (lane == 0) {
                    unsigned mm = m;
                    while (mm) {
                        const int b = __ffs(mm) - 1;
                        mm &= mm - 1;
                        const int vv = c0 + b, pm = (int)place[vv];
                        cpu[pm] -= e.sz64[cpuc[vv] & 0x7f];
                        mem[pm] -= e.sz64[memc[vv]];
                        if (vmstat) vmstat_close(vs.slots + vv * 4, tnow - vs.slots[vv * 4], 0u, vs.hist, vs.totals);
                    }
                }
                __syncwarp();
                if (term) { place[v] = (PT)(P + 1); cpuc[v] = 0; memc[v] = 0; rem[v] = 0; }
                need_full_refresh = true;
            }
        }
        }
    }
    if (scan) {
        // the earliest departure among the VMs still running (distance d >= 1 from this step), for the steps to come
        if (!(TM && sizeof(PT) != 1)) dmin = __reduce_min_sync(FULL, dmin);
        __syncwarp();
        if (lane == 0) *e.next_dep() = dmin == 0xffffffffu ? 0xffffffffu : tnow + dmin;
    }
    __syncwarp();
    PROF_ADD(8);                                               // 8: departures
    // ---- 4. clamp (env.py:267-268): values only shrink when something was subtracted this step ----
    if ((served > 0 && !(TM && sizeof(PT) != 1)) || n_susp > 0) {      // (the team path clamps at each departure)
        for (int q = lane; q < P; q += 32) {
            bool ch = need_full_refresh;
            if (cpu[q] < 1e-7 && cpu[q] != 0.0) { cpu[q] = 0.0; ch = true; }
            if (mem[q] < 1e-7 && mem[q] != 0.0) { mem[q] = 0.0; ch = true; }
            if (ch) refresh_cap(e, q);
        }
        __syncwarp();
    }

    // ---- 5b. admissions (_accept_vm_requests, env.py:273-293) ----
    int quota = n_arr;                                         // admissions still allowed this step
    if (trace_mode == VMGYM_TRACE_PRESAMPLED) {
        const long long left = tr.admissions_len - (long long)admission_pos;
        if ((long long)quota > left) { quota = (int)(left > 0 ? left : 0); exhausted = 1; }
    }
    const int n_empty0 = (int)sc->n_empty + served;            // empty slots before admission
    int admitted = 0;
    long long csum = 0, msum = 0;
    if (quota > 0 && n_empty0 > 0) {
        // When no slot was empty before this step, the empty slots are exactly the ones this step's departures freed,
        // already known in slot order: lane k admits into the k-th of them.  Otherwise scan for the lowest-index empties.
        const bool known = sizeof(PT) == 1 && sc->n_empty == 0 && served <= 4;
        auto admit_chunk = [&](int v, bool empty) {
            const unsigned m = __ballot_sync(FULL, empty);
            const int rank = admitted + __popc(m & ((1u << lane) - 1u));
            if (empty && rank < quota) {                       // lowest-index empty slots (:275-277)
                const uint32_t j = admission_pos + (uint32_t)rank;
                uint32_t cc, mc, svc;
                if (trace_mode == VMGYM_TRACE_PRESAMPLED) {
                    const uint32_t w = tr.d_admissions[env_id * tr.admissions_len + j];
                    cc = w & 0xff; mc = (w >> 8) & 0xff; svc = w >> 16;
                } else {
                    const Philox4 r = philox_dev(j, 2u, k0, k1);
                    const uint32_t span = 2u * (uint32_t)(tr.size_hi_code - tr.size_lo_code);
                    cc = (uint32_t)tr.size_lo_code + ((mulhi32(r.x, span) + 1u) >> 1);
                    mc = (uint32_t)tr.size_lo_code + ((mulhi32(r.y, span) + 1u) >> 1);
                    const uint64_t us = ((uint64_t)r.z << 32) | r.w;
                    const int ks = e.svc_bracket ? cdf_search_bracketed(e.svc_cdf, tr.service_cdf_len, e.svc_bracket, us)
                                                 : cdf_search(e.svc_cdf, tr.service_cdf_len, us);
                    svc = (uint32_t)(tr.service_kmin + ks) + 1u;                                  // Poisson + 1 (:289)
                }
                place[v] = (PT)P;
                cpuc[v] = (uint8_t)cc; memc[v] = (uint8_t)mc; rem[v] = (uint16_t)svc;
                csum += cc; msum += mc;
                if (vmstat) *reinterpret_cast<uint4*>(vs.slots + v * 4) = make_uint4(tnow, 0u, 0u, 0u);   // arrival (env.py:293)
            }
            admitted += __popc(m);
        };
        if (TM && sizeof(PT) != 1) {
            // team mode: the countdown phase left a bitmap of the empty slots (+ this step's departures): only the
            // chunks that hold one are visited, lowest index first
            const unsigned* emk = e.emask();
            const int n_chunks = (V + 31) / 32;
            for (int cb0 = 0; cb0 < n_chunks && admitted < quota; cb0 += 32) {
                unsigned nz = __ballot_sync(FULL, cb0 + lane < n_chunks && emk[cb0 + lane] != 0u);
                while (nz && admitted < quota) {
                    const int c = cb0 + __ffs(nz) - 1;
                    nz &= nz - 1;
                    admit_chunk(32 * c + lane, ((emk[c] >> lane) & 1u) != 0u);
                }
            }
        } else {
            for (int c0 = 0; c0 < (known ? 32 : V) && admitted < quota; c0 += 32) {
                int v = c0 + lane;
                bool empty = v < V && !known && (int)place[v] == P + 1;
                if (known) {
                    v = lane == 0 ? freed0 : (lane == 1 ? freed1 : (lane == 2 ? freed2 : freed3));
                    empty = lane < served;
                }
                admit_chunk(v, empty);
            }
        }
        admitted = min(admitted, quota);
        csum = (long long)__reduce_add_sync(FULL, (unsigned)csum);
        msum = (long long)__reduce_add_sync(FULL, (unsigned)msum);
        __syncwarp();
    }

    PROF_ADD(9);                                               // 9: clamp + admissions
    // ---- 6. metrics (env.py:112-121) from the incrementally maintained slot counters ----
    const int n_empty = n_empty0 - admitted;
    const int waiting = (int)sc->n_waiting - n_place + n_susp + admitted;
    const int arrived = V - n_empty;

    // ---- 7. reward (env.py:123-156): a function of the state alone, so a step that changed nothing repeats the last one ----
    double reward = 0.0;
    if (n_place + n_susp + served + admitted == 0 && tnow > 1u) {
        reward = sc->last_reward;
    } else if (arrived > 0) {
        if (reward_fn == VMGYM_REWARD_WR) {
            reward = -((double)waiting / (double)arrived);
        } else if (reward_fn == VMGYM_REWARD_UT) {
            reward = reward_ut(cpu, mem, P, p.beta);
        } else {
            reward = reward_kl_env(e, arrived, p.cap_target);
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
        sc->arrival_pos = arrival_pos + 1;
        sc->admission_pos = admission_pos + (uint32_t)admitted;
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
    PROF_ADD(10);                                              // 10: reward + counters
    StepResult res;
    res.reward = reward; res.terminated = terminated; res.rejected = rejected; res.waiting = waiting; res.arrived = arrived;
    res.changed = n_place + n_susp + served + admitted;      // anything that can change which waiting VMs fit
    return res;
}

template <typename PT>
__device__ __noinline__ void write_obs_generic(const Env<PT> e, float* __restrict__ o)
{
    const int P = e.P, V = e.V;
    const PT* place = e.place();
    const uint8_t* cpuc = e.cpuc();
    const uint8_t* memc = e.memc();
    const double* cpu = e.cpu();
    const double* mem = e.mem();
#pragma unroll 1
    for (int v = e.lane; v < V; v += 32) { o[v] = (float)place[v]; o[V + v] = e.sz32[cpuc[v] & 0x7f]; o[2 * V + v] = e.sz32[memc[v]]; }
#pragma unroll 1
    for (int q = e.lane; q < P; q += 32) { o[3 * V + q] = (float)cpu[q]; o[3 * V + P + q] = (float)mem[q]; }
}

// The observation row for a caller with HOST-resident observations: `o` (device) is the reference copy of the row, `host` its
// device-mapped mirror in pinned host memory; an entry is stored to both only when its bit pattern changed.
template <typename PT>
__device__ __noinline__ void write_obs_mirrored(const Env<PT> e, float* o, float* host)
{
    const int P = e.P, V = e.V;
    const PT* place = e.place();
    const uint8_t* cpuc = e.cpuc();
    const uint8_t* memc = e.memc();
    const double* cpu = e.cpu();
    const double* mem = e.mem();
#pragma unroll 1
    for (int v = e.lane; v < V; v += 32) {
        const float a = (float)place[v], c = e.sz32[cpuc[v] & 0x7f], m = e.sz32[memc[v]];
        if (__float_as_uint(o[v]) != __float_as_uint(a)) { o[v] = a; host[v] = a; }
        if (__float_as_uint(o[V + v]) != __float_as_uint(c)) { o[V + v] = c; host[V + v] = c; }
        if (__float_as_uint(o[2 * V + v]) != __float_as_uint(m)) { o[2 * V + v] = m; host[2 * V + v] = m; }
    }
#pragma unroll 1
    for (int q = e.lane; q < P; q += 32) {
        const float c = (float)cpu[q], m = (float)mem[q];
        if (__float_as_uint(o[3 * V + q]) != __float_as_uint(c)) { o[3 * V + q] = c; host[3 * V + q] = c; }
        if (__float_as_uint(o[3 * V + P + q]) != __float_as_uint(m)) { o[3 * V + P + q] = m; host[3 * V + P + q] = m; }
    }
}

// observation row (env.py:295-296): f32[ placement | vm_cpu | vm_memory | cpu | memory ]
template <typename PT>
__device__ __forceinline__ void write_obs(const Env<PT>& e, float* __restrict__ o)
{
    const int P = e.P, V = e.V;
    const PT* place = e.place();
    const uint8_t* cpuc = e.cpuc();
    const uint8_t* memc = e.memc();
    const double* cpu = e.cpu();
    const double* mem = e.mem();
    if (sizeof(PT) == 1 && (V & 3) == 0 && (P & 3) == 0) {
        // 128-bit stores: the row and its five segments are 16-byte aligned when V and P are multiples of 4
        float4* o4 = reinterpret_cast<float4*>(o);
        const uint32_t* pl4 = reinterpret_cast<const uint32_t*>(place);
        const uint32_t* cc4 = reinterpret_cast<const uint32_t*>(cpuc);
        const uint32_t* mc4 = reinterpret_cast<const uint32_t*>(memc);
        const int vg = V >> 2, pg = P >> 2;
        for (int g = e.lane; g < vg; g += 32) {
            const uint32_t a = pl4[g], c = cc4[g] & 0x7f7f7f7fu, m = mc4[g];
            o4[g] = make_float4((float)(a & 0xff), (float)((a >> 8) & 0xff), (float)((a >> 16) & 0xff), (float)(a >> 24));
            o4[vg + g] = make_float4(e.sz32[c & 0xff], e.sz32[(c >> 8) & 0xff], e.sz32[(c >> 16) & 0xff], e.sz32[c >> 24]);
            o4[2 * vg + g] = make_float4(e.sz32[m & 0xff], e.sz32[(m >> 8) & 0xff], e.sz32[(m >> 16) & 0xff], e.sz32[m >> 24]);
        }
        const double2* c2 = reinterpret_cast<const double2*>(cpu);
        const double2* m2 = reinterpret_cast<const double2*>(mem);
        for (int g = e.lane; g < pg; g += 32) {
            const double2 a = c2[2 * g], b = c2[2 * g + 1], c = m2[2 * g], d = m2[2 * g + 1];
            o4[3 * vg + g] = make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
            o4[3 * vg + pg + g] = make_float4((float)c.x, (float)c.y, (float)d.x, (float)d.y);
        }
        return;
    }
    write_obs_generic(e, o);
}

__device__ __forceinline__ void fill_tables(double* sz64, float* sz32)
{
    for (int k = threadIdx.x; k < SIZE_TABLE; k += blockDim.x) {
        const double x = code_to_f64(k);         // == k / 100.0 == np.around(u, 2) for the code k (env.py:212-219)
        sz64[k] = x;
        sz32[k] = (float)x;                        // env.py:296 float32 cast
    }
}

// plain 128-bit copy of a record (the non-bulk fallback of the staging path)
static __device__ __noinline__ void copy16(void* dst, const void* src, int bytes, int lane)
{
    __syncwarp();
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll 1
    for (int i = lane; i < bytes / 16; i += 32) d4[i] = s4[i];
    __syncwarp();
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
// Grid-stride over envs, one warp per env, <= 4 warps per CTA.
// ---------------------------------------------------------------------------------------------------
// PC / VC: compile-time pms / vms of the instantiation (0 = take them from the layout at run time); the named
// configs of the reference (config/10.yml, config/100.yml) get fully unrolled loops.
// SPEC >= 0 additionally fixes (agent, tiebreak, reward, trace mode) = spec_* fields at compile time (SPEC < 0: run time).
// `post_agent` != 0: the kernel also computes that agent's act() on the new state (vmgym_outputs.d_next_action, stable ties)
constexpr int make_spec(int agent, int tiebreak, int reward, int mode, int post_agent = 0)
{
    return agent | (tiebreak << 4) | (reward << 8) | (mode << 12) | (post_agent << 16);
}

// TM: team mode (one env per CTA, warp 0 + helper warps; see "Team mode" above) — instantiated for u16 placements.
// DB: double-buffered records for launches in which a warp steps several envs one after the other (grid capped at the
// resident CTAs): the next env's record is fetched (bulk-async, second buffer + second mbarrier) while the current one is
// stepped, and the write-back of the previous one drains in the background.  Needs bulk loads and stores (use_bulk bits 0, 1).
// ROT_CT: rotation launch known at compile time (1 yes, 0 no, -1 = look at rot_batches): the specialised kernels drop the other
// loop nest's index arithmetic from the per-record scaffolding.
template <typename PT, int PC, int VC, int SPEC, bool TM = false, bool DB = false, int ROT_CT = -1>
__global__ void __launch_bounds__(TM ? 256 : (PC == 10 ? 256 : 896), TM ? 3 : (PC == 10 ? 4 : 1)) step_kernel(const __grid_constant__ StepParams p)
{
    constexpr int REWARD_CT = SPEC >= 0 ? ((SPEC >> 8) & 0xf) : 0;
    constexpr int MODE_CT = SPEC >= 0 ? ((SPEC >> 12) & 0xf) : -1;
    const int agent_k = SPEC >= 0 ? (SPEC & 0xf) : p.agent;
    const int tiebreak_k = SPEC >= 0 ? ((SPEC >> 4) & 0xf) : p.tiebreak;
    extern __shared__ __align__(128) unsigned char smem[];
    const DevLayout& L = p.L;
    const int cP = PC ? PC : L.P, cV = VC ? VC : L.V, cD = 3 * cV + 2 * cP;
    double* sz64 = reinterpret_cast<double*>(smem);
    float* sz32 = reinterpret_cast<float*>(smem + SIZE_TABLE * 8);
    uint32_t* arr_cdf_s = reinterpret_cast<uint32_t*>(smem + SIZE_TABLE * 12);
    uint16_t* svc_bracket_s = reinterpret_cast<uint16_t*>(reinterpret_cast<uint64_t*>(smem + SIZE_TABLE * 12) + ARR_CDF_SMEM);
    // team mode: every warp of the CTA works on the same env (region 0); otherwise a warp per env
    const int lane = threadIdx.x & 31, warp = TM ? 0 : (int)(threadIdx.x >> 5), wpc = TM ? 1 : (int)(blockDim.x >> 5);
    const bool helper = TM && threadIdx.x >= 32;
    const int nth = TM ? (int)blockDim.x : 32;
    const int wstride = DB ? align_up(L.sm_stride + L.rec_bytes, 128) : L.sm_stride;     // DB: + the second record buffer
    unsigned char* base = smem + L.sm_tables + (size_t)warp * wstride;
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + L.sm_bar);                        // DB: bar[1] belongs to the second buffer
    const bool BULK = DB || (p.use_bulk & 1) != 0, BULK_ST = DB || (p.use_bulk & 2) != 0;
    const long long stride = (long long)gridDim.x * wpc;
    const long long env0 = (long long)blockIdx.x * wpc + warp;
    // the warp's work list.  Plain launches walk env0, env0 + stride, ...; a rotation launch takes the rot_steps batch steps of
    // each env index it owns (batch rot_first, rot_first + 1, ... modulo rot_batches) before moving on to the next index
    const bool ROT = ROT_CT >= 0 ? (ROT_CT != 0) : (p.rot_batches > 0);
    const long long n_idx = ROT ? p.rot_envs : p.n_envs;
    const int per_idx = ROT ? p.rot_steps : 1;
    // Team-mode rotation ("balanced", BAL): a CTA steps ONE env at a time and a batch rarely holds a multiple of the resident
    // CTAs (1024 envs on 592 CTA slots at 1000 PMs: two rounds, the second 73 % full).  So the CTAs do not own env indexes but
    // RECORDS: record r = batch * rot_envs + idx belongs to CTA r mod gridDim.x — still one fixed CTA per record, hence every
    // record's steps stay in one CTA's program order — and the loop nest is batch step outside, the CTA's records of that batch
    // inside.  The remainder of each batch then lands on different CTAs from batch to batch and the load evens out.
    const bool BAL = TM && ROT;                       // compile-time false outside team mode
    const long long bal_g = (long long)gridDim.x;     // (grid <= rot_envs: every CTA owns >= 1 record of every batch)
    auto bal_idx0 = [&](int rb) -> int {              // lowest env index of batch rb owned by this CTA
        return (int)(((long long)blockIdx.x - ((long long)rb * p.rot_envs) % bal_g + bal_g) % bal_g);
    };
    const bool any_item = BAL ? per_idx > 0 : (env0 < n_idx && per_idx > 0);
    const long long first_env = BAL ? (long long)p.rot_first * p.rot_envs + bal_idx0(p.rot_first)
                                    : env0 + (ROT ? (long long)p.rot_first * p.rot_envs : 0ll);
    // Programmatic dependent launch (use_bulk bit 2): let the NEXT kernel of the stream start launching right away (its CTAs
    // take the slots this grid frees as its fast CTAs finish) ...
    const bool PDL = (p.use_bulk & 4) != 0;
    if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // Start staging this warp's first record before anything else: the bulk copy only needs the warp's own mbarrier, so
    // its DRAM round trip overlaps the table set-up below instead of following it.  (With PDL the records may still be
    // written by the previous grid: the load waits for it, after the tables.)
    if (BULK && lane == 0 && !helper) {
        mbar_init(bar, 1);
        if (DB) mbar_init(bar + 1, 1);
        fence_barrier_init();
        if (!PDL && any_item) {
            mbar_arrive_expect_tx(bar, (uint32_t)L.rec_bytes);
            bulk_g2s(base, p.state + first_env * (long long)L.rec_bytes, (uint32_t)L.rec_bytes, bar);
        }
    }
    const bool philox = (MODE_CT >= 0 ? MODE_CT : p.tr.mode) == VMGYM_TRACE_PHILOX;
    // the specialised Philox kernels are only launched with a small arrival table and service brackets (launch_step checks)
    constexpr bool TABLES_CT = SPEC >= 0 && MODE_CT == VMGYM_TRACE_PHILOX;
    const bool arr_in_smem = TABLES_CT || (philox && p.tr.arrival_cdf_len <= ARR_CDF_SMEM);
    const bool have_bracket = TABLES_CT || (philox && p.tr.d_service_bracket != nullptr);
    {
        // CTA-wide tables: every global load is issued before the first shared-memory store so that the round trips
        // overlap each other (and the record's bulk copy) instead of queueing behind one another
        const int t = threadIdx.x, nt = blockDim.x;
        const uint64_t a0 = (arr_in_smem && t < p.tr.arrival_cdf_len) ? p.tr.d_arrival_cdf[t] : 0ull;
        const uint16_t b0 = (have_bracket && t < SVC_BRACKETS + 1) ? p.tr.d_service_bracket[t] : (uint16_t)0;
        fill_tables(sz64, sz32);
        if (arr_in_smem) {
            if (t < p.tr.arrival_cdf_len) arr_cdf_s[t] = (uint32_t)(a0 >> 32);
            for (int k = t + nt; k < p.tr.arrival_cdf_len; k += nt) arr_cdf_s[k] = (uint32_t)(p.tr.d_arrival_cdf[k] >> 32);
        }
        if (have_bracket) {
            if (t < SVC_BRACKETS + 1) svc_bracket_s[t] = b0;
            for (int k = t + nt; k < SVC_BRACKETS + 1; k += nt) svc_bracket_s[k] = p.tr.d_service_bracket[k];
        }
    }
    __syncthreads();
    if (PDL) {
        // ... and wait here, tables built, for the previous grid to complete and flush before touching any record / output
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (BULK && lane == 0 && !helper && any_item) {
            mbar_arrive_expect_tx(bar, (uint32_t)L.rec_bytes);
            bulk_g2s(base, p.state + first_env * (long long)L.rec_bytes, (uint32_t)L.rec_bytes, bar);
        }
    }

    Env<PT> e;
    e.base = base; e.rec = base; e.L = &p.L; e.sz64 = sz64; e.sz32 = sz32; e.P = cP; e.V = cV; e.lane = lane; e.tune = p.use_bulk;
    e.arr_cdf = p.tr.d_arrival_cdf;
    e.arr_cdf32 = arr_in_smem ? arr_cdf_s : nullptr;
    e.svc_cdf = p.tr.d_service_cdf;                  // searched from a 64-way bracket, on admissions only
    e.svc_bracket = have_bracket ? svc_bracket_s : nullptr;
    uint32_t phase = 0;
    if (helper) {
        // helper warps: wait for each record of this CTA, then serve the main warp's phases until it closes the env
        for (long long o = BAL ? 0ll : env0; o < (BAL ? (long long)per_idx : n_idx); o += BAL ? 1ll : stride) {
            const int i0 = BAL ? bal_idx0((int)((p.rot_first + o) % p.rot_batches)) : 0;
            for (int i = i0; i < (BAL ? (int)n_idx : per_idx); i += BAL ? (int)stride : 1) {
                if (BULK) { mbar_wait(bar, phase); phase ^= 1; }
                team_serve(e, (int)threadIdx.x, nth);
            }
        }
        return;
    }
    uint32_t phase1 = 0;
    int cur = 0;                             // DB: which record buffer holds the current env
    // loop nest: env index outside, its rotation steps inside — or, balanced team rotation, batch step outside, records inside
    for (long long o = BAL ? 0ll : env0; o < (BAL ? (long long)per_idx : n_idx); o += BAL ? 1ll : stride) {
    const int rb_o = BAL ? (int)((p.rot_first + o) % p.rot_batches) : 0;
    const int i0 = BAL ? bal_idx0(rb_o) : 0;
    for (int i = i0, rbi = ROT ? p.rot_first : 0; i < (BAL ? (int)n_idx : per_idx); i += BAL ? (int)stride : 1, rbi = (rbi + 1 == p.rot_batches) ? 0 : rbi + 1) {
        const long long idx = BAL ? (long long)i : o;
        const int rk = BAL ? (int)o : i;
        const int rb = BAL ? rb_o : rbi;
        const long long env = ROT ? (long long)rb * p.rot_envs + idx : idx;
        const bool first_item = BAL ? (o == 0 && i == i0) : (idx == env0 && rk == 0);
        // 32 x 32 -> 64-bit products for the per-record addresses (launches hold < 2^31 envs, fill_params): one IMAD.WIDE each
        // instead of a 64 x 64 multiply chain per output pointer
        const unsigned env_u = (unsigned)env;
        unsigned char* grec = p.state + (unsigned long long)env_u * (unsigned)L.rec_bytes;
#ifdef VMGYM_PROF
        const long long _pload = clock64();
#endif
        // ---- stage the record into shared memory ----
        // rotation: this warp's next record (the same env index in the next batch) is 3.4 KB that nobody has touched for B - 1
        // batch steps, i.e. a DRAM round trip on the warp's serial chain.  Ask L2 for it now; the load one step later then hits
        // L2.  (The bytes still cross HBM once per step; a second shared-memory buffer per warp — DB — would cost resident
        // warps at this record size.)  Measured: 6.54 -> 6.46 us per 4096-env step.
        if (!DB && ROT && !BAL && lane == 0 && rk + 1 < per_idx && (p.use_bulk & 64) == 0) {
            const long long nxt = (long long)((rb + 1 == p.rot_batches) ? 0 : rb + 1) * p.rot_envs + idx;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.state + (unsigned long long)(unsigned)nxt * (unsigned)L.rec_bytes), "r"((uint32_t)L.rec_bytes) : "memory");
        }
        if (BAL && lane == 0 && (p.use_bulk & 64) == 0) {
            // balanced team rotation: the CTA's next record = its next one in this batch, else its first one in the next batch
            long long nxt = -1;
            if (idx + stride < n_idx) nxt = (long long)rb * p.rot_envs + idx + stride;
            else if (rk + 1 < per_idx) {
                const int nb = (rb + 1 == p.rot_batches) ? 0 : rb + 1;
                nxt = (long long)nb * p.rot_envs + bal_idx0(nb);
            }
            if (nxt >= 0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.state + (unsigned long long)(unsigned)nxt * (unsigned)L.rec_bytes), "r"((uint32_t)L.rec_bytes) : "memory");
        }
        if (DB) {
            unsigned char* other = cur ? base : base + L.sm_stride;
            e.rec = cur ? base + L.sm_stride : base;
            // the next item of this warp: the same env index in the next batch of the rotation, else the next env index
            const bool more_rot = rk + 1 < per_idx;
            const long long nxt = more_rot ? (long long)((rb + 1 == p.rot_batches) ? 0 : rb + 1) * p.rot_envs + idx
                                           : idx + stride + (ROT ? (long long)p.rot_first * p.rot_envs : 0ll);
            if (lane == 0 && (more_rot || idx + stride < n_idx)) {
                bulk_wait_read0();            // the other buffer's write-back (previous env) has left shared memory
                mbar_arrive_expect_tx(bar + (cur ^ 1), (uint32_t)L.rec_bytes);
                bulk_g2s(other, p.state + (unsigned long long)(unsigned)nxt * (unsigned)L.rec_bytes, (uint32_t)L.rec_bytes, bar + (cur ^ 1));
            }
            if (cur) { mbar_wait(bar + 1, phase1); phase1 ^= 1; }
            else { mbar_wait(bar, phase); phase ^= 1; }
        } else if (BULK) {
            if (lane == 0 && !first_item) {
                mbar_arrive_expect_tx(bar, (uint32_t)L.rec_bytes);
                bulk_g2s(base, grec, (uint32_t)L.rec_bytes, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
        } else {
            copy16(base, grec, L.rec_bytes, lane);
        }

#ifdef VMGYM_PROF
        if (threadIdx.x == 0) atomicAdd(&g_prof[0], (unsigned long long)(clock64() - _pload));      // 0: record load wait
#endif
        PROF_T0();
        uint8_t* valid_g = p.out.d_valid ? p.out.d_valid + (unsigned long long)env_u * (unsigned)cV : nullptr;
        StepResult res;
        res.reward = 0.0; res.terminated = 0; res.rejected = 0; res.waiting = 0; res.arrived = 0; res.changed = 0;
        double* st_acc = reinterpret_cast<double*>(base + L.sm_stats);     // per-launch stats sums (lane 0)
        if (p.out.d_stats && lane < VMGYM_STATS) st_acc[lane] = 0.0;
        // STATUS_QUIET: "a fused agent's act() followed by the env's apply loop would change nothing".  It is
        // established by a full evaluation after which the step changed no placement (the agent proposed nothing, or
        // every proposal was rejected by the fp64 capacity check — SURVEY App. B-2) and nothing departed or was
        // admitted: the next act() then sees the same float32 PM loads and the same waiting VMs, proposes the same
        // actions and the env rejects them again.  While it holds, act() and the apply loop are skipped.  The key
        // records which agent established it (0 = "no waiting VM fits anywhere", which holds for every agent).
        const bool need_vectors = p.out.d_action != nullptr || p.out.d_valid != nullptr;
        const uint32_t my_key = (uint32_t)(agent_k | (tiebreak_k << 4));
        const uint32_t status0 = e.sc()->status;
        bool obs_stale = (status0 & STATUS_OBS_STALE) != 0;
        bool quiet = (status0 & STATUS_QUIET) != 0;
        uint32_t quiet_key = (status0 & STATUS_KEY_MASK) >> STATUS_KEY_SHIFT;
        int quiet_rejected = (int)(status0 >> 16);
        for (int s = 0; s < p.n_steps; s++) {
            bool have_actions, evaluated = false;
            int n_found = 0;
            if (agent_k != VMGYM_AGENT_NONE) {
                if (!(quiet && (quiet_key == 0 || quiet_key == my_key)) || need_vectors) {
                    // the agent sees the float32 observation of the current state (env.py:296)
                    if (!TM) {                   // (team mode: part of the team's PREP phase inside agent_act)
                        const double* cpu = e.cpu();
                        const double* mem = e.mem();
                        for (int q = lane; q < cP; q += 32) { e.cpu32()[q] = (float)cpu[q]; e.mem32()[q] = (float)mem[q]; }
                        __syncwarp();
                    }
                    AgentView<PT> av;
                    av.place = e.place(); av.cc = e.cpuc(); av.mc = e.memc(); av.c32 = nullptr; av.m32 = nullptr; av.sz32 = sz32;
                    PROF_ADD(1);                                                                         // 1: setup
                    n_found = agent_act<PT, TM>(e, av, agent_k, tiebreak_k, true, nth);
                    PROF_ADD(2);                                                                         // 2: agent act
                    evaluated = true;
                    if (p.out.d_action) {        // the action vector: proposals, else the current placement (firstfit.py:29)
                        PT* ao = reinterpret_cast<PT*>(p.out.d_action) + (unsigned long long)env_u * (unsigned)cV;
                        for (int v = lane; v < cV; v += 32)
                            ao[v] = ((e.prop()[v >> 5] >> (v & 31)) & 1u) ? (PT)e.act()[v] : e.place()[v];
                    }
                }
                have_actions = n_found > 0;
            } else {
                // external actions: stage the row and mark the slots whose action differs from their placement
                const int adt = p.action_dtype;
                const unsigned char* arow = reinterpret_cast<const unsigned char*>(p.action) + (unsigned long long)env_u * (unsigned)(cV * dtype_bytes(adt));
                unsigned any = 0;
                for (int c0 = 0; c0 < cV; c0 += 32) {
                    const int v = c0 + lane;
                    const int a = v < cV ? load_action(arow, adt, v) : 0;
                    const bool diff = v < cV && a != (int)e.place()[v];
                    if (diff) e.act()[v] = (uint16_t)a;
                    const unsigned m = __ballot_sync(FULL, diff);
                    if (lane == 0) e.prop()[c0 >> 5] = m;
                    any |= m;
                }
                __syncwarp();
                have_actions = any != 0;
            }
            res = env_step<PT, REWARD_CT, MODE_CT, (SPEC < 0), TM>(e, p, env, valid_g, have_actions, nth);
            PROF_ADD(3);                                                                                 // 3: env step
            if (res.changed) {
                quiet = false;
                obs_stale = true;
            } else if (evaluated) {
                quiet = true;
                quiet_key = n_found == 0 ? 0u : my_key;
                quiet_rejected = res.rejected;
            } else if (quiet && agent_k != VMGYM_AGENT_NONE) {
                res.rejected = quiet_rejected;        // the skipped proposals would have been rejected again
            }
            if (p.out.d_stats) stats_update(e, res, st_acc, p.cap_target);
            if (res.terminated) break;
        }

        // ---- the agent's act() on the state just produced (vmgym_outputs.d_next_action; generic kernels and the specs with a post agent) ----
        constexpr int POST_CT = SPEC >= 0 ? ((SPEC >> 16) & 0xf) : -1;
        if constexpr (POST_CT != 0) {
            if (p.out.d_next_action) {
                const int post_agent = SPEC >= 0 ? POST_CT : p.out.next_agent;
                const int post_tie = SPEC >= 0 ? VMGYM_TIE_STABLE : p.out.next_tiebreak;
                if (!TM) {
                    const double* cpu = e.cpu();
                    const double* mem = e.mem();
                    for (int q = lane; q < cP; q += 32) { e.cpu32()[q] = (float)cpu[q]; e.mem32()[q] = (float)mem[q]; }
                    __syncwarp();
                }
                AgentView<PT> av;
                av.place = e.place(); av.cc = e.cpuc(); av.mc = e.memc(); av.c32 = nullptr; av.m32 = nullptr; av.sz32 = sz32;
                agent_act<PT, TM>(e, av, post_agent, post_tie, true, nth);
                PT* ao = reinterpret_cast<PT*>(p.out.d_next_action) + (unsigned long long)env_u * (unsigned)cV;
                for (int v = lane; v < cV; v += 32)
                    ao[v] = ((e.prop()[v >> 5] >> (v & 31)) & 1u) ? (PT)e.act()[v] : e.place()[v];
                __syncwarp();
            }
        }

        // ---- outputs ----
        if (p.out.d_obs) {
            // a persistent observation buffer keeps the rows of envs whose state did not change (a quiet step changes
            // nothing: no 4(3V+2P)-byte store); the mirrored variant (host-resident observations) is persistent by definition
            // and is compiled where external actions are possible (generic kernels and the agent-NONE spec)
            constexpr bool MIRROR_OK = SPEC < 0 || (SPEC & 0xf) == VMGYM_AGENT_NONE;
            const bool mirrored = MIRROR_OK && p.out.d_obs_mirror != nullptr;
            const bool persistent = p.out.obs_persistent != 0 || mirrored;
            if (!persistent || obs_stale) {
                float* orow = p.out.d_obs + (unsigned long long)env_u * (unsigned)cD;
                if (mirrored) write_obs_mirrored(e, orow, p.out.d_obs_mirror + (unsigned long long)env_u * (unsigned)cD);
                else if (TM) {
                    if (lane == 0) {
                        e.ctl()[2] = (int)(unsigned)((unsigned long long)orow & 0xffffffffull);
                        e.ctl()[3] = (int)(unsigned)((unsigned long long)orow >> 32);
                    }
                    team_run(e, TEAM_OBS, nth);
                } else write_obs(e, orow);
            }
            if (persistent) obs_stale = false;
        }
        if (lane == 0) {
            e.sc()->status = (e.sc()->status & STATUS_EXHAUSTED) | (obs_stale ? STATUS_OBS_STALE : 0u) |
                             (quiet ? (STATUS_QUIET | (quiet_key << STATUS_KEY_SHIFT) | ((uint32_t)quiet_rejected << 16)) : 0u);
            if (p.out.d_reward) p.out.d_reward[env_u] = res.reward;
            if (p.out.d_terminated) p.out.d_terminated[env_u] = (uint8_t)res.terminated;
            if (p.out.d_stats) {
                double* st = p.out.d_stats + (unsigned long long)env_u * (unsigned)VMGYM_STATS;
                for (int k = 0; k < VMGYM_STATS; k++) st[k] += st_acc[k];
            }
        }

        PROF_ADD(4);                                                                                     // 4: outputs (obs row)
        if (TM) team_run(e, TEAM_END, nth);       // helpers move on to the next record of this CTA
        // ---- write the record back ----
        if (DB) {
            fence_proxy_async();          // generic-proxy writes to smem -> visible to the async proxy
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(grec, e.rec, (uint32_t)L.rec_bytes);
                bulk_commit();            // drained before this buffer is loaded again (wait_group.read in front of the prefetch)
            }
            __syncwarp();
            cur ^= 1;
        } else if (BULK_ST) {
            fence_proxy_async();          // generic-proxy writes to smem -> visible to the async proxy
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(grec, base, (uint32_t)L.rec_bytes);
                bulk_commit();
                bulk_wait_read0();        // smem may be overwritten by the next bulk load after this
            }
            __syncwarp();
        } else {
            copy16(grec, base, L.rec_bytes, lane);
            if (BULK) fence_proxy_async();        // generic-proxy reads of smem before the next record's async-proxy write
        }
        PROF_ADD(5);                                                                                     // 5: write-back
    }
    }
    if (DB && lane == 0) bulk_wait_read0();       // shared memory must outlive the last write-backs
}


// Record's per-VM lists as of now: running histograms + the VMs that still occupy a slot (one warp per env).
template <typename PT>
__global__ void vmstats_finalize_kernel(DevLayout L, const unsigned char* state, long long n_envs, const uint32_t* slots,
                                        const uint32_t* hist, const unsigned long long* totals, uint32_t* hist_out,
                                        unsigned long long* totals_out)
{
    const int lane = threadIdx.x & 31;
    const long long env = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (env >= n_envs) return;
    const unsigned char* rec = state + env * (long long)L.rec_bytes;
    const PT* place = reinterpret_cast<const PT*>(rec + L.off_place);
    const vmgym_env_scalars* sc = reinterpret_cast<const vmgym_env_scalars*>(rec + L.off_scal);
    uint32_t* ho = hist_out + env * 2ll * VMGYM_VMSTAT_BINS;
    const uint32_t* hi = hist + env * 2ll * VMGYM_VMSTAT_BINS;
    for (int i = lane; i < 2 * VMGYM_VMSTAT_BINS; i += 32) ho[i] = hi[i];
    unsigned long long* to = totals_out + env * 4;
    if (lane < 4) to[lane] = totals[env * 4 + lane];
    __syncwarp();
    if (lane == 0) {
        const uint32_t next = (uint32_t)sc->timestep;            // last executed step + 1: samples run through step next - 1
        const uint32_t* s = slots + env * (long long)L.V * 4;
        for (int v = 0; v < L.V; v++) {
            if ((int)place[v] <= L.P)
                vmstat_close(s + v * 4, next - s[v * 4], s[v * 4 + 3] ? next - s[v * 4 + 3] : 0u, ho, to);
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
    uint16_t* rcap = reinterpret_cast<uint16_t*>(rec + L.off_cap);
    for (int q = lane; q < L.P; q += 32) rcap[q] = (uint16_t)(100u | (100u << 8));   // empty PM: every size code fits
    if (lane == 0) {
        sc->timestep = 1;                                                 // env.py:197
        sc->n_waiting = 0;
        sc->n_empty = (uint16_t)L.V;
        sc->seed = seeds ? seeds[env] : keep.seed;
        sc->arrival_pos = rewind ? 0u : keep.arrival_pos;
        sc->admission_pos = rewind ? 0u : keep.admission_pos;
        sc->status = (rewind ? 0u : (keep.status & STATUS_EXHAUSTED)) | (obs ? 0u : STATUS_OBS_STALE);
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
// per-warp shared memory: obs row f32[D] | place PT[Vp] | cc u8[Vp] | mc u8[Vp] | act u16[Vp] | tmp | fitm | cap | prop
struct ActLayout { int row, place, cc, mc, act, tmp, fit, prop, stride; };
__host__ __device__ inline ActLayout act_layout(const DevLayout& L)
{
    ActLayout a;
    a.row = 0;
    a.place = align_up(4 * L.D, 16);
    a.cc = a.place + 2 * L.Vp;
    a.mc = a.cc + L.Vp;
    a.act = a.mc + L.Vp;
    a.tmp = a.act + 2 * L.Vp;
    a.fit = a.tmp + align_up(6 * L.Pp, 16);
    a.prop = a.fit + 512 + align_up(2 * L.Pp, 16);
    a.stride = align_up(a.prop + 4 * ((L.Vp + 31) / 32), 128);
    return a;
}

template <typename PT>
__global__ void act_kernel(DevLayout Lg, int agent, int tiebreak, const float* obs, long long n_envs, void* action, int adt)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ DevLayout L;                       // per-CTA layout whose scratch offsets point into the act layout
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    float* sz32 = reinterpret_cast<float*>(smem);
    for (int k = threadIdx.x; k < SIZE_TABLE; k += blockDim.x) sz32[k] = (float)code_to_f64(k);
    const ActLayout al = act_layout(Lg);
    if (threadIdx.x == 0) {
        L = Lg;
        L.sm_cpu32 = al.row + 4 * 3 * Lg.V;       // the row's cpu / memory segments are the agent's local loads
        L.sm_mem32 = al.row + 4 * (3 * Lg.V + Lg.P);
        L.sm_act = al.act; L.sm_tmp = al.tmp; L.sm_fit = al.fit; L.sm_prop = al.prop;
    }
    __syncthreads();
    const long long env = (long long)blockIdx.x * wpc + warp;
    if (env >= n_envs) return;
    unsigned char* base = smem + SIZE_TABLE * 4 + (size_t)warp * al.stride;
    float* row = reinterpret_cast<float*>(base + al.row);
    PT* place = reinterpret_cast<PT*>(base + al.place);
    uint8_t* cc = base + al.cc;
    uint8_t* mc = base + al.mc;
    const int V = Lg.V, P = Lg.P;
    const float* o = obs + env * (long long)Lg.D;
    for (int i = lane; i < Lg.D; i += 32) row[i] = o[i];
    __syncwarp();
    // slot arrays as the agent reads them (utils.py:41 astype(int)); a size that is not the float32 image of a
    // hundredth gets code 0 so the fit filter never rejects it
    for (int v = lane; v < Lg.Vp; v += 32) {
        int pl = P + 1, kc = 0, km = 0;
        if (v < V) {
            pl = (int)row[v];
            const float xc = row[V + v], xm = row[2 * V + v];
            const int a = __float2int_rn(xc * 100.0f), b = __float2int_rn(xm * 100.0f);
            const bool exact = a >= 0 && a <= 100 && b >= 0 && b <= 100 && sz32[a] == xc && sz32[b] == xm;
            kc = exact ? a : 0; km = exact ? b : 0;
            pl = pl < 0 ? P + 1 : min(pl, (sizeof(PT) == 1) ? 255 : 65535);
        }
        place[v] = (PT)pl; cc[v] = (uint8_t)kc; mc[v] = (uint8_t)km;
    }
    __syncwarp();
    Env<PT> e;
    e.base = base; e.rec = base; e.L = &L; e.sz64 = nullptr; e.sz32 = sz32; e.arr_cdf = nullptr; e.arr_cdf32 = nullptr; e.svc_cdf = nullptr; e.svc_bracket = nullptr;
    e.P = P; e.V = V; e.lane = lane;
    AgentView<PT> av;
    av.place = place; av.cc = cc; av.mc = mc; av.c32 = row + V; av.m32 = row + 2 * V; av.sz32 = sz32;
    agent_act(e, av, agent, tiebreak, false);
    unsigned char* ao = reinterpret_cast<unsigned char*>(action) + env * (long long)V * dtype_bytes(adt);
    for (int v = lane; v < V; v += 32) {
        const int a = ((e.prop()[v >> 5] >> (v & 31)) & 1u) ? (int)e.act()[v] : (int)row[v];
        if (adt == VMGYM_U8) ao[v] = (uint8_t)a;
        else if (adt == VMGYM_I16) reinterpret_cast<int16_t*>(ao)[v] = (int16_t)a;
        else reinterpret_cast<long long*>(ao)[v] = a;
    }
}

}  // namespace vmgym
