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
#include <stdlib.h>
#include <string.h>

#include "vmgym_env_kernels.cuh"

namespace vmgym {

// ---------------------------------------------------------------------------------------------------
// host-side layout
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int g_warps_per_cta = 0;
static int g_use_bulk = 7;      // bit 0: bulk-async record loads, bit 1: bulk-async record stores, bit 2: programmatic dependent launch

static int fail(int code, const char* fmt, const char* detail = "")
{
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}

// per-env shared-memory scratch behind the record.  `tmp` holds the numpy-introsort keys / permutation (6 P bytes, inside
// agent_act) or the `kl` reward's compacted size codes (2 V bytes, inside env_step).  small_tmp (team-mode launches that do not
// sort): 64 bytes for the team's scan results, and the `kl` compaction moves onto the agent's float32 view, which is dead outside
// agent_act — at 1000 PMs that is the 6 KB between three and four envs per SM.
static void layout_scratch(DevLayout& l, int pb, bool small_tmp)
{
    l.sm_cpu32 = l.rec_bytes;
    l.sm_mem32 = l.sm_cpu32 + align_up(4 * l.P, 16);
    l.sm_act = l.sm_mem32 + align_up(4 * l.P, 16);
    l.sm_tmp = l.sm_act + 2 * l.Vp;
    const int tmp_bytes = small_tmp ? 64 : align_up((2 * l.Vp > 6 * l.Pp) ? 2 * l.Vp : 6 * l.Pp, 16);
    l.sm_kl = small_tmp ? l.sm_cpu32 : l.sm_tmp;
    l.sm_fit = l.sm_tmp + tmp_bytes;                       // u32 fitm[128] | u16 cap[Pp]
    l.sm_prop = l.sm_fit + 512 + align_up(2 * l.Pp, 16);      // u32 prop[ceil(Vp/32)]: slots whose action differs
    l.sm_stats = l.sm_prop + align_up(4 * ((l.Vp + 31) / 32), 16);   // f64[VMGYM_STATS] eval-summary sums of the launch
    l.sm_bar = l.sm_stats + 8 * VMGYM_STATS;
    // team mode (u16 placements = large shapes, one env per CTA): command words + three slot-chunk bitmaps (departures, candidates, empty slots)
    l.sm_team = l.sm_bar + 16;
    const int team_bytes = pb == 2 ? 16 + 3 * align_up(4 * ((l.Vp + 31) / 32), 16) : 0;
    l.sm_stride = align_up(l.sm_team + team_bytes, 128);
}

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
    l.off_cap = l.off_memc + l.Vp;
    l.off_scal = l.off_cap + align_up(2 * l.Pp, 16);
    // + 16 B: parked Philox words (arrival draws), + 16 B: next-departure clock (u32) and reserve
    l.rec_bytes = align_up(l.off_scal + (int)sizeof(vmgym_env_scalars) + 16 + 16, 128);
    layout_scratch(l, pb, false);
    l.svc_cdf_smem = 0;                                    // service table stays in global memory (used on admissions only)
    l.sm_tables = align_up(SIZE_TABLE * 8 + SIZE_TABLE * 4 + ARR_CDF_SMEM * 8 + l.svc_cdf_smem * 8 + (SVC_BRACKETS + 1) * 2, 128);
    if (L) *L = l;
    if (pub) {
        pub->record_bytes = l.rec_bytes; pub->pms_padded = l.Pp; pub->vms_padded = l.Vp; pub->place_bytes = pb;
        pub->off_cpu = 0; pub->off_memory = l.off_mem; pub->off_remaining = l.off_rem; pub->off_placement = l.off_place;
        pub->off_cpu_code = l.off_cpuc; pub->off_mem_code = l.off_memc; pub->off_scalars = l.off_scal; pub->off_capacity = l.off_cap;
        pub->obs_dim = l.D; pub->action_dim = l.A; pub->smem_bytes_per_env = l.sm_stride;
    }
    return VMGYM_OK;
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

// SM count of the CURRENT device (cached per device: one process may drive several)
static int sm_count()
{
    static int counts[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& c = counts[dev & 63];
    if (c == 0) {
        cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
        if (c <= 0) c = 148;
    }
    return c;
}

// warps per CTA: keep >= ~6 CTAs of work per SM when the batch is small (balance across the 148 SMs), up to 8
// warps when it is large.
static int pick_warps(long long n_envs, int smem_per_warp, int smem_fixed)
{
    if (g_warps_per_cta > 0) {
        int w = g_warps_per_cta > 28 ? 28 : g_warps_per_cta;       // step_kernel is compiled for <= 896 threads per CTA
        while (w > 1 && smem_fixed + w * smem_per_warp > 200 * 1024) w--;
        return w;
    }
    int w = 4;
    // one wave that fills the machine (28 resident warps per SM at 72 registers): 4 CTAs of 7 warps per SM instead of 7 of 4 —
    // fewer CTAs to dispatch and fewer table prologues (measured 14.1 -> 13.6 us per 4096-env step; 8 or 16 warps, which do
    // not divide 28, are slower)
    if (n_envs > (long long)sm_count() * 21 && n_envs <= (long long)sm_count() * 28 && smem_fixed + 7 * smem_per_warp <= 56 * 1024) return 7;
    while (w > 1 && n_envs < (long long)sm_count() * w * 6) w >>= 1;
    while (w > 1 && smem_fixed + w * smem_per_warp > 200 * 1024) w >>= 1;
    return w;
}

// warp-per-env kernel by record buffering (db) and launch kind (rot); the specialised fused-agent kernels have the launch kind
// compiled in (external-action kernels are never rotation launches; generic kernels test rot_batches)
template <typename PT, int PC, int VC, int SPEC>
static void (*pick_db(bool db, bool rot))(const StepParams)
{
    if constexpr (SPEC >= 0 && (SPEC & 0xf) != VMGYM_AGENT_NONE) {
        if (rot) {
            if (db) return step_kernel<PT, PC, VC, SPEC, false, true, 1>;
            return step_kernel<PT, PC, VC, SPEC, false, false, 1>;
        }
        if (db) return step_kernel<PT, PC, VC, SPEC, false, true, 0>;
        return step_kernel<PT, PC, VC, SPEC, false, false, 0>;
    } else if constexpr (SPEC >= 0) {
        if (db) return step_kernel<PT, PC, VC, SPEC, false, true, 0>;
        return step_kernel<PT, PC, VC, SPEC, false, false, 0>;
    } else {
        if (db) return step_kernel<PT, PC, VC, SPEC, false, true>;
        return step_kernel<PT, PC, VC, SPEC, false, false>;
    }
}

// specialised team-mode kernel by launch kind
template <typename PT, int PC, int VC, int SPEC>
static void (*pick_team(bool rot))(const StepParams)
{
    if (rot) return step_kernel<PT, PC, VC, SPEC, true, false, 1>;
    return step_kernel<PT, PC, VC, SPEC, true, false, 0>;
}

template <typename PT>
static int launch_step(StepParams& sp, cudaStream_t st)
{
    sp.use_bulk = g_use_bulk;
    // u16 placements = large shapes: team mode, one env per CTA with `w` warps (warp 0 steps, the others join the bulk phases)
    static const bool team_ok = getenv("VMGYM_NO_TEAM") == nullptr;                        // A/B switch for experiments
    const bool team = team_ok && sizeof(PT) == 2;
    if (team) {
        // launches that never sort (stable ties / first-fit / external actions) do without the sort scratch
        const bool sorts = (sp.agent == VMGYM_AGENT_BESTFIT && sp.tiebreak == VMGYM_TIE_NUMPY_INTROSORT) ||
                           (sp.out.d_next_action && sp.out.next_agent == VMGYM_AGENT_BESTFIT && sp.out.next_tiebreak == VMGYM_TIE_NUMPY_INTROSORT);
        const bool kl_fits = 2 * align_up(sp.L.V, 16) <= sp.L.sm_act - sp.L.sm_cpu32;
        static const bool small_ok = getenv("VMGYM_TEAM_FULL_SCRATCH") == nullptr;          // A/B switch for experiments
        if (small_ok && !sorts && kl_fits) layout_scratch(sp.L, 2, true);
    }
    const DevLayout& L = sp.L;
    // env indexes the warps share out: a rotation launch schedules one batch's worth, each warp then walks all the batches
    const long long n_sched = sp.rot_batches > 0 ? sp.rot_envs : sp.n_envs;
    int w;
    if (team) {
        const long long one = (long long)L.sm_tables + L.sm_stride + 1024;                  // + the per-CTA reservation
        long long per_sm = (228 * 1024) / one;
        const long long need = (n_sched + sm_count() - 1) / sm_count();
        if (per_sm > need) per_sm = need;
        if (per_sm < 1) per_sm = 1;
        // (the specialised team kernels below are compiled under the same launch bounds: at most the generic kernel's registers)
        static int team_regs = 0;
        if (team_regs == 0) {
            cudaFuncAttributes fa;
            team_regs = cudaFuncGetAttributes(&fa, step_kernel<PT, 0, 0, -1, (sizeof(PT) == 2)>) == cudaSuccess && fa.numRegs > 0 ? fa.numRegs : 80;
        }
        const long long by_regs = 65536 / (per_sm * 32 * team_regs);                        // warps per CTA the register file allows
        w = g_warps_per_cta > 0 ? g_warps_per_cta : (int)(48 / per_sm < by_regs ? 48 / per_sm : by_regs);
        w = w < 1 ? 1 : (w > 8 ? 8 : w);                                                    // compiled for <= 256 threads
    } else {
        w = pick_warps(n_sched, L.sm_stride, L.sm_tables);
        if (L.P == 10 && L.V == 30 && w > 8) w = 8;        // the 10-PM kernels are compiled for <= 256 threads: 64 registers, 32 resident warps per SM (72: 28 warps; 1.59 -> 1.68 G env-steps/s; 56 registers: 1.60)
    }
    // double-buffered records once a warp steps several envs per launch (more envs than ~1.4 x the resident warps) — for
    // small records only: measured +7.5 % at the 10-PM shape (2^20 envs), but -6 % at 100 PMs, where the second 3.4 KB
    // buffer per warp costs more resident warps than the prefetch wins (0.79 -> 0.73 of the roofline at 32768 envs).
    // Needs bulk loads + stores; use_bulk bit 4 switches it off, bit 5 forces it for any record size (A/B runs, tests).
    const int wstride_db = align_up(L.sm_stride + L.rec_bytes, 128);
    const bool db = !team && (sp.use_bulk & 3) == 3 && (sp.use_bulk & 16) == 0 && (L.rec_bytes <= 1024 || (sp.use_bulk & 32) != 0) &&
                    n_sched * (sp.rot_batches > 0 ? sp.rot_steps : 1) * 10 > (long long)sm_count() * 7 * w * 14 && (size_t)L.sm_tables + (size_t)w * wstride_db <= 227 * 1024;
    const size_t smem = (size_t)L.sm_tables + (size_t)(team ? 1 : w) * (db ? wstride_db : L.sm_stride);
    if (smem > 227 * 1024) return fail(VMGYM_EUNSUPPORTED, "env record does not fit in shared memory (pms/vms too large)");
    const bool rot = sp.rot_batches > 0;
    void (*kern)(const StepParams) = pick_db<PT, 0, 0, -1>(db, rot);
    if (team) {
        kern = step_kernel<PT, 0, 0, -1, (sizeof(PT) == 2)>;
        // the benchmark configurations of the large shapes: fused heuristic agent, reward wr, stable ties, Philox arrivals, no per-VM
        // statistics — a third of the generic kernel's code (agent / reward / trace mode compiled in)
        static const bool specialise_team_ok = getenv("VMGYM_NO_SPECIALIZE") == nullptr;
        // (the specialised kernels take the small arrival table and the service brackets for granted)
        const bool specialise_team = specialise_team_ok && sp.tr.arrival_cdf_len <= ARR_CDF_SMEM && sp.tr.d_service_bracket != nullptr;
        if constexpr (sizeof(PT) == 2) {
            if (specialise_team && sp.reward_fn == VMGYM_REWARD_WR && sp.tiebreak == VMGYM_TIE_STABLE && !sp.out.d_vm_slots && !sp.out.d_next_action &&
                sp.tr.mode == VMGYM_TRACE_PHILOX) {
                if (sp.agent == VMGYM_AGENT_BESTFIT && L.P == 1000 && L.V == 3000)         // BASELINE config 5's shape
                    kern = pick_team<PT, 1000, 3000, make_spec(VMGYM_AGENT_BESTFIT, 0, VMGYM_REWARD_WR, VMGYM_TRACE_PHILOX)>(rot);
                else if (sp.agent == VMGYM_AGENT_BESTFIT)
                    kern = pick_team<PT, 0, 0, make_spec(VMGYM_AGENT_BESTFIT, 0, VMGYM_REWARD_WR, VMGYM_TRACE_PHILOX)>(rot);
                else if (sp.agent == VMGYM_AGENT_FIRSTFIT)
                    kern = pick_team<PT, 0, 0, make_spec(VMGYM_AGENT_FIRSTFIT, 0, VMGYM_REWARD_WR, VMGYM_TRACE_PHILOX)>(rot);
            }
        }
    }
    // (the specialised kernels are all Philox kernels and take the small arrival table and the service brackets for granted: philox_fast)
    static const bool specialise = getenv("VMGYM_NO_SPECIALIZE") == nullptr;               // A/B switch for experiments
    const bool philox_fast = sp.tr.mode == VMGYM_TRACE_PHILOX && sp.tr.arrival_cdf_len <= ARR_CDF_SMEM && sp.tr.d_service_bracket != nullptr;
    if constexpr (sizeof(PT) == 1) {
    if (specialise && L.P == 100 && L.V == 300) {                                           // config/100.yml
        kern = pick_db<PT, 100, 300, -1>(db, rot);
        // the benchmark configurations of this shape: fused heuristic agents / external actions, reward wr, stable ties
        if (sp.reward_fn == VMGYM_REWARD_WR && sp.agent == VMGYM_AGENT_NONE && philox_fast && !sp.out.d_vm_slots &&
            sp.out.d_next_action && sp.out.next_agent == VMGYM_AGENT_BESTFIT && sp.out.next_tiebreak == VMGYM_TIE_STABLE) {
            // HostVecEnv's step: external actions, then best-fit's act on the new state
            kern = pick_db<PT, 100, 300, make_spec(VMGYM_AGENT_NONE, 0, VMGYM_REWARD_WR, VMGYM_TRACE_PHILOX, VMGYM_AGENT_BESTFIT)>(db, rot);
        } else if (sp.reward_fn == VMGYM_REWARD_WR && sp.tiebreak == VMGYM_TIE_STABLE && !sp.out.d_vm_slots && !sp.out.d_next_action) {   // per-VM stats, other next-action outputs: generic kernel
            if (sp.agent == VMGYM_AGENT_BESTFIT && philox_fast)
                kern = pick_db<PT, 100, 300, make_spec(VMGYM_AGENT_BESTFIT, 0, VMGYM_REWARD_WR, VMGYM_TRACE_PHILOX)>(db, rot);
            else if (sp.agent == VMGYM_AGENT_FIRSTFIT && philox_fast)
                kern = pick_db<PT, 100, 300, make_spec(VMGYM_AGENT_FIRSTFIT, 0, VMGYM_REWARD_WR, VMGYM_TRACE_PHILOX)>(db, rot);
            else if (sp.agent == VMGYM_AGENT_NONE && philox_fast)
                kern = pick_db<PT, 100, 300, make_spec(VMGYM_AGENT_NONE, 0, VMGYM_REWARD_WR, VMGYM_TRACE_PHILOX)>(db, rot);
        }
    } else if (specialise && L.P == 10 && L.V == 30) {
        kern = pick_db<PT, 10, 30, -1>(db, rot);                                                // config/10.yml
        if (sp.reward_fn == VMGYM_REWARD_WR && sp.tiebreak == VMGYM_TIE_STABLE && !sp.out.d_vm_slots && !sp.out.d_next_action && philox_fast) {
            if (sp.agent == VMGYM_AGENT_FIRSTFIT)                                          // BASELINE configs[0]: first-fit evaluation
                kern = pick_db<PT, 10, 30, make_spec(VMGYM_AGENT_FIRSTFIT, 0, VMGYM_REWARD_WR, VMGYM_TRACE_PHILOX)>(db, rot);
            else if (sp.agent == VMGYM_AGENT_BESTFIT)
                kern = pick_db<PT, 10, 30, make_spec(VMGYM_AGENT_BESTFIT, 0, VMGYM_REWARD_WR, VMGYM_TRACE_PHILOX)>(db, rot);
        }
    }
    }
    // per (kernel, smem, warps) launch plan, computed once (also keeps these calls out of CUDA-graph capture)
    static thread_local size_t plan_smem = 0;
    static thread_local int plan_w = 0, plan_occ = 1, plan_dev = -1;
    static thread_local void (*plan_kern)(const StepParams) = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    if (plan_smem != smem || plan_w != w || plan_kern != kern || plan_dev != dev) {
        int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute");
        if (rc) return rc;
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, w * 32, smem);
        plan_occ = occ < 1 ? 1 : occ;
        plan_smem = smem; plan_w = w; plan_kern = kern; plan_dev = dev;
    }
    const int occ = plan_occ;
    long long blocks = team ? n_sched : (n_sched + w - 1) / w;
    const long long cap = (long long)sm_count() * occ;
    if (blocks > cap) blocks = cap;
    if (sp.use_bulk & 4) {
        // programmatic stream serialization: this launch may begin before the previous kernel of the stream has finished; the
        // kernel itself waits (griddepcontrol.wait) before it touches records or outputs
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)blocks); lc.blockDim = dim3((unsigned)(w * 32)); lc.dynamicSmemBytes = smem; lc.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = attr; lc.numAttrs = 1;
        return check_cuda(cudaLaunchKernelEx(&lc, kern, sp), "step_kernel launch (PDL)");
    }
    kern<<<(unsigned)blocks, w * 32, smem, st>>>(sp);
    return check_cuda(cudaGetLastError(), "step_kernel launch");
}

static int fill_params(StepParams* sp, const vmgym_config* cfg, void* d_state, int64_t n_envs, const vmgym_trace* trace,
                       const vmgym_outputs* out)
{
    int rc = make_layout(cfg, &sp->L, nullptr);
    if (rc) return rc;
    if (n_envs == 0) return VMGYM_OK;     // callers test n_envs == 0 again before launching
    if (!d_state || n_envs < 0) return fail(VMGYM_EINVAL, "null state / negative n_envs");
    if (n_envs > 0x7fffffffLL) return fail(VMGYM_EINVAL, "more than 2^31 - 1 envs in one launch");   // the step kernel's record addresses are 32 x 32-bit products
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
    const int n_vm = (sp->out.d_vm_slots != nullptr) + (sp->out.d_vm_hist != nullptr) + (sp->out.d_vm_totals != nullptr);
    if (n_vm != 0 && n_vm != 3) return fail(VMGYM_EINVAL, "d_vm_slots / d_vm_hist / d_vm_totals must be all set or all NULL");
    sp->action = nullptr; sp->action_dtype = VMGYM_U8; sp->use_bulk = 1; sp->agent = VMGYM_AGENT_NONE; sp->tiebreak = 0; sp->n_steps = 1;
    sp->rot_batches = 0; sp->rot_steps = 0; sp->rot_first = 0; sp->rot_envs = 0;
    return VMGYM_OK;
}

extern "C" {

const char* vmgym_last_error(void) { return g_err; }
void vmgym_internal_set_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }
int vmgym_abi_version(void) { return VMGYM_ABI_VERSION; }

int vmgym_set_tuning(int warps_per_cta, int use_bulk_copy)
{
    if (warps_per_cta < 0 || warps_per_cta > 28) return fail(VMGYM_EINVAL, "warps_per_cta must be 0..28");
    g_warps_per_cta = warps_per_cta;
    g_use_bulk = use_bulk_copy & 127;     // bit 6: no L2 prefetch of the rotation's next record (A/B runs)
    return VMGYM_OK;
}

#ifdef VMGYM_PROF
/* development builds only: cycles per section of the step kernel's decision warp (see PROF_ADD), reset after reading */
int vmgym_debug_prof(unsigned long long* out16)
{
    unsigned long long zero[16] = {0};
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, vmgym::g_prof, sizeof(zero));
    cudaMemcpyToSymbol(vmgym::g_prof, zero, sizeof(zero));
    return VMGYM_OK;
}
#endif

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
    if (n_envs == 0) return VMGYM_OK;
    if (!d_state || n_envs < 0) return fail(VMGYM_EINVAL, "null state / negative n_envs");
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
    if (n_envs == 0) return VMGYM_OK;
    if (!d_action) return fail(VMGYM_EINVAL, "null action");
    if (sp.out.d_next_action) {
        if (sp.out.next_agent != VMGYM_AGENT_FIRSTFIT && sp.out.next_agent != VMGYM_AGENT_BESTFIT)
            return fail(VMGYM_EUNSUPPORTED, "d_next_action: next_agent must be firstfit or bestfit");
        if (sp.out.next_tiebreak != VMGYM_TIE_STABLE && sp.out.next_tiebreak != VMGYM_TIE_NUMPY_INTROSORT)
            return fail(VMGYM_EINVAL, "d_next_action: unknown tiebreak");
    }
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
    sp.out.d_next_action = nullptr;            // (vmgym_step only)
    sp.agent = agent; sp.tiebreak = tiebreak; sp.n_steps = n_steps;
    cudaStream_t st = (cudaStream_t)stream;
    return sp.L.P <= 253 ? launch_step<uint8_t>(sp, st) : launch_step<uint16_t>(sp, st);
}

int vmgym_agent_step_rotation(const vmgym_config* cfg, void* d_state, int64_t envs_per_batch, int32_t n_batches, int32_t first_batch,
                              int32_t n_batch_steps, const vmgym_trace* trace, int agent, int tiebreak, int n_steps,
                              const vmgym_outputs* out, void* stream)
{
    if (n_batches < 1 || n_batch_steps < 0 || first_batch < 0 || first_batch >= n_batches || envs_per_batch < 0)
        return fail(VMGYM_EINVAL, "rotation: need n_batches >= 1, 0 <= first_batch < n_batches, n_batch_steps >= 0");
    StepParams sp;
    int rc = fill_params(&sp, cfg, d_state, envs_per_batch * n_batches, trace, out);
    if (rc) return rc;
    if (agent != VMGYM_AGENT_FIRSTFIT && agent != VMGYM_AGENT_BESTFIT)
        return fail(VMGYM_EUNSUPPORTED, "fused agent must be firstfit or bestfit");
    if (tiebreak != VMGYM_TIE_STABLE && tiebreak != VMGYM_TIE_NUMPY_INTROSORT) return fail(VMGYM_EINVAL, "unknown tiebreak");
    if (n_steps < 1) return fail(VMGYM_EINVAL, "n_steps must be >= 1");
    if (envs_per_batch == 0 || n_batch_steps == 0) return VMGYM_OK;
    sp.out.d_next_action = nullptr;            // (vmgym_step only)
    sp.agent = agent; sp.tiebreak = tiebreak; sp.n_steps = n_steps;
    sp.rot_batches = n_batches; sp.rot_steps = n_batch_steps; sp.rot_first = first_batch; sp.rot_envs = envs_per_batch;
    cudaStream_t st = (cudaStream_t)stream;
    return sp.L.P <= 253 ? launch_step<uint8_t>(sp, st) : launch_step<uint16_t>(sp, st);
}

int vmgym_agent_act(const vmgym_config* cfg, int agent, int tiebreak, const float* d_obs, int64_t n_envs, void* d_action,
                    int action_dtype, void* stream)
{
    DevLayout L;
    int rc = make_layout(cfg, &L, nullptr);
    if (rc) return rc;
    if (n_envs == 0) return VMGYM_OK;
    if (!d_obs || !d_action || n_envs < 0) return fail(VMGYM_EINVAL, "null obs/action");
    if (agent != VMGYM_AGENT_FIRSTFIT && agent != VMGYM_AGENT_BESTFIT) return fail(VMGYM_EUNSUPPORTED, "agent must be firstfit or bestfit");
    if (n_envs == 0) return VMGYM_OK;
    const int per_warp = act_layout(L).stride;
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
    if (n_envs == 0) return VMGYM_OK;
    if (!d_state || !d_obs || n_envs < 0) return fail(VMGYM_EINVAL, "null state/obs");
    if (n_envs == 0) return VMGYM_OK;
    const int threads = 256;
    const long long blocks = (n_envs * 32 + threads - 1) / threads;
    if (L.P <= 253) observe_kernel<uint8_t><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(L, (const unsigned char*)d_state, n_envs, d_obs);
    else observe_kernel<uint16_t><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(L, (const unsigned char*)d_state, n_envs, d_obs);
    return check_cuda(cudaGetLastError(), "observe_kernel launch");
}

// Host mirror of the observation buffer (HostVecEnv): entries whose bits differ from the shadow copy are stored to the mapped host
// buffer and to the shadow; reward / done rows follow.  Element-wise, grid-stride, 128-bit loads.
__global__ void obs_mirror_kernel(const float4* __restrict__ obs, float4* __restrict__ prev, float* __restrict__ host, long long n4,
                                  const double* __restrict__ d_reward, double* __restrict__ h_reward,
                                  const uint8_t* __restrict__ d_term, uint8_t* __restrict__ h_term, long long n_envs)
{
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = t0; i < n4; i += stride) {
        const float4 a = obs[i], b = prev[i];
        const bool cx = __float_as_uint(a.x) != __float_as_uint(b.x), cy = __float_as_uint(a.y) != __float_as_uint(b.y);
        const bool cz = __float_as_uint(a.z) != __float_as_uint(b.z), cw = __float_as_uint(a.w) != __float_as_uint(b.w);
        if (cx | cy | cz | cw) {
            prev[i] = a;
            float* h = host + 4 * i;
            if (cx) h[0] = a.x;
            if (cy) h[1] = a.y;
            if (cz) h[2] = a.z;
            if (cw) h[3] = a.w;
        }
    }
    for (long long e = t0; e < n_envs; e += stride) {
        if (h_reward) h_reward[e] = d_reward[e];
        if (h_term) h_term[e] = d_term[e];
    }
}

int vmgym_obs_mirror_update(const float* d_obs, float* d_shadow, float* h_obs, int64_t n_floats, const double* d_reward, double* h_reward,
                            const uint8_t* d_terminated, uint8_t* h_terminated, int64_t n_envs, void* stream)
{
    if (!d_obs || !d_shadow || !h_obs || n_floats < 0 || n_envs < 0 || (h_reward && !d_reward) || (h_terminated && !d_terminated))
        return fail(VMGYM_EINVAL, "vmgym_obs_mirror_update: null operand");
    if ((n_floats & 3) || ((uintptr_t)d_obs & 15) || ((uintptr_t)d_shadow & 15) || ((uintptr_t)h_obs & 15))
        return fail(VMGYM_EINVAL, "vmgym_obs_mirror_update: buffers must be 16-byte aligned and a multiple of 4 floats long");
    if (n_floats == 0 && n_envs == 0) return VMGYM_OK;
    const int threads = 256;
    long long blocks = (n_floats / 4 + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    obs_mirror_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(d_obs), reinterpret_cast<float4*>(d_shadow),
                                                                                  h_obs, n_floats / 4, d_reward, h_reward, d_terminated, h_terminated, n_envs);
    return check_cuda(cudaGetLastError(), "obs_mirror_kernel launch");
}

int vmgym_invalid_action_mask(const vmgym_config* cfg, const void* d_state, int64_t n_envs, uint8_t* d_mask, void* stream)
{
    DevLayout L;
    int rc = make_layout(cfg, &L, nullptr);
    if (rc) return rc;
    if (n_envs == 0) return VMGYM_OK;
    if (!d_state || !d_mask || n_envs < 0) return fail(VMGYM_EINVAL, "null state/mask");
    if (n_envs == 0) return VMGYM_OK;
    const int threads = 256;
    const long long blocks = (n_envs * 32 + threads - 1) / threads;
    if (L.P <= 253) mask_kernel<uint8_t><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(L, (const unsigned char*)d_state, n_envs, d_mask);
    else mask_kernel<uint16_t><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(L, (const unsigned char*)d_state, n_envs, d_mask);
    return check_cuda(cudaGetLastError(), "mask_kernel launch");
}

int vmgym_vmstats_finalize(const vmgym_config* cfg, const void* d_state, int64_t n_envs, const uint32_t* d_vm_slots,
                           const uint32_t* d_vm_hist, const uint64_t* d_vm_totals, uint32_t* d_hist_out,
                           uint64_t* d_totals_out, void* stream)
{
    DevLayout L;
    int rc = make_layout(cfg, &L, nullptr);
    if (rc) return rc;
    if (n_envs == 0) return VMGYM_OK;
    if (!d_state || !d_vm_slots || !d_vm_hist || !d_vm_totals || !d_hist_out || !d_totals_out || n_envs < 0)
        return fail(VMGYM_EINVAL, "null per-VM statistics buffer");
    const int threads = 128;
    const long long blocks = (n_envs * 32 + threads - 1) / threads;
    if (L.P <= 253)
        vmstats_finalize_kernel<uint8_t><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
            L, (const unsigned char*)d_state, n_envs, d_vm_slots, d_vm_hist, (const unsigned long long*)d_vm_totals, d_hist_out,
            (unsigned long long*)d_totals_out);
    else
        vmstats_finalize_kernel<uint16_t><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
            L, (const unsigned char*)d_state, n_envs, d_vm_slots, d_vm_hist, (const unsigned long long*)d_vm_totals, d_hist_out,
            (unsigned long long*)d_totals_out);
    return check_cuda(cudaGetLastError(), "vmstats_finalize_kernel launch");
}

}  // extern "C"
