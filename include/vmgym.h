/*
 * vmgym.h — C ABI of the B200-native batched VM placement/migration env (libvmgym.so).
 *
 * This is the drop-in boundary of the hot path: plain pointers and sizes, no torch types.  All pointers
 * named `d_*` are DEVICE pointers; every call enqueues its kernels on the caller's `stream` (a
 * cudaStream_t passed as void*), never synchronises, and is CUDA-graph capturable.  Return value: 0 on
 * success, a negative VMGYM_E* code otherwise (nothing is thrown across the boundary;
 * vmgym_last_error() gives the message of the calling thread's last failure).
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference root,
 * yzh503/vm-placement-migration-gym).  The reference has no FFI of its own (it is pure Python), so the
 * "binding a maintainer would add" is the ctypes stub shown in INTEGRATION.md.
 */
#ifndef VMGYM_H_
#define VMGYM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMGYM_ABI_VERSION 9

enum vmgym_status {
    VMGYM_OK = 0,
    VMGYM_EINVAL = -1,      /* bad shape / null pointer / unsupported value */
    VMGYM_ECUDA = -2,       /* a CUDA runtime call failed */
    VMGYM_EARCH = -3,       /* device is not sm_100 */
    VMGYM_EUNSUPPORTED = -4
};

/* vmenv/envs/env.py:125,151,153 and main.py:94 — "reward 1/2/3" of the paper */
enum vmgym_reward { VMGYM_REWARD_WR = 1, VMGYM_REWARD_UT = 2, VMGYM_REWARD_KL = 3 };
/* src/agents/{firstfit,bestfit}.py; src/agents/drlvmp.py:549-617 (worst-fit / min-dot / min-L2 choices) */
enum vmgym_agent {
    VMGYM_AGENT_NONE = 0, VMGYM_AGENT_FIRSTFIT = 1, VMGYM_AGENT_BESTFIT = 2, VMGYM_AGENT_WORSTFIT = 3,
    VMGYM_AGENT_MINDOT = 4, VMGYM_AGENT_MINL2 = 5
};
/* best-fit tie rule (src/agents/bestfit.py:33 uses an unstable np.argsort; see DESIGN.md "tie-break") */
enum vmgym_tiebreak { VMGYM_TIE_STABLE = 0, VMGYM_TIE_NUMPY_INTROSORT = 1 };
/* element type of an action tensor */
enum vmgym_dtype { VMGYM_U8 = 1, VMGYM_I16 = 2, VMGYM_I64 = 3 };
enum vmgym_trace_mode { VMGYM_TRACE_PRESAMPLED = 0, VMGYM_TRACE_PHILOX = 1 };

/* vmenv/envs/config.py:3-16 (the fields the device path needs) + the active step limit. */
typedef struct vmgym_config {
    int32_t pms;                /* P: physical machines */
    int32_t vms;                /* V: VM slots */
    int32_t allow_null_action;  /* action_dim A = P+2 if set else P+1 (env.py:26) */
    int32_t reward_function;    /* enum vmgym_reward */
    int32_t cap_target_util;    /* env.py:117-121 */
    int32_t step_limit;         /* eval_steps in eval mode else training_steps (env.py:160-163) */
    double beta;                /* `ut` mixing weight (env.py:152) */
} vmgym_config;

/* Byte layout of one env record in HBM (struct-of-arrays inside a record, records contiguous).
 * Filled by vmgym_get_layout; the host mirror builds strided tensor views from it. */
typedef struct vmgym_layout {
    int32_t record_bytes;       /* multiple of 128 */
    int32_t pms_padded, vms_padded;
    int32_t place_bytes;        /* 1 (P <= 253) or 2 */
    int32_t off_cpu;            /* f64[P]  PM cpu utilisation accumulators (env.py:190) */
    int32_t off_memory;         /* f64[P]  PM memory utilisation accumulators (env.py:191) */
    int32_t off_remaining;      /* u16[V]  vm_remaining_runtime (env.py:192).  WAITING slots hold the steps left; RUNNING slots hold the
                                   step in which the VM departs, modulo 2^16 (placed in step t with r steps left: t + r - 1), so that
                                   no counter is decremented per step.  With T = the record's timestep (the next step to run),
                                   steps left of a running slot = ((value - T) & 0xffff) + 1.  A u32 "earliest departure step"
                                   follows the scalars at off_scalars + sizeof(vmgym_env_scalars) + 16. */
    int32_t off_placement;      /* u8|u16[V] vm_placement: 0..P-1 running, P waiting, P+1 empty (env.py:187) */
    int32_t off_cpu_code;       /* u8[V]   vm_cpu in hundredths, bit 7 = vm_suspended (env.py:188,204) */
    int32_t off_mem_code;       /* u8[V]   vm_memory in hundredths (env.py:189) */
    int32_t off_scalars;        /* struct vmgym_env_scalars */
    int32_t obs_dim;            /* 3V + 2P (env.py:27,295-296) */
    int32_t action_dim;         /* A */
    int32_t smem_bytes_per_env; /* shared memory one resident env needs in the step kernels */
    int32_t off_capacity;       /* u16[P]  derived cache: per-PM max admissible size codes in the float32 view
                                   (cpu | mem << 8); maintained by every kernel that changes PM utilisation */
} vmgym_layout;

/* Per-env scalar block inside the record (env.py:193-208). */
typedef struct vmgym_env_scalars {
    int32_t timestep;           /* starts at 1 */
    int32_t total_requests, served_requests, dropped_requests, suspend_actions, place_actions;
    uint32_t arrival_pos;       /* arrivals consumed from the arrival stream (rng3, env.py:272) */
    uint32_t admission_pos;     /* entries consumed from the size/service streams (rng1,2,4; env.py:279-289) */
    uint32_t status;            /* bit 0: pre-sampled trace exhausted (the reference would raise at env.py:282);
                                   bit 1: QUIET — a fused agent's act()+apply would change nothing (skipped while set);
                                   bits 8-15: agent/tiebreak that established QUIET (0 = any); bits 16-31: rejected
                                   proposals per quiet step */
    uint16_t n_waiting;         /* slots with vm_placement == P   (env.py:114) */
    uint16_t n_empty;           /* slots with vm_placement == P+1 (env.py:273) */
    uint64_t seed;              /* Philox key of this env */
    int64_t cpu_code_sum;       /* 100 * total_cpu_requested (exact integer) */
    int64_t mem_code_sum;       /* 100 * total_memory_requested */
    double episode_return;      /* sum of rewards since reset */
    double last_reward;
} vmgym_env_scalars;

/* Source of the env's randomness.  PRESAMPLED: arrays drawn by the host exactly like the reference draws
 * them (numpy PCG64, env.py:172-178,211-219,272,289) so trajectories are bit-identical to the reference.
 * PHILOX: counter-based generation inside the kernel (key = env seed). */
typedef struct vmgym_trace {
    int32_t mode;                    /* enum vmgym_trace_mode */
    int32_t reserved;
    const uint16_t* d_arrivals;      /* [n_envs, arrivals_len]  Poisson(arrival_rate) per step */
    int64_t arrivals_len;
    const uint32_t* d_admissions;    /* [n_envs, admissions_len] cpu_code | mem_code<<8 | (Poisson(service)+1)<<16 */
    int64_t admissions_len;
    /* Philox mode: inverse-CDF tables as 64-bit thresholds, P(X <= kmin+i) * 2^64 */
    const uint64_t* d_arrival_cdf;  int32_t arrival_cdf_len;  int32_t arrival_kmin;
    const uint64_t* d_service_cdf;  int32_t service_cdf_len;  int32_t service_kmin;
    int32_t size_lo_code, size_hi_code;  /* 10..100 uniform, 10..65 lowuniform, 25..100 highuniform (env.py:211-219) */
    const uint16_t* d_service_bracket;   /* optional u16[65]: #{i : service_cdf[i] <= b << 58} for b = 0..64 (search start) */
} vmgym_trace;

/* Optional per-step outputs; any pointer may be NULL. */
typedef struct vmgym_outputs {
    float* d_obs;            /* [n_envs, 3V+2P] float32 observation after the step (env.py:295-296) */
    double* d_reward;        /* [n_envs] (env.py:123-156) */
    uint8_t* d_terminated;   /* [n_envs] (env.py:160-163) */
    uint8_t* d_valid;        /* [n_envs, V] info["valid"] (env.py:69-74,91-94) */
    void* d_action;          /* [n_envs, V] actions chosen by a fused agent, element type = placement type */
    double* d_stats;         /* [n_envs, VMGYM_STATS] running sums over the steps of the call for Record.get_summary (src/record.py:
                                98-134): drop rate, waiting ratio, mean_p cpu, var_p cpu, mean_p mem, var_p mem, rejected actions,
                                steps, (mean_p cpu)^2, (mean_p mem)^2, target cpu mean, target mem mean, used PMs (rank), 3 reserved */
    /* Per-VM episode statistics of Record (src/record.py:34-96: pending rate, slowdown rate, lifetime of every VM that
     * occupied a slot), all three NULL or all three set.  The step kernels keep four clocks per slot and add a VM to
     * the histograms when it departs; vmgym_vmstats_finalize adds the VMs that still exist.  Zero the three buffers
     * when the envs are reset. */
    uint32_t* d_vm_slots;    /* [n_envs, V, 4] arrival step, first-placement step (0 = never), waiting steps after the
                                first placement, step of the pending suspension (0 = none) */
    uint32_t* d_vm_hist;     /* [n_envs, 2, VMGYM_VMSTAT_BINS] counts of rint(1000 * rate): [0] pending, [1] slowdown */
    uint64_t* d_vm_totals;   /* [n_envs, 4] VMs seen, VMs ever placed, sum of lifetimes, reserved */
    /* Optional mirror of d_obs for a caller whose observations live in HOST memory (pinned, device-mapped): with it set,
     * d_obs is treated as the persistent reference copy and an entry is stored to d_obs and to the mirror only when its value
     * changed — a quiet step changes nothing, so the PCIe traffic is a few entries per env instead of 4(3V+2P) bytes.
     * The mirror must hold the same contents as d_obs when the first such call is made (copy it once after reset). */
    float* d_obs_mirror;     /* [n_envs, 3V+2P] or NULL */
    /* 1: d_obs is the SAME buffer on every call for these envs (and is not written by anyone else): an env whose state did
     * not change since its observation row was last stored keeps that row (identical contents, no store).  The env records
     * carry the "changed since last stored" bit; vmgym_reset with a d_obs counts as a store.  0: every call stores. */
    int32_t obs_persistent;
    int32_t reserved0;
    /* vmgym_step only: after the step, the heuristic agent's act() on the NEW state (= on the float32 observation just produced:
     * firstfit.py:21-38 / bestfit.py:21-40), as [n_envs, V] actions of the placement type — what the caller's next
     * `agent.act(obs)` returns, computed while the record is still in shared memory (HostVecEnv's eager act).  NULL: off.
     * May alias the d_action argument of vmgym_step when that is of the placement type (an env reads its row before it writes it). */
    void* d_next_action;
    int32_t next_agent;      /* VMGYM_AGENT_FIRSTFIT / VMGYM_AGENT_BESTFIT */
    int32_t next_tiebreak;   /* VMGYM_TIE_* */
} vmgym_outputs;
#define VMGYM_STATS 16
#define VMGYM_VMSTAT_BINS 1024   /* rates are rounded to 3 decimals (record.py:61,79): bins 0..1000 are used */

const char* vmgym_last_error(void);
int vmgym_abi_version(void);

/* Record layout for a config.  Replaces the attribute set created by VmEnv.reset (env.py:186-208). */
int vmgym_get_layout(const vmgym_config* cfg, vmgym_layout* out);

/* VmEnv.reset (env.py:180-226): zero the state of the selected envs (all when d_env_mask is NULL, else
 * those with d_env_mask[i] != 0).  d_seeds (u64[n_envs], may be NULL) re-keys the env's streams and rewinds
 * the stream cursors; with d_seeds == NULL the cursors continue (reset() without seed, drlvmp.py:450-452).
 * d_obs (may be NULL) receives the reset observation. */
int vmgym_reset(const vmgym_config* cfg, void* d_state, int64_t n_envs, const uint8_t* d_env_mask,
                const uint64_t* d_seeds, int rewind_streams, float* d_obs, void* stream);

/* VmEnv.step (env.py:66-103) for n_envs envs: apply action[n_envs, V] sequentially per env with fp64
 * validity checks, service countdown and departures (_run_vms, :244-268), arrivals (_accept_vm_requests,
 * :271-293), metrics + reward (_process_action, :108-170), observation, termination flag, timestep += 1.
 * Out-of-range action values are invalid no-ops, as in validate() (:35-42). */
int vmgym_step(const vmgym_config* cfg, void* d_state, int64_t n_envs, const vmgym_trace* trace,
               const void* d_action, int action_dtype, const vmgym_outputs* out, void* stream);

/* Record's per-VM lists at this moment (record.py:34-96 evaluated on the episode so far): copies the histograms /
 * totals the step kernels accumulated for departed VMs into d_hist_out [n_envs, 2, VMGYM_VMSTAT_BINS] /
 * d_totals_out [n_envs, 4] and adds every VM that still occupies a slot (its sample list ends at the last step).
 * Does not modify the env state or the running buffers. */
int vmgym_vmstats_finalize(const vmgym_config* cfg, const void* d_state, int64_t n_envs, const uint32_t* d_vm_slots,
                           const uint32_t* d_vm_hist, const uint64_t* d_vm_totals, uint32_t* d_hist_out,
                           uint64_t* d_totals_out, void* stream);

/* Fused agent.act(obs) + env.step(action) (the body of Base.test's loop, src/agents/base.py:71-86) for the
 * heuristic agents, n_steps times per launch with the env state resident in shared memory.
 * FirstFitAgent.act: firstfit.py:21-38; BestFitAgent.act: bestfit.py:21-40.  Stops early at termination. */
int vmgym_agent_step(const vmgym_config* cfg, void* d_state, int64_t n_envs, const vmgym_trace* trace, int agent,
                     int tiebreak, int n_steps, const vmgym_outputs* out, void* stream);

/* The same fused act + step for SEVERAL independent env batches resident in one state buffer (the seeds / load points / batches
 * the reference's drivers fan out over processes, exp_performance.py:63-83), as ONE persistent launch: d_state holds
 * n_batches x envs_per_batch records (batch b = records [b * envs_per_batch, (b + 1) * envs_per_batch)), outputs are sized for
 * all of them, and the launch executes n_batch_steps consecutive "batch steps" — batch step k advances every env of batch
 * (first_batch + k) % n_batches by n_steps steps of the Base.test loop (base.py:71-86).  Equivalent to n_batch_steps calls of
 * vmgym_agent_step on the batches in rotation; a warp owns env index i of every batch (large shapes, pms > 253: a CTA owns the records
 * r = batch * envs_per_batch + i with r mod grid == its index), so every record is always stepped by the same warp / CTA in program
 * order, the grid stays resident and no launch boundary separates the steps.  Results are identical to the per-batch calls (tested).
 * n_batches * envs_per_batch must be < 2^31. */
int vmgym_agent_step_rotation(const vmgym_config* cfg, void* d_state, int64_t envs_per_batch, int32_t n_batches, int32_t first_batch,
                              int32_t n_batch_steps, const vmgym_trace* trace, int agent, int tiebreak, int n_steps,
                              const vmgym_outputs* out, void* stream);

/* agent.act(observation) on a batch of float32 observations (firstfit.py:21-38, bestfit.py:21-40,
 * drlvmp.py:549-617 heuristics).  d_action element type given by action_dtype. */
int vmgym_agent_act(const vmgym_config* cfg, int agent, int tiebreak, const float* d_obs, int64_t n_envs,
                    void* d_action, int action_dtype, void* stream);

/* VmEnv._get_obs (env.py:295-296). */
int vmgym_observe(const vmgym_config* cfg, const void* d_state, int64_t n_envs, float* d_obs, void* stream);

/* Host-resident observations (the caller of env.step of the reference holds numpy arrays, src/agents/base.py:71-86): keeps a mapped
 * pinned HOST copy `h_obs` of the device observation buffer `d_obs` current by storing only the entries whose bits differ from the
 * device-side shadow `d_shadow` (which is updated alongside); also forwards the step's reward [n_envs] f64 and done flags [n_envs] u8
 * to host buffers when those are given.  `n_floats` = n_envs * obs_dim, a multiple of 4; all three observation buffers 16-byte
 * aligned.  Runs on `stream`: HostVecEnv puts it on a side stream so that it is off the path to the next actions. */
int vmgym_obs_mirror_update(const float* d_obs, float* d_shadow, float* h_obs, int64_t n_floats, const double* d_reward, double* h_reward,
                            const uint8_t* d_terminated, uint8_t* h_terminated, int64_t n_envs, void* stream);

/* VmEnv.get_invalid_action_mask (env.py:45-53): d_mask[n_envs, V, A], 1 = invalid. */
int vmgym_invalid_action_mask(const vmgym_config* cfg, const void* d_state, int64_t n_envs, uint8_t* d_mask,
                              void* stream);

/* Network.get_action (src/agents/ppo.py:115-126) on the actor's logits d_logits[n_envs, V*A]: masked logits
 * (-1e7, :119), V categoricals of A, sample (Gumbel-max on a Philox stream keyed by seed/counter) or evaluate
 * d_action_in, sum of log-probs and of entropies per env.  The invalid-action mask (env.py:45-53) is built on the fly
 * from the env records d_state, or read as packed bits d_mask_in[n_envs, V, ceil(A/32)] (1 = invalid); masked == 0
 * disables it (get_invalid_action_mask(False)).  migration_ratio >= 0 applies PPOAgent.act's gating (ppo.py:153-155);
 * pass a negative value for training rollouts (ppo.py:196-197).  d_mask_out (may be NULL) receives the effective
 * packed mask for the update pass. */
int vmgym_policy_heads(const vmgym_config* cfg, const void* d_state, const uint32_t* d_mask_in, int masked,
                       const float* d_logits, int64_t n_envs, const void* d_action_in, int action_dtype,
                       float migration_ratio, uint64_t seed, uint64_t counter, void* d_action_out, float* d_logprob,
                       float* d_entropy, uint32_t* d_mask_out, void* stream);

/* Gradient of (sum_i g_logprob[i]*logprob[i] + g_entropy[i]*entropy[i]) w.r.t. the logits, for PPOAgent.update
 * (ppo.py:257-258,284-285); masked columns get zero gradient like the in-place fill of ppo.py:119. */
int vmgym_policy_heads_backward(const vmgym_config* cfg, const uint32_t* d_mask_in, int masked, const float* d_logits,
                                int64_t n_envs, const void* d_action_in, int action_dtype, const float* d_g_logprob,
                                const float* d_g_entropy, float* d_g_logits, void* stream);

/* The dense layers of PPOAgent.update (ppo.py:91-109 Network, :229-295) on tcgen05 tensor cores, bf16 operands, fp32 accumulate:
 *   C[M, N] = sum_k A(m, k) B(n, k)
 * Operand X is K-major (x_mn = 0: row-major [rows, K], row stride ldx) or MN-major (x_mn = 1: row-major [K, rows]) — the
 * forward pass (activations x nn.Linear weights) is K-major x K-major, the input-gradient GEMM K-major x MN-major, the
 * weight-gradient GEMM (contracting over samples) MN-major x MN-major; no transposed copies are needed.  Epilogue, fused:
 * + d_bias[N], act (bits 0-1: 0 none, 1 tanh, 2 relu; bit 2: the fp32 output takes the value BEFORE the activation; bit 3: the bf16
 * output is written as the split operand [hi | lo | hi] of vmgym_cast_split_bf16, segments N wide, ldc_bf16 >= 3 N),
 * * (1 - y^2) with d_mul_y bf16 [M, ldy] (tanh backward), fp32 output d_c_f32 (= or +=
 * when `accumulate`) and / or bf16 output d_c_bf16, and d_row_sum[M] (= or +=) sum_k A(m, k) — the bias gradient of the
 * weight-gradient GEMM.  ld* in elements, multiples of 8 for the operands; operand bases 16-byte aligned. */
int vmgym_tc_gemm(const void* d_a, int32_t a_mn, int64_t lda, const void* d_b, int32_t b_mn, int64_t ldb, int64_t M, int64_t N,
                  int64_t K, const float* d_bias, int32_t act, const void* d_mul_y, int64_t ldy, float* d_c_f32, int64_t ldc_f32,
                  int32_t accumulate, void* d_c_bf16, int64_t ldc_bf16, float* d_row_sum, void* stream);

/* fp32 [rows, cols] (row stride lds) -> bf16 [rows, cols_pad], zero padded (cols_pad % 8 == 0): the GEMMs' operand format. */
int vmgym_cast_pad_bf16(const float* d_src, int64_t rows, int64_t cols, int64_t lds, void* d_dst_bf16, int64_t cols_pad, void* stream);

/* fp32 [rows, cols] -> bf16 [rows, 3 cols_pad] = [hi | lo | hi] (order 0: activations) or [hi | hi | lo] (order 1: weights) with
 * hi = bf16(x), lo = bf16(x - hi): one bf16 GEMM over the concatenated K then yields x w to ~2^-16 relative.  Used for the
 * networks' first layer, whose inputs are raw observations (PM indices up to P + 1 next to sizes in [0, 1], env.py:295-296). */
int vmgym_cast_split_bf16(const float* d_src, int64_t rows, int64_t cols, int64_t lds, void* d_dst_bf16, int64_t cols_pad, int32_t order,
                          void* stream);

/* The critic's last layer Linear(hidden, 1) (ppo.py:95-101) on bf16 activations, and its backward: d_dz = dv w (1 - h^2) as
 * bf16 [rows, hidden], d_dw[hidden] += sum_m dv[m] h[m, :], *d_db += sum dv. */
int vmgym_value_head(const void* d_h_bf16, int64_t rows, int32_t hidden, const float* d_w, const float* d_b, float* d_out, void* stream);
int vmgym_value_head_backward(const void* d_h_bf16, int64_t rows, int32_t hidden, const float* d_w, const float* d_dv, void* d_dz_bf16,
                              float* d_dw, float* d_db, void* stream);

/* Per-sample PPO loss of ppo.py:259-282 (clipped surrogate, optionally clipped value loss, entropy bonus; means over
 * 1 / inv_n_total samples) and its derivatives w.r.t. each sample's summed log-prob (d_c_logprob) and value (d_c_value);
 * d_sums[0] += sum of log-ratios (the KL estimate of ppo.py:263), d_sums[1] += loss (fp64). */
int vmgym_ppo_loss(const float* d_new_logprob, const float* d_old_logprob, const float* d_adv, const float* d_entropy, const float* d_value,
                   const float* d_old_value, const float* d_return, int64_t n, float eps_clip, float ent_coef, float vf_coef,
                   int32_t vf_loss_clip, float inv_n_total, float* d_c_logprob, float* d_c_value, double* d_sums, void* stream);

/* The fused actor head evaluating STORED actions under stored masks (PPOAgent.update, ppo.py:257-258): as vmgym_policy_fused
 * with d_action_in, plus the softmax statistics of every (env, VM) row — d_stat_max / d_stat_sum [M, V]: row maximum and
 * sum of e^(z - max) over the valid columns — for vmgym_policy_fused_grad.  The epilogue visits only the columns that are
 * valid in at least one row of a warp (masked columns contribute exactly 0). */
int vmgym_policy_fused_eval(const void* d_h_bf16, const void* d_wpad_bf16, const float* d_bias_pad, const uint32_t* d_mask_bits,
                            const void* d_action_in, int64_t M, int64_t V, int64_t A, int64_t K, float* d_logprob, float* d_entropy,
                            float* d_stat_max, float* d_stat_sum, void* stream);

/* Backward of the fused actor head: recomputes the logits of each (128 envs x 1 VM) tile in tensor memory and, with the
 * forward's d_entropy / d_stat_max / d_stat_sum [M, V], writes d/dlogits of sum_e c_logprob[e] logprob(e) + c_entropy entropy(e)
 * in ONE pass over the accumulator as bf16 d_g_bf16[M, ldg] (column R v + a; zeros for masked and padding columns) — the
 * operand of the output layer's two backward GEMMs (vmgym_tc_gemm), and the only [samples, V x 128] tensor of the update that
 * reaches HBM. */
int vmgym_policy_fused_grad(const void* d_h_bf16, const void* d_wpad_bf16, const float* d_bias_pad, const uint32_t* d_mask_bits,
                            const void* d_action_in, int64_t M, int64_t V, int64_t A, int64_t K, const float* d_c_logprob,
                            float c_entropy, const float* d_entropy, const float* d_stat_max, const float* d_stat_sum, void* d_g_bf16,
                            int64_t ldg, void* stream);

/* The optimiser step of PPOAgent.update (ppo.py:143,284-287) on flat fp32 buffers of n elements:
 * nn.utils.clip_grad_norm_(max_grad_norm; <= 0 disables) on d_grad * grad_scale, then torch.optim.AdamW's update
 * (decoupled weight decay, bias-corrected moments; formulas of torch's reference implementation).  Device-resident control:
 * *d_step is the optimiser's step counter (incremented here), *d_skip != 0 (may be NULL) turns the whole call into a no-op —
 * the KL early stop of ppo.py:263-264 decided without a host round trip.  d_workspace: >= 1024 doubles of scratch;
 * d_grad_norm_out (may be NULL) receives the pre-clip global norm.  d_grad must be 16-byte aligned. */
int vmgym_adamw_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, float max_grad_norm, float grad_scale,
                     double* d_workspace, const int32_t* d_skip, int32_t* d_step, float* d_grad_norm_out, void* stream);

/* GAE advantages and returns (ppo.py:237-243) for time-major tensors [T, n_envs]; `dones` masks both the bootstrap
 * and the recursion. */
int vmgym_gae(const float* d_rewards, const float* d_values, const float* d_next_values, const uint8_t* d_dones, int32_t T,
              int64_t n_envs, float gamma, float lambda, float* d_advantages, float* d_returns, void* stream);

/* DRLVMPAgent._convert_action heuristics (src/agents/drlvmp.py:517-617) for one waiting VM per env:
 * d_choice[i] in {0 worst-fit, 1 min dot-product, 2 min L2 distance, 3 best-fit}, d_vm_index[i] = slot (or -1 to skip);
 * d_pm_out[i] = chosen PM, or -1 when worst-/best-fit find no PM that fits (the placement then stays WAIT). */
int vmgym_drlvmp_choice(const vmgym_config* cfg, const float* d_obs, const int32_t* d_vm_index, const int32_t* d_choice,
                        int64_t n_envs, int32_t* d_pm_out, void* stream);

/* One iteration of DRLVMPAgent.act (src/agents/drlvmp.py:504-530) after the head GEMMs, for every env with a k-th waiting VM
 * (k = *d_k + k_offset, the counter kept on the device so a captured graph can be replayed): d_heads [n, heads_ld >= (4 + 1) * atoms] = advantage atoms of the
 * four actions followed by the value atoms (Network.dist, :355-368) -> q-values -> argmax -> heuristic PM choice on the
 * working observation d_obs [n, 3V+2P] (as vmgym_drlvmp_choice) -> placement written into d_obs -> d_pre [n, hidden] (the
 * feature layer's pre-activation) corrected by the changed entry's weight column (d_w_cols [3V+2P, hidden], row j = column j
 * of the layer's weight) and d_feat = relu(d_pre); d_feat_split_bf16 (may be NULL) [n, 3 hidden] receives the same activations as
 * the [hi | lo | hi] bf16 split operand of the tensor-core head GEMMs (vmgym_tc_gemm).  d_order [n, V]: waiting slots first, in
 * slot order; d_n_wait [n]. */
int vmgym_drlvmp_iter(const vmgym_config* cfg, int32_t hidden, int32_t n_actions, int32_t atoms, const float* d_heads,
                      int32_t heads_ld, const float* d_support, float* d_obs, const int64_t* d_order, const int64_t* d_n_wait, const int64_t* d_k,
                      int32_t k_offset, const float* d_w_cols, float* d_pre, float* d_feat, void* d_feat_split_bf16, int64_t n_envs,
                      void* stream);

/* The actor's output layer (src/agents/ppo.py:103-109, nn.Linear(hidden, V*A)) on tcgen05 tensor cores:
 * d_c[M, N] (fp32, row stride ldc) = d_a[M, K] (bf16) . d_w[N, K]^T (bf16, the nn.Linear weight layout) + d_bias[N].
 * K must be a multiple of 8, operands 16-byte aligned.  bf16 inputs are a throughput mode (rollouts); parity of
 * logits with the fp32 reference is stated per math mode in DESIGN.md. */
int vmgym_linear_bf16(const void* d_a_bf16, const void* d_w_bf16, const float* d_bias, float* d_c, int64_t M, int64_t N,
                      int64_t K, int64_t ldc, void* stream);

/* Fused actor head (ppo.py:103-109 output layer + :115-126 get_action): logits = h . Wpad^T + bias_pad are produced
 * per (128 envs x 1 VM) tile in tensor memory and consumed there — masked (-1e7), sampled (or d_action_in evaluated),
 * log-prob and entropy per (env, VM) — so the [M, V*A] logits never touch HBM.  d_wpad_bf16[V*R, K] / d_bias_pad[V*R]
 * hold VM v's A rows at [R v, R v + A) (zero above; R = vmgym_policy_fused_rows(A, K)); d_mask_bits[M, V, 4] are the packed invalid bits written by
 * vmgym_policy_heads(d_logits = NULL, d_mask_out = ...), or NULL for no mask.  Requires action_dim <= 128 (u8 actions).
 * Sum d_logprob / d_entropy [M, V] over V for the per-env values of ppo.py:126. */
/* Rows per VM (R) of the padded output-layer operands of the fused head: d_wpad_bf16[V * R, K] / d_bias_pad[V * R] hold VM v's A
 * rows at [R v, R v + A) (zeros above), and vmgym_policy_fused_grad writes VM v's gradient columns at [R v, R v + R).
 * R = A rounded up to 16 for the persistent kernel (K <= 512), 128 for the tile-per-CTA fallback. */
int vmgym_policy_fused_rows(int64_t A, int64_t K);

int vmgym_policy_fused(const void* d_h_bf16, const void* d_wpad_bf16, const float* d_bias_pad, const uint32_t* d_mask_bits,
                       const void* d_action_in, int64_t M, int64_t V, int64_t A, int64_t K, uint64_t seed, uint64_t counter,
                       uint8_t* d_action_out, float* d_logprob, float* d_entropy, void* stream);

/* SumSegmentTree / MinSegmentTree of the prioritized replay buffer (src/segment_tree.py:8-142): d_*_tree are fp64
 * arrays of 2*capacity nodes (root at 1, leaves at capacity..2*capacity-1; capacity a power of two, :30-32), either may
 * be NULL.  Update = `tree[idx] = val` for a batch (:63-71, last write wins on duplicates); the root holds sum() / min(). */
int vmgym_segtree_update(double* d_sum_tree, double* d_min_tree, int64_t capacity, const int64_t* d_idx, const double* d_val,
                         int32_t n, void* stream);
/* SumSegmentTree.retrieve (src/segment_tree.py:103-118) for a batch of upper bounds -> leaf indices. */
int vmgym_segtree_retrieve(const double* d_sum_tree, int64_t capacity, const double* d_upper, int32_t n, int64_t* d_out,
                           void* stream);

/* Tuning knobs (process-wide): warps per CTA of the step kernels (0 = auto; 1..8: byte-placement shapes run that many
 * envs per CTA, at most 4; u16-placement shapes run ONE env per CTA with that many warps teaming on it) and bulk-async record copies
 * (cp.async.bulk): use_bulk_copy bit 0 = loads, bit 1 = stores, bit 2 = programmatic dependent launch of the step kernels
 * (default 7); bit 3 = team mode builds its fit table with the main warp alone; bit 4 = no double-buffered records in
 * multi-round launches, bit 5 = double-buffer records of any size (default: records up to 1 KB).  For experiments. */
int vmgym_set_tuning(int warps_per_cta, int use_bulk_copy);

/* PrioritizedReplayBuffer.sample_batch (src/agents/drlvmp.py:178-241, src/segment_tree.py:35-62,103-118): stratified
 * proportional sampling from the sum tree with caller-supplied uniforms d_u[batch] in [0, 1) (the reference draws
 * random.uniform(a, b) = a + (b - a) * u), plus the importance weights.  len = number of stored transitions. */
int vmgym_per_sample(const double* d_sum_tree, const double* d_min_tree, int64_t capacity, int64_t len, int32_t batch,
                     const double* d_u, double beta, int64_t* d_idx_out, double* d_weight_out, void* stream);

/* Categorical-DQN target projection (src/agents/drlvmp.py:676-699): d_next_dist [n, atoms] (the target net's distribution of
 * the double-DQN action), d_reward f32[n], d_done i32[n], d_support f32[atoms] -> d_proj f32[n, atoms]. */
int vmgym_c51_project(const float* d_next_dist, const float* d_reward, const int32_t* d_done, const float* d_support,
                      float gamma, float v_min, float v_max, int32_t atoms, int64_t n, float* d_proj, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VMGYM_H_ */
