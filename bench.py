#!/usr/bin/env python
"""bench.py — env-steps/sec of the batched VM-placement hot path (BASELINE.json metric).

Workload (N=1): BASELINE.json configs[1] — config/100.yml, best-fit evaluation, 100 PMs / 300 VM slots, uniform VM
sizes, 4096 envs per GPU.  One "step" = agent.act + env.step for every env of a 4096-env batch (one launch of the fused
kernel, observation written to HBM every step as the gym API does); launches rotate over --batches independent batches
so that the records of a launch are never L2-resident (inputs larger than L2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--envs E]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (see DESIGN.md §7 for every field).  `--impl reference` times the CPU oracle port of
the reference path (the reference is pure Python and is not on the GPU box) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200"), os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import yaml  # noqa: E402

METRIC = "env-steps/sec (100 PMs, best-fit act+step, whole job)"
UNIT = "env-steps/s"
WARM_STEPS = 3000          # reach saturation (~300/300 slots occupied) before timing, SURVEY §8d
PERIOD = 1000              # service_length of config/100.yml: one departure wave per period


def load_env_cfg():
    cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "100.yml")))["environment"]
    cfg["reward_function"] = "wr"        # main.py:94 CLI default overrides the YAML (main.py:34)
    return cfg


def algorithmic_bytes(P, V):
    """state record S = 16P + 5V + 48 read and written once, observation 4(3V+2P) written, plus
    16 B of per-step scalars (reward, done, trace words)."""
    S = 16 * P + 5 * V + 48
    return 2 * S + 4 * (3 * V + 2 * P) + 16


# --------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, one env per process like the reference's drivers (exp.py:1)
# --------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, warm, steps = args
    import vmoracle as vo
    cfg = load_env_cfg()
    cfg["seed"] = int(seed)
    env = vo.OracleVmEnv(vo.OracleConfig(**cfg), trace_steps=warm + steps + 8, trace_adm=4 * (warm + steps) + 4096)
    env.rollout(vo.AGENT_BESTFIT, warm)
    t0 = time.perf_counter()
    n, _ = env.rollout(vo.AGENT_BESTFIT, steps)
    return n, time.perf_counter() - t0


def cpu_arm(steps_per_env=2000, warm=WARM_STEPS, cores=None):
    import vmoracle as vo
    vo.build()
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(s, warm, steps_per_env) for s in range(cores)])
    wall = time.perf_counter() - t0
    total = sum(n for n, _ in res)
    slowest = max(dt for _, dt in res)
    return dict(value=total / slowest, unit=UNIT, cores=cores, kind="port",
                sample=f"{cores} envs x {steps_per_env} best-fit act+step after {warm} warm-up steps, one env per "
                       f"process (C oracle port of env.py/bestfit.py; the Python reference measured 62-68 steps/s/process, "
                       f"BASELINE.md §3)", per_process=total / cores / slowest, wall_s=wall)


# --------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.samples.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for n, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from vmgym import Config, VecVmEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from vmgym.sharding import bind_to_gpu_numa_node
    prev_affinity = bind_to_gpu_numa_node(local) if world > 1 else None
    cfg = load_env_cfg()
    E = args.envs
    if args.bulk is not None or args.warps:
        from vmgym import _native as nv
        nv.lib().vmgym_set_tuning(int(args.warps), 7 if args.bulk is None else int(args.bulk))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- batches: NB independent batches of E envs each, resident in ONE state buffer; a timed "step" is one fused
    # act+step pass over ONE batch, consecutive steps rotate over the batches.  Working set per step = 14 MB records read +
    # 14 MB written (+ the observation rows that changed); NB steps touch NB x 28 MB >> 126 MB L2 before a batch comes round
    # again, so every step reads its records from HBM (inputs larger than L2; no flush kernel inside the timed region).
    # Service times are Poisson(1000): departures come in waves one service period apart and a step costs more inside a
    # wave.  Batch b is therefore warmed up to phase b * PERIOD / NB of the period, so the rotation samples all phases.
    # Envs shard contiguously over ranks: batch b of rank g holds global env ids [(b * world + g) * E, +E).
    NB = args.batches
    seeds = np.concatenate([cfg["seed"] + (b * world + rank) * E + np.arange(E, dtype=np.int64) for b in range(NB)])
    vec = VecVmEnv(Config(**cfg), NB * E, device=dev, rng="philox", seeds=seeds)
    P, V, D = vec.P, vec.V, vec.obs_dim
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2
    quiet = dict(want_obs=False, want_action=False, want_valid=False)
    for b in range(NB):
        vec.agent_step("bestfit", n_steps=WARM_STEPS + (b * PERIOD) // NB, envs=(b * E, (b + 1) * E), **quiet)

    K, W, R = args.steps, max(3, args.warmup), max(1, args.replays)
    # ---- timed region A (value): the rotation as ONE persistent launch of K batch steps (vmgym_agent_step_rotation): a warp
    # owns env index i of every batch, the grid stays resident across the K steps.  W untimed warm-up steps of the same call
    # (after one full untimed rotation, which also stores every observation row once), then R timed replays of the K-step
    # launch, each with its own event pair; the rotation continues from replay to replay (all phases sampled).
    nxt = vec.agent_step_rotation("bestfit", E, NB, first_batch=0)
    nxt = vec.agent_step_rotation("bestfit", E, W, first_batch=nxt)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    # R + 1 boundary events: replay r is timed from boundary r to boundary r + 1 (nothing between two replays is left out of the
    # timed spans, and only one event record sits between consecutive launches)
    evb = [torch.cuda.Event(enable_timing=True) for _ in range(R + 1)]
    barrier()
    evb[0].record()
    for r in range(R):
        nxt = vec.agent_step_rotation("bestfit", E, K, first_batch=nxt)
        evb[r + 1].record()
    barrier()
    replay_ms = np.array([evb[r].elapsed_time(evb[r + 1]) for r in range(R)], dtype=np.float64)
    total_ms = float(np.median(replay_ms))

    # observation rows the step kernel actually stores: a row is re-stored only when its env's state changed since the row was
    # written; counted on the device, outside the timed region, over one full rotation (every batch once)
    before = vec.obs.clone()
    nxt = vec.agent_step_rotation("bestfit", E, NB, first_batch=nxt)
    rows_changed = int((vec.obs != before).any(dim=1).sum().item())
    del before
    obs_rows_frac = rows_changed / float(NB * E)

    # ---- timed region A' (per_launch): the round-1 protocol — the same K batch steps as K separate launches of the fused step
    # kernel (one CUDA graph, programmatic dependent launch between them), median of 10 replays ----
    def rotate(n, start):
        for k in range(n):
            b_ = (start + k) % NB
            vec.agent_step("bestfit", 1, want_obs=True, want_action=False, want_valid=False, envs=(b_ * E, (b_ + 1) * E))

    with vec._on_device():
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            rotate(NB, nxt)
        torch.cuda.current_stream(dev).wait_stream(side)
        launch_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(launch_graph):
            rotate(K, nxt)
    launch_graph.replay()
    barrier()
    pl = []
    for _ in range(10):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); launch_graph.replay(); s1.record()
        torch.cuda.synchronize()
        pl.append(s0.elapsed_time(s1))
    per_launch_ms = float(np.median(pl)) / K

    # ---- timed region B (rollout): same work, 100 steps per launch with the state resident in shared memory ----
    chunk, n_chunks = 100, 10                   # 1000 steps = one full service period
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(n_chunks):
        vec.agent_step("bestfit", chunk, want_obs=True, want_action=False, want_valid=False, envs=(0, E))
    r1.record()
    barrier()
    rollout_ms = r0.elapsed_time(r1)

    # ---- timed region C (e2e): the reference-facing loop `action = agent.act(obs); obs, r, done = env.step(action)` with
    # HOST (pinned) observation / action / reward buffers: every step moves obs + action over PCIe in both directions
    # inside the timed region.  HostVecEnv splits the E envs into groups on separate streams so that one group's
    # device->host copies overlap another group's host->device copies (PCIe is full duplex). ----
    from vmgym.host_vec import HostVecEnv
    Ke = max(3, min(K, 100))
    hv = HostVecEnv(Config(**cfg), E, groups=args.e2e_groups, device=dev, rng="philox", agent="bestfit",
                    seeds=cfg["seed"] + 3 * 10**6 + rank * E + np.arange(E, dtype=np.int64))
    # like the batches of the main metric, the groups are warmed to different phases of the service period (departure waves)
    e2e_warm = [WARM_STEPS + (gi * PERIOD) // len(hv.groups) for gi in range(len(hv.groups))]
    hv.fast_forward(e2e_warm)
    hv.run_pipelined(max(3, 10 * W))        # untimed warm-up: graph capture + upload, clocks; the first dozens of steps of a fresh loop run slow
    barrier()
    t0 = time.perf_counter()
    hv.run_pipelined(Ke)
    barrier()
    e2e_s = time.perf_counter() - t0
    # bytes the step kernel actually stores to the host observation buffer per step (only entries that changed), counted on
    # the device outside the timed region
    changed = 0
    for _ in range(10):
        prev = [g.vec.obs.clone() for g in hv.groups]
        hv.act(); hv.step()
        changed += sum(int((g.vec.obs != p_).sum().item()) for g, p_ in zip(hv.groups, prev))
    d2h_obs = changed / 10 * 4
    # the same loop as round 1 ran it: the agent's input re-uploaded from the host observation buffer every step (19 MB H2D per
    # 4096 envs) — what a caller pays when it hands act() observations of its own instead of the env's buffer
    hv1 = HostVecEnv(Config(**cfg), E, groups=args.e2e_groups, device=dev, rng="philox", agent="bestfit", resident_obs=False,
                     seeds=cfg["seed"] + 3 * 10**6 + rank * E + np.arange(E, dtype=np.int64))
    hv1.fast_forward(e2e_warm)
    hv1.run_pipelined(max(3, 10 * W))
    barrier()
    t0 = time.perf_counter()
    hv1.run_pipelined(Ke)
    barrier()
    e2e1_s = time.perf_counter() - t0
    h2d = hv.h2d_bytes_per_step
    d2h = int(E * V * hv.action.element_size() + d2h_obs + E * 9)
    hv.close(); hv1.close()
    del hv, hv1
    clocks = sampler.finish() if sampler else None

    # ---- extra A: PPO training throughput at BASELINE config 4's per-GPU share: 8192 envs, rollout T = batch_size = 100
    # (config/100.yml), 4 sequential minibatches of 25 steps, k_epochs 4, NCCL gradient all-reduce per optimiser step ----
    extras_on = set() if args.no_extras else set(args.extras.split(","))
    ppo = None
    if "ppo" in extras_on:
        from vmgym.ppo import PPOAgent, PPOConfig
        Np, Tp = args.ppo_envs, args.ppo_steps
        vp = VecVmEnv(Config(**cfg), Np, device=dev, rng="philox", seeds=cfg["seed"] + 2 * 10**6 + rank * Np + np.arange(Np, dtype=np.int64))
        torch.set_float32_matmul_precision("high")      # as the reference does (main.py:45): TF32 for the fp32 layers
        agent_p = PPOAgent(vp, PPOConfig(hidden_size=512, batch_size=Tp, minibatch_size=max(1, Tp // 4), episodes=1, env_chunk=32768,
                                         masked=True, kl_max=1e9, fused_rollout=True))
        if world > 1:
            for p_ in agent_p.model.parameters():
                dist.broadcast(p_.data, 0)
            agent_p.weights_changed()
        vp.agent_step("bestfit", n_steps=WARM_STEPS, **quiet)          # saturated envs, as in training after the first episode steps
        agent_p.learn(episodes=1, max_updates=1, reset=False)           # warm-up (allocations, kernel plans)
        ppo_s = []
        for _ in range(2):                                              # two timed rollout + update rounds, the faster one reported
            barrier()
            t0 = time.perf_counter()
            agent_p.learn(episodes=1, max_updates=1, reset=False)
            barrier()
            ppo_s.append(time.perf_counter() - t0)
        ppo = {"seconds": min(ppo_s), "seconds_all": ppo_s, "env_steps": Np * Tp, "T": Tp, "envs": Np}
        del vp, agent_p
        # PPO evaluation rollouts (BASELINE config 3 shape): mask + gating + fused tcgen05 actor head + env.step, E envs
        ve = VecVmEnv(Config(**cfg), E, device=dev, rng="philox", seeds=cfg["seed"] + 6 * 10**6 + rank * E + np.arange(E, dtype=np.int64))
        ve.agent_step("bestfit", n_steps=WARM_STEPS, **quiet)
        agent_e = PPOAgent(ve, PPOConfig(hidden_size=512, masked=True, migration_ratio=0.002))
        agent_e.rollout(3)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        agent_e.rollout(20)
        e1.record()
        barrier()
        ppo["eval_ms_per_step"] = e0.elapsed_time(e1) / 20
        del agent_e, ve

    # ---- extra B: the synthetic 1000-PM shape (BASELINE config 5: highuniform sizes at 100 % load, V = 3P) ----
    s1000 = None
    if "s1000" in extras_on:
        kw1000 = dict(cfg, pms=1000, vms=3000, sequence="highuniform", arrival_rate=1000 / 0.625 / cfg["service_length"])
        E1, NB1, K1 = 1024, 4, 50
        # the headline's protocol at this shape: NB1 phase-staggered batches of E1 envs (4 x 37 MB of records > 126 MB L2), K1 batch
        # steps per rotation launch (team-mode kernel, records shared out over the CTAs), median of 7 replays
        v1 = VecVmEnv(Config(**kw1000), NB1 * E1, device=dev, rng="philox",
                      seeds=cfg["seed"] + 4 * 10**6 + rank * NB1 * E1 + np.arange(NB1 * E1, dtype=np.int64))
        v1.agent_step("bestfit", n_steps=WARM_STEPS, **quiet)
        for b in range(1, NB1):
            v1.agent_step("bestfit", n_steps=(b * PERIOD) // NB1, envs=(b * E1, (b + 1) * E1), **quiet)
        nx1 = v1.agent_step_rotation("bestfit", E1, NB1 + 3, first_batch=0)
        barrier()
        r1 = []
        for _ in range(7):
            s0_, s1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0_.record()
            nx1 = v1.agent_step_rotation("bestfit", E1, K1, first_batch=nx1)
            s1_.record()
            torch.cuda.synchronize()
            r1.append(s0_.elapsed_time(s1_) / K1)
        barrier()
        s1000 = {"envs_per_gpu": E1, "ms_per_step": float(np.median(r1)), "batches": NB1, "steps_per_launch": K1}
        # round-1 protocol for comparison: one launch per step on one batch (batch 0: the start of a departure wave)
        s0_, s1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0_.record()
        for _ in range(10):
            v1.agent_step("bestfit", 1, want_obs=True, want_action=False, want_valid=False, envs=(0, E1))
        s1_.record()
        barrier()
        s1000["per_launch_ms_per_step"] = s0_.elapsed_time(s1_) / 10
        del v1
        v1 = VecVmEnv(Config(**kw1000), E1, device=dev, rng="philox", seeds=cfg["seed"] + 4 * 10**6 + rank * E1 + np.arange(E1, dtype=np.int64))
        v1.agent_step("bestfit", n_steps=WARM_STEPS, **quiet)
        # DRL-VMP rollout at this shape (drlvmp.py:504-512: one network evaluation + heuristic per WAITING VM, sequentially)
        from vmgym.drlvmp import DRLVMPAgent, DRLVMPConfig
        torch.set_float32_matmul_precision("high")      # as the reference does (main.py:45)
        ag1 = DRLVMPAgent(v1, DRLVMPConfig(hidden_size=512))
        ag1.eval()
        for mode in ("tc", "torch"):
            # "tc": every GEMM of the per-VM iteration on the hand-written tcgen05 kernel with split-bf16 operands (fp32-accurate:
            # the q-value argmax equals the reference's); "torch": the same loop on cuBLAS TF32 GEMMs (library baseline)
            ag1.gemm = mode
            o1 = v1.observe()
            o1, *_ = v1.step(ag1.act(o1), want_valid=False)            # graph capture / allocations
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                o1, *_ = v1.step(ag1.act(o1), want_valid=False)
            barrier()
            s1000["drlvmp_s_per_step" + ("" if mode == "tc" else "_cublas")] = (time.perf_counter() - t0) / 2
        del v1, ag1

    # ---- extra C: the small shape of BASELINE configs[0] (config/10.yml, first-fit) at 2^20 envs per GPU ----
    s10 = None
    if "s10" in extras_on:
        cfg10 = yaml.safe_load(open(os.path.join(ROOT, "configs", "10.yml")))["environment"]
        cfg10["reward_function"] = "wr"
        E10 = 1 << 20
        v10 = VecVmEnv(Config(**cfg10), E10, device=dev, rng="philox", seeds=cfg10["seed"] + 5 * 10**6 + rank * E10 + np.arange(E10, dtype=np.int64))
        v10.agent_step("firstfit", n_steps=WARM_STEPS, **quiet)
        v10.agent_step("firstfit", 1, want_obs=True, want_action=False, want_valid=False)
        barrier()
        s0_, s1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0_.record()
        for _ in range(10):
            v10.agent_step("firstfit", 1, want_obs=True, want_action=False, want_valid=False)
        s1_.record()
        barrier()
        s10 = {"envs_per_gpu": E10, "ms_per_step": s0_.elapsed_time(s1_) / 10}
        del v10

    if world > 1:
        t = torch.tensor([total_ms, rollout_ms, e2e_s, per_launch_ms, ppo["seconds"] if ppo else 0.0,
                          ppo["eval_ms_per_step"] if ppo else 0.0, e2e1_s, s1000["ms_per_step"] if s1000 else 0.0,
                          s1000["drlvmp_s_per_step"] if s1000 else 0.0, s10["ms_per_step"] if s10 else 0.0,
                          float(np.percentile(replay_ms, 10)), float(np.percentile(replay_ms, 90)), float(replay_ms.mean())],
                         dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tl = t.tolist()
        total_ms, rollout_ms, e2e_s, per_launch_ms = tl[:4]
        e2e1_s = tl[6]
        if s1000:
            s1000["ms_per_step"], s1000["drlvmp_s_per_step"] = tl[7], tl[8]
        if s10:
            s10["ms_per_step"] = tl[9]
        if ppo:
            ppo["seconds"], ppo["eval_ms_per_step"] = tl[4], tl[5]
        p10_ms, p90_ms, mean_ms = tl[10], tl[11], tl[12]
        fr = torch.tensor([obs_rows_frac], dtype=torch.float64, device=dev)
        dist.all_reduce(fr)
        obs_rows_frac = fr.item() / world
    else:
        p10_ms, p90_ms, mean_ms = float(np.percentile(replay_ms, 10)), float(np.percentile(replay_ms, 90)), float(replay_ms.mean())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    # Algorithmic bytes per env-step: the fused heuristic step reads and writes the state record once,
    # B_fused = 2 S + 16 with S = 16 P + 5 V + 48, plus the observation bytes the kernel really stores (rows of envs whose
    # state changed; unchanged rows are kept in the persistent observation buffer), measured above.
    S = 16 * P + 5 * V + 48
    B_fused = 2 * S + 16
    obs_stored = obs_rows_frac * 4 * D
    B = B_fused + obs_stored
    step_s = (total_ms / K) * 1e-3
    achieved = B * E / step_s / 1e9
    value = world * E * K / (total_ms * 1e-3)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_rotation_traffic.json")))
        if int(tj.get("envs_per_batch", 0)) == E:
            traffic, traffic_src = float(tj["dram_bytes_per_batch_step"]), tj.get("source")
    except Exception:
        pass
    extras = {}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (Philox arrivals, uniform sizes)",
        "config": {"workload": "config/100.yml best-fit evaluation: 100 PMs, 300 VM slots, uniform sizes, "
                               f"{E} envs per GPU per batch, reward wr, saturated after {WARM_STEPS} warm-up steps",
                   "envs_per_gpu": E, "batches": NB,
                   "l2": f"inputs larger than L2: consecutive steps rotate over {NB} independent {E}-env batches "
                         f"({NB} x {2 * 3456 * E / 1e6:.0f} MB of records touched between two visits of a batch, L2 = 126 MB)",
                   "phase_sampling": f"batch b warmed up to phase b*{PERIOD}/{NB} of the service period (departure waves)",
                   "timing": f"one step = one fused act+step pass over one {E}-env batch; K consecutive steps = ONE persistent launch "
                             f"(vmgym_agent_step_rotation); {R} replays of the K-step launch back to back, replay r timed between boundary events r and r + 1; "
                             "ms_per_step = median replay / K",
                   "rng": "philox", "tiebreak": "stable",
                   "obs_written": "into the env's persistent observation buffer; rows of envs whose state did not change are kept, not re-stored",
                   "state_types": "f64 PM accumulators, u8 placements / size codes, u16 runtimes, f32 observation"},
        "gpu_launches": R,
        "timing_stats": {"replays": R, "steps_per_replay": K, "median_ms": total_ms, "p10_ms": p10_ms, "p90_ms": p90_ms,
                         "mean_ms": mean_ms, "spread": (p90_ms - p10_ms) / total_ms},
        "e2e": {"value": world * E * Ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": Ke, "groups": args.e2e_groups, "obs_reupload_value": world * E * Ke / e2e1_s,
                "obs_reupload_h2d_bytes_per_step": E * D * 4 + E * V,
                "path": "HostVecEnv: action = agent.act(); obs, reward, done = env.step(action) with HOST (pinned) action / observation / "
                        "reward / done buffers. Per step the actions cross PCIe twice (agent kernel -> host buffer, host buffer -> step "
                        "kernel) and the step kernel stores reward, done and the CHANGED observation entries to the host buffers "
                        "(d2h_bytes_per_step = actions + measured changed entries + reward/done). The host observation buffer is a mirror "
                        "the env keeps current, so act() on it reads the identical device copy instead of re-uploading 4(3V+2P) bytes per "
                        "env; obs_reupload_value = the round-1 loop that re-uploads it every step (what act(obs) costs for a caller's own "
                        "array). The step enqueue ends with the agent's act() on the observation it just produced (eager_act: one graph "
                        "launch and one host round trip per step, the act itself computed by the step kernel on the record it still holds, "
                        "vmgym_outputs.d_next_action; act() then only waits for the actions, which still travel device -> host "
                        "-> device and may be replaced by the caller before step(); the host observation mirror is brought up to date by a kernel "
                        "on a side stream, vmgym_obs_mirror_update, off the path to the next actions). The env groups are warmed to different phases of the service period, like the batches of the main metric; "
                        "%d untimed steps of the same loop precede the %d timed ones" % (max(3, 10 * W), Ke)},
        "per_launch": {"value": world * E / (per_launch_ms * 1e-3), "unit": UNIT, "ms_per_step": per_launch_ms,
                       "note": "round-1 protocol: the same rotation as K separate launches of the fused step kernel in one CUDA graph "
                               "(programmatic dependent launch), median of 10 replays",
                       "roofline_frac": B * E / (per_launch_ms * 1e-3) / 1e9 / peak},
        "rollout": {"value": world * E * chunk * n_chunks / (rollout_ms * 1e-3), "unit": UNIT,
                    "steps_per_launch": chunk, "note": "same fused kernel, env state resident in shared memory across steps"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "vmgym::step_kernel<u8,100,300,bestfit/wr/philox> (rotation launch)",
                     "bytes_per_env_step": B, "B_fused": B_fused, "obs_bytes_stored_per_env_step": obs_stored,
                     "obs_rows_stored_frac": obs_rows_frac,
                     "frac_B_fused_only": B_fused * E / step_s / 1e9 / peak,
                     "frac_if_every_obs_row_counted": (B_fused + 4 * D) * E / step_s / 1e9 / peak,
                     "note": "achieved = (B_fused + measured stored observation bytes) x envs per step / (median replay / K); "
                             "B_fused = 2(16P+5V+48)+16 (SURVEY §8d); frac_if_every_obs_row_counted is the round-1 accounting, secondary",
                     "peak_source": peak_src},
        "clocks": clocks,
    }
    extras["per_launch_M"] = round(world * E / (per_launch_ms * 1e-3) / 1e6, 1)
    extras["rollout100_M"] = round(world * E * chunk * n_chunks / (rollout_ms * 1e-3) / 1e6, 1)
    if ppo:
        out["ppo_train"] = {"value": world * ppo["env_steps"] / ppo["seconds"], "unit": "PPO train env-steps/s",
                            "config": f"config/100.yml, {ppo['envs']} envs/GPU (BASELINE config 4's per-GPU share), rollout T={ppo['T']} "
                                      "(reference batch_size), k_epochs=4, 4 sequential minibatches, H=512, "
                                      "NCCL gradient all-reduce per optimiser step when N > 1; the faster of two timed rollout + update rounds",
                            "seconds": ppo["seconds"], "seconds_all": ppo["seconds_all"], "T": ppo["T"], "envs_per_gpu": ppo["envs"]}
        out["ppo_eval"] = {"value": world * E / (ppo["eval_ms_per_step"] * 1e-3), "unit": UNIT, "ms_per_step": ppo["eval_ms_per_step"],
                           "config": f"config/100.yml PPO evaluation rollouts, {E} envs/GPU, reference-shaped MLP (H=512, random init: the "
                                     "100-PM weights are not shipped), masked, migration_ratio 0.002, fused tcgen05 actor head (bf16)"}
        extras[f"ppo_train_T{ppo['T']}_M"] = round(out["ppo_train"]["value"] / 1e6, 3)
        extras["ppo_eval_M"] = round(out["ppo_eval"]["value"] / 1e6, 2)
    if s1000:
        B1 = 2 * (16 * 1000 + 6 * 3000 + 48) + 16
        out["s1000"] = {"value": world * s1000["envs_per_gpu"] / (s1000["ms_per_step"] * 1e-3), "unit": UNIT,
                        "ms_per_step": s1000["ms_per_step"], "envs_per_gpu": s1000["envs_per_gpu"],
                        "roofline_frac_B_fused": B1 * s1000["envs_per_gpu"] / (s1000["ms_per_step"] * 1e-3) / 1e9 / peak,
                        "drlvmp_rollout": {"value": world * s1000["envs_per_gpu"] / s1000["drlvmp_s_per_step"], "unit": UNIT,
                                           "cublas_tf32_value": world * s1000["envs_per_gpu"] / s1000["drlvmp_s_per_step_cublas"],
                                           "note": "DRLVMPAgent.act (H=512 dueling C51 net, one network evaluation + heuristic per waiting "
                                                   "VM, ~1000 strictly sequential evaluations per step) + env.step; value: head GEMMs on the "
                                                   "hand-written tcgen05 kernel, split-bf16 operands (fp32-accurate argmax); cublas_tf32_value: "
                                                   "the same loop on cuBLAS TF32 GEMMs — a ~1 GFLOP GEMM per launch is latency-bound and the "
                                                   "TMA / TMEM set-up of a tcgen05 kernel (~10 us) costs more than the library's mma.sync kernel"},
                        "per_launch_value": world * s1000["envs_per_gpu"] / (s1000["per_launch_ms_per_step"] * 1e-3),
                        "config": "synthetic 1000 PMs / 3000 VM slots, highuniform sizes, arrival 1.6 (100 % load), fused best-fit act+step, "
                                  f"team-mode kernel; value: rotation launch over {s1000['batches']} phase-staggered batches of "
                                  f"{s1000['envs_per_gpu']} envs ({s1000['steps_per_launch']} batch steps per launch, median of 7 replays; "
                                  "records of one rotation > L2); per_launch_value: round-1 protocol, one launch per step on one batch"}
        extras["s1000_M"] = round(out["s1000"]["value"] / 1e6, 2)
        extras["s1000_drlvmp_k"] = round(out["s1000"]["drlvmp_rollout"]["value"] / 1e3, 1)
    if s10:
        B10 = 2 * (16 * 10 + 5 * 30 + 48) + 16
        out["s10"] = {"value": world * s10["envs_per_gpu"] / (s10["ms_per_step"] * 1e-3), "unit": UNIT, "ms_per_step": s10["ms_per_step"],
                      "envs_per_gpu": s10["envs_per_gpu"],
                      "roofline_frac_B_fused": B10 * s10["envs_per_gpu"] / (s10["ms_per_step"] * 1e-3) / 1e9 / peak,
                      "config": "config/10.yml shape (10 PMs / 30 VM slots, BASELINE configs[0]), first-fit fused act+step, 2^20 envs per GPU, "
                                "one launch per step"}
        extras["s10_G"] = round(out["s10"]["value"] / 1e9, 3)
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)             # the CPU baseline uses every host core
    out["config"]["numa_bound"] = prev_affinity is not None
    if not args.no_cpu:
        out["cpu_baseline"] = cpu_arm(steps_per_env=args.cpu_steps)
    # compact copy of the secondary numbers, last in the line (survives a truncated tail) and inside the parsed `config`
    out["config"]["extras"] = extras
    out["extras"] = extras
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = max(1, args.steps), max(0, args.warmup)
    per_step = max(100, args.cpu_steps // max(1, K))
    vals = []
    base = None
    for k in range(W + K):
        r = cpu_arm(steps_per_env=per_step)
        if k >= W:
            vals.append(r["value"])
            base = r
    value = float(np.mean(vals))
    base["value"] = value
    cfg = load_env_cfg()
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
           "steps": K, "warmup": W, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic (numpy PCG64 traces, uniform sizes)",
           "config": {"workload": "config/100.yml best-fit evaluation: 100 PMs, 300 VM slots, uniform sizes, one env per "
                                  f"host process x {base['cores']} processes, reward wr, saturated after {WARM_STEPS} warm-up steps",
                      "arrival_rate": cfg["arrival_rate"]},
           "cpu_baseline": base,
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU (per batch)")
    ap.add_argument("--ppo-envs", type=int, default=8192, help="envs per GPU of the ppo_train extra")
    ap.add_argument("--ppo-steps", type=int, default=100, help="rollout length T of the ppo_train extra (config/100.yml batch_size)")
    ap.add_argument("--replays", type=int, default=50, help="timed replays of the K-step rotation launch (median reported)")
    ap.add_argument("--warps", type=int, default=0, help="vmgym_set_tuning warps per CTA, 0 = auto (experiments)")
    ap.add_argument("--bulk", type=int, default=None, help="vmgym_set_tuning use_bulk_copy bits (experiments)")
    ap.add_argument("--e2e-groups", type=int, default=4, help="env groups (streams) of the host-buffer e2e loop")
    ap.add_argument("--batches", type=int, default=20, help="independent env batches the timed launches rotate over")
    ap.add_argument("--cpu-steps", type=int, default=6000, help="timed CPU steps per env in the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the PPO / 1000-PM / 10-PM extras")
    ap.add_argument("--extras", default="ppo,s1000,s10", help="comma list of the extras to run (ppo, s1000, s10)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e loop (experiments)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
