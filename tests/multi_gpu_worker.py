"""Worker of tests/test_multi_gpu.py: one process per GPU under torch.distributed.run (NCCL), world size >= 2.

A. env sharding (SURVEY §4 tier 3 / §8e): global env ids [0, N) sharded contiguously over the ranks produce byte-identical
   records, observations and rewards to the same envs stepped by ONE rank — fused best-fit steps, then external-action steps.
B. PPO data parallelism (ppo.py:229-295 + NCCL gradient all-reduce): on a fixed rollout, the gradient of a minibatch computed
   on env shards and all-reduced == the gradient of the whole batch on one GPU (global advantage normalisation), and a full
   sharded update() leaves every rank with the same parameters as the single-GPU large-batch update.
Rank 0 prints one JSON line with the measured differences; any mismatch raises (non-zero exit)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from vmgym import Config, VecVmEnv
    from vmgym.agents import FirstFitAgent
    from vmgym.ppo import PPOAgent, PPOConfig
    from vmgym.sharding import shard_range, shard_seeds
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}

    # ---------------- A. env shards == one rank ----------------
    kw = dict(pms=100, vms=300, arrival_rate=1.8182, service_length=60, training_steps=10000, eval_steps=100000,
              reward_function="wr", allow_null_action=True, seed=5)
    N = 64 * world + 3                                   # uneven shards
    lo, hi = shard_range(N, rank, world)

    def run(vec):
        vec.agent_step("bestfit", n_steps=150, want_action=False, want_valid=False)
        ff = FirstFitAgent(vec)
        obs = vec.obs
        rsum = torch.zeros(vec.num_envs, dtype=torch.float64, device=dev)
        for _ in range(25):
            obs, r, term, _, _ = vec.step(ff.act(obs))
            rsum += r
        return vec.state.clone(), vec.obs.clone(), rsum

    mine = VecVmEnv(Config(**kw), hi - lo, device=dev, rng="philox", seeds=shard_seeds(kw["seed"], N, rank, world))
    st, ob, rs = run(mine)
    sizes = [shard_range(N, r, world)[1] - shard_range(N, r, world)[0] for r in range(world)]
    pad = max(sizes)

    def gather(x):
        buf = torch.zeros((pad,) + x.shape[1:], dtype=x.dtype, device=dev)
        buf[:x.shape[0]] = x
        outl = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(outl, buf)
        return torch.cat([o[:n] for o, n in zip(outl, sizes)])

    st_all, ob_all, rs_all = gather(st), gather(ob), gather(rs)
    if rank == 0:
        full = VecVmEnv(Config(**kw), N, device=dev, rng="philox", seeds=kw["seed"] + np.arange(N))
        st1, ob1, rs1 = run(full)
        assert torch.equal(st_all, st1), "sharded env records differ from the single-rank run"
        assert torch.equal(ob_all, ob1) and torch.equal(rs_all, rs1)
        out["env_records_identical"] = True
        out["env_count"] = N

    # ---------------- B. sharded PPO gradients == large-batch gradients ----------------
    kw2 = dict(pms=10, vms=30, arrival_rate=0.4, service_length=30, training_steps=400, eval_steps=1000,
               reward_function="wr", allow_null_action=True, seed=3)
    Ng, T = 16 * world, 8
    prev_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    for math in ("fp32", "bf16"):
        # fp32: torch autograd layers; bf16: the hand-written tensor-core forward / backward (same bf16 roundings per sample on
        # every rank, so shards and the whole batch differ by summation order only — incl. the split-K atomics)
        pc = PPOConfig(hidden_size=64, batch_size=T, minibatch_size=4, k_epochs=2, env_chunk=4096, kl_max=None, update_math=math)
        rtol, atol_rel = (1e-4, 1e-5) if math == "fp32" else (2e-3, 2e-4)
        torch.manual_seed(1234)                                              # identical initial weights on every rank
        full = VecVmEnv(Config(**kw2), Ng, device=dev, rng="philox", seeds=kw2["seed"] + np.arange(Ng))
        a_full = PPOAgent(full, pc)
        a_full.data_parallel = False
        for p in a_full.model.parameters():
            dist.broadcast(p.data, 0)
        a_full.weights_changed()
        w0 = a_full._flat.clone()
        # the same rollout on every rank: deterministic env + Philox-sampled actions from identical weights
        full.agent_step("firstfit", n_steps=60, want_action=False, want_valid=False)
        obs = full.observe().clone()
        buf = dict(obs=[], next_obs=[], action=[], mask=[], logprob=[], reward=[], done=[])
        with torch.no_grad():
            for _ in range(T):
                logits = a_full.model.actor(obs).contiguous()
                action, logprob, _, mask = a_full._heads(logits, -1.0, want_mask=True)
                nobs, reward, term, _, _ = full.step(action, want_valid=False)
                for k, v in (("obs", obs), ("next_obs", nobs), ("action", action), ("mask", mask), ("logprob", logprob),
                             ("reward", reward.float()), ("done", full.terminated_u8)):
                    buf[k].append(v.clone())
                obs = nobs.clone()
        buf = {k: torch.stack(v) for k, v in buf.items()}
        chk = buf["obs"].double().sum() + buf["action"].double().sum() + buf["logprob"].double().sum()
        chks = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(chks, chk)
        assert all(float(c) == float(chks[0]) for c in chks), "ranks disagree on the rollout"

        lo, hi = shard_range(Ng, rank, world)
        shard = {k: v[:, lo:hi].contiguous() for k, v in buf.items()}
        sh_env = VecVmEnv(Config(**kw2), hi - lo, device=dev, rng="philox", seeds=shard_seeds(kw2["seed"], Ng, rank, world))
        a_sh = PPOAgent(sh_env, pc)
        a_sh._flat.copy_(w0)
        a_sh.weights_changed()

        def one_minibatch(agent, b, w):
            from vmgym.ppo import gae
            Tn, Nn = b["reward"].shape
            tc = agent._tc_network()
            with torch.no_grad():
                if tc is not None:
                    values = tc.values(b["obs"]).reshape(Tn, Nn)
                    nvals = tc.values(b["next_obs"]).reshape(Tn, Nn)
                    o, m = tc.cast_obs(b["obs"]).reshape(Tn, Nn, tc.Dx), agent._mask4(b["mask"])
                else:
                    values = agent.model.get_value(b["obs"].reshape(Tn * Nn, -1)).reshape(Tn, Nn)
                    nvals = agent.model.get_value(b["next_obs"].reshape(Tn * Nn, -1)).reshape(Tn, Nn)
                    o, m = b["obs"], b["mask"]
                adv, ret = gae(b["reward"], values, nvals, b["done"], pc.gamma, pc.lamda)
            t0, t1 = 0, pc.minibatch_size
            n_mb = (t1 - t0) * Nn
            adv_mb = agent._normalise_advantages(adv[t0:t1], w)
            agent._minibatch_backward(o[t0:t1].reshape(n_mb, -1), b["action"][t0:t1].reshape(n_mb, -1),
                                      m[t0:t1].reshape(n_mb, agent.V, m.shape[-1]), b["logprob"][t0:t1].reshape(-1), adv_mb,
                                      values[t0:t1].reshape(-1), ret[t0:t1].reshape(-1), n_mb * w)
            if w > 1:
                dist.all_reduce(agent._flat_grad)
            return agent._flat_grad.clone()

        g_sh = one_minibatch(a_sh, shard, world)
        g_full = one_minibatch(a_full, buf, 1)
        d = (g_sh - g_full).abs().max().item()
        scale = g_full.abs().max().item()
        assert torch.allclose(g_sh, g_full, rtol=rtol, atol=atol_rel * scale), f"{math}: sharded gradient differs: max |d| {d} at scale {scale}"
        out[f"{math}_grad_max_abs_diff"], out[f"{math}_grad_scale"] = d, scale

        # full update: sharded (collectives inside update) vs large batch on one GPU
        a_sh._flat.copy_(w0); a_full._flat.copy_(w0)
        a_sh.weights_changed(); a_full.weights_changed()
        a_sh.update(**shard)
        a_full.update(**buf)
        dp = (a_sh._flat - a_full._flat).abs().max().item()
        moved = (a_full._flat - w0).abs().max().item()
        assert moved > 0 and dp <= (2e-2 if math == "fp32" else 1e-1) * moved + 1e-7, f"{math}: parameters after the sharded update differ: {dp} (update size {moved})"
        ws = [torch.zeros_like(a_sh._flat) for _ in range(world)]
        dist.all_gather(ws, a_sh._flat)
        assert all(torch.equal(w, ws[0]) for w in ws), "ranks ended the update with different parameters"
        out[f"{math}_update_param_max_abs_diff"], out[f"{math}_update_size"] = dp, moved
    torch.backends.cuda.matmul.allow_tf32 = prev_tf32
    if rank == 0:
        print("MULTI_GPU_OK " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
