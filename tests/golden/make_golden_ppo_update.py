#!/usr/bin/env python
"""Golden vectors for PPOAgent.update (src/agents/ppo.py:229-295) and the actor / critic forward of the shipped 10-PM
weights, produced by the UNMODIFIED reference classes.  Build container only (needs /root/reference).

    python tests/golden/make_golden_ppo_update.py   ->  tests/golden/ppo_update.npz, tests/golden/ppo_weights10.npz

The rollout batch is collected exactly like PPOAgent.learn does (ppo.py:190-214: mask -> get_action -> env.step, buffers of
batch_size = 100 steps of ONE env, an episode boundary inside the batch), then `agent.update(...)` runs unmodified.  What the
update computes internally is observed from outside, without touching the reference file:
  * `model.get_value` / `model.get_action` are wrapped on the INSTANCE to log their outputs (values, new log-probs, entropies);
  * `torch.Tensor.backward`, `nn.utils.clip_grad_norm_` and `optimizer.step` are wrapped to log the loss, the pre-clip
    gradient norm and the number of optimiser steps;
  * a `sys.settrace` hook on the frame of `update` copies its locals `advantages` and `returns`.

ppo_update.npz, per case `<c>.`:
  cfg_json, ppo_cfg_json, mask u8[T, ceil(V*A/8)] (np.packbits of the bool masks), action i16[T,V], obs f32[T,D], next_obs f32[T,D],
  logprob f32[T], reward f32[T], done u8[T], values f32[T], next_values f32[T], advantages f32[T], returns f32[T],
  attempt_epoch i32[K], attempt_mb i32[K], attempt_kl f64[K], attempt_stepped u8[K]  (every minibatch the update looked at),
  attempt_newlogprob f32[K, mb], attempt_entropy f32[K, mb],
  step_loss f64[S], step_grad_norm f64[S]  (one per optimiser step),
  and per parameter tensor `<c>.p.<name>.`: idx i64[n] (flat sample positions), pre f32[n], post f32[n], delta_l2 f64, delta_sum f64.
ppo_weights10.npz: the 12 tensors of weights-10/ppo-wr.pt (keys without the `_orig_mod.` prefix); for ppo-ut / ppo-kl the
  output layer (`<name>.actor.4.weight/bias`, `<name>.critic.4.*`) and, for the two anchor observations of SURVEY §8c, the
  hidden activations feeding it (`<name>.h_actor`, `<name>.h_critic`), the logits and the values of the reference Network.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VMGYM_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "stubs"), REF, HERE]
os.environ.setdefault("OMP_NUM_THREADS", "1")

import numpy as np  # noqa: E402
import torch  # noqa: E402
from make_golden import base_cfg  # noqa: E402

CASES = {
    # default hyper-parameters of config/10.yml (lr 5e-5): all 4 x 4 minibatches step
    "upd_default": dict(env=dict(reward_function="wr", arrival_rate=0.3, service_length=30, training_steps=60), ppo={}),
    # a learning rate large enough that the KL early stop (ppo.py:263-264) fires after the first steps
    "upd_klstop": dict(env=dict(reward_function="wr", arrival_rate=0.3, service_length=30, training_steps=60), ppo=dict(lr=2e-3)),
    # vf_loss_clip off, unmasked, other reward, odd minibatch split 100 = 3 x 30 + 10
    "upd_unclipped_ut": dict(env=dict(reward_function="ut", arrival_rate=0.2, service_length=40, training_steps=10000),
                             ppo=dict(vf_loss_clip=False, masked=False, minibatch_size=30, k_epochs=2)),
}


def seed_all(seed):
    torch.manual_seed(seed)          # main.py:40-45
    random.seed(seed)
    np.random.seed(seed)


def sample_idx(n, k=2048):
    if n <= k:
        return np.arange(n, dtype=np.int64)
    return np.unique(np.linspace(0, n - 1, k).astype(np.int64))


def run_case(name, spec):
    from vmenv.envs.env import VmEnv
    from vmenv.envs.config import Config
    from src.agents.ppo import PPOAgent, PPOConfig
    import yaml
    cfg = base_cfg("10")
    cfg.update(spec["env"])
    agent_cfg = yaml.safe_load(open(os.path.join(REF, "config", "10.yml")))["agents"]["ppo"]
    agent_cfg.update(spec["ppo"])
    seed_all(cfg["seed"])
    env = VmEnv(Config(**cfg))
    agent = PPOAgent(env, PPOConfig(**agent_cfg))
    agent.load_model(os.path.join(REF, "weights-10", spec.get("weights", "ppo-wr.pt")))
    agent.eval(False)
    c = agent.config
    T, V, A, D = c.batch_size, cfg["vms"], env.action_dim, env.observation_space.shape[0]
    # ---- rollout buffers exactly as ppo.py:178-214 ----
    mask_b = torch.zeros((T, V, A), dtype=bool)
    action_b = torch.zeros((T, V), dtype=int)
    obs_b = torch.zeros(T, D, dtype=torch.float32)
    nobs_b = torch.zeros(T, D, dtype=torch.float32)
    lp_b = torch.zeros(T, dtype=torch.float32)
    rew_b = torch.zeros(T, dtype=torch.float32)
    done_b = torch.zeros(T, dtype=int)
    i, ep = 0, 0
    while i < T:
        obs, _ = env.reset(seed=env.config.seed + ep)
        obs = torch.tensor(obs, dtype=torch.float32)
        done = False
        while not done and i < T:
            invalid_mask = torch.tensor(env.get_invalid_action_mask(c.masked))
            action, logprob, _ = agent.model.get_action(obs.unsqueeze(0), invalid_mask=invalid_mask)
            action = torch.flatten(action)
            next_obs, reward, done, _, _ = env.step(action.cpu().numpy())
            next_obs = torch.tensor(next_obs, dtype=torch.float32)
            mask_b[i], action_b[i], obs_b[i], nobs_b[i] = invalid_mask, action, obs, next_obs
            lp_b[i], rew_b[i], done_b[i] = logprob.item(), reward, done
            i += 1
            obs = next_obs
        ep += 1
    pre = {k: v.detach().clone() for k, v in agent.model.state_dict().items()}

    # ---- observe the unmodified update from outside ----
    log = dict(values=[], attempts=[], losses=[], norms=[], steps=0)
    model = agent.model
    orig_value, orig_action = model.get_value, model.get_action

    def get_value(obs):
        out = orig_value(obs)
        log["values"].append(out.detach().clone())
        return out

    def get_action(obs, action=None, invalid_mask=None):
        out = orig_action(obs, action=action, invalid_mask=invalid_mask)
        log["attempts"].append(dict(newlogprob=out[1].detach().clone(), entropy=out[2].detach().clone(), stepped=0))
        return out

    model.get_value, model.get_action = get_value, get_action
    orig_backward, orig_clip, orig_step = torch.Tensor.backward, torch.nn.utils.clip_grad_norm_, agent.optimizer.step

    def backward(self, *a, **k):
        log["losses"].append(float(self.detach().double()))
        return orig_backward(self, *a, **k)

    def clip(params, max_norm, *a, **k):
        n = orig_clip(params, max_norm, *a, **k)
        log["norms"].append(float(n))
        return n

    def step(*a, **k):
        log["steps"] += 1
        log["attempts"][-1]["stepped"] = 1
        return orig_step(*a, **k)

    grabbed = {}

    def tracer(frame, event, arg):
        if frame.f_code.co_name != "update" or "ppo.py" not in frame.f_code.co_filename:
            return None

        def local(frame, event, arg):
            if "returns" in frame.f_locals and "returns" not in grabbed:
                grabbed["advantages"] = frame.f_locals["advantages"].detach().clone()
                grabbed["returns"] = frame.f_locals["returns"].detach().clone()
            return local
        return local

    torch.Tensor.backward, torch.nn.utils.clip_grad_norm_, agent.optimizer.step = backward, clip, step
    sys.settrace(tracer)
    try:
        agent.update(mask_b, action_b, obs_b, nobs_b, lp_b, rew_b, done_b)
    finally:
        sys.settrace(None)
        torch.Tensor.backward, torch.nn.utils.clip_grad_norm_ = orig_backward, orig_clip
    post = agent.model.state_dict()

    # minibatch bookkeeping: sequential minibatches, a break ends the epoch (ppo.py:251-252,263-264)
    mbs = [list(range(s, min(T, s + c.minibatch_size))) for s in range(0, T, c.minibatch_size)]
    ep_i, mb_i = 0, 0
    ae, am, akl, ast = [], [], [], []
    mbw = max(len(m) for m in mbs)
    anl = np.zeros((len(log["attempts"]), mbw), np.float32)
    aen = np.zeros((len(log["attempts"]), mbw), np.float32)
    for k, at in enumerate(log["attempts"]):
        idx = mbs[mb_i]
        lr = at["newlogprob"] - lp_b[idx]
        ae.append(ep_i); am.append(mb_i); akl.append(float(-lr.double().mean())); ast.append(at["stepped"])
        anl[k, :len(idx)] = at["newlogprob"].numpy()
        aen[k, :len(idx)] = at["entropy"].numpy()
        if at["stepped"] and mb_i + 1 < len(mbs):
            mb_i += 1
        else:
            ep_i, mb_i = ep_i + 1, 0
    assert log["steps"] == len(log["losses"]) == len(log["norms"]) == sum(ast)
    out = {f"{name}.cfg_json": json.dumps(cfg), f"{name}.ppo_cfg_json": json.dumps(agent_cfg),
           f"{name}.mask": np.packbits(mask_b.numpy().reshape(T, -1), axis=1), f"{name}.action": action_b.numpy().astype(np.int16),
           f"{name}.obs": obs_b.numpy(), f"{name}.next_obs": nobs_b.numpy(), f"{name}.logprob": lp_b.numpy(),
           f"{name}.reward": rew_b.numpy(), f"{name}.done": done_b.numpy().astype(np.uint8),
           f"{name}.values": log["values"][0].flatten().numpy(), f"{name}.next_values": log["values"][1].flatten().numpy(),
           f"{name}.advantages": grabbed["advantages"].numpy(), f"{name}.returns": grabbed["returns"].numpy(),
           f"{name}.attempt_epoch": np.array(ae, np.int32), f"{name}.attempt_mb": np.array(am, np.int32),
           f"{name}.attempt_kl": np.array(akl, np.float64), f"{name}.attempt_stepped": np.array(ast, np.uint8),
           f"{name}.attempt_newlogprob": anl, f"{name}.attempt_entropy": aen,
           f"{name}.step_loss": np.array(log["losses"], np.float64), f"{name}.step_grad_norm": np.array(log["norms"], np.float64),
           f"{name}.weights": spec.get("weights", "ppo-wr.pt")}
    for k in pre:
        kk = k[len("_orig_mod."):] if k.startswith("_orig_mod.") else k
        a, b = pre[k].flatten().numpy(), post[k].detach().flatten().numpy()
        idx = sample_idx(a.size)
        d = (b.astype(np.float64) - a.astype(np.float64))
        out[f"{name}.p.{kk}.idx"], out[f"{name}.p.{kk}.pre"], out[f"{name}.p.{kk}.post"] = idx, a[idx], b[idx]
        out[f"{name}.p.{kk}.delta_l2"], out[f"{name}.p.{kk}.delta_sum"] = float(np.sqrt((d * d).sum())), float(d.sum())
    print(f"{name}: attempts={len(ae)} steps={log['steps']} kl={np.round(akl, 5).tolist()} done_at={np.nonzero(done_b.numpy())[0].tolist()} "
          f"loss0={log['losses'][0]:.6f} norm0={log['norms'][0]:.4f}", flush=True)
    return out


def weights_and_anchors():
    """The shipped 10-PM weights for the GPU box + the reference Network's own forward values on two observations."""
    from vmenv.envs.env import VmEnv
    from vmenv.envs.config import Config
    from src.agents.ppo import Network
    from src.agents.firstfit import FirstFitAgent
    cfg = base_cfg("10")
    cfg["reward_function"] = "wr"
    env = VmEnv(Config(**cfg))
    ff = FirstFitAgent(env)
    obs0, _ = env.reset(seed=1)
    obs = obs0
    for _ in range(2000):                                      # the golden run of SURVEY §8c
        obs, _, _, _, _ = env.step(ff.act(obs))
    x = torch.tensor(np.stack([obs0, obs]), dtype=torch.float32)
    out = {"anchor_obs": x.numpy(), "anchor_mask": np.packbits(env.get_invalid_action_mask(True).reshape(-1))}
    for fname in ("ppo-wr.pt", "ppo-ut.pt", "ppo-kl.pt"):
        tag = fname[:-3]
        net = Network(110, env.action_space, 512, torch.float32)
        sd = {k[len("_orig_mod."):]: v for k, v in torch.load(os.path.join(REF, "weights-10", fname), map_location="cpu").items()}
        net.load_state_dict(sd)
        with torch.no_grad():
            h_a, h_c = net.actor[:4](x), net.critic[:4](x)
            out[f"{tag}.logits"], out[f"{tag}.values"] = net.actor(x).numpy(), net.get_value(x).flatten().numpy()
            out[f"{tag}.h_actor"], out[f"{tag}.h_critic"] = h_a.numpy(), h_c.numpy()
            # Network.get_action with the stored mask and the unmasked greedy action as the evaluated action (ppo.py:115-126)
            m = torch.tensor(env.get_invalid_action_mask(True))
            lg = net.actor(x[1:2]).clone()
            act = torch.where(m, torch.full_like(lg.reshape(30, 12), -1e9), lg.reshape(30, 12)).argmax(1)[None]
            _, lp, ent = net.get_action(x[1:2], action=act, invalid_mask=m)
            out[f"{tag}.eval_action"], out[f"{tag}.eval_logprob"], out[f"{tag}.eval_entropy"] = act.numpy().astype(np.int16), lp.numpy(), ent.numpy()
        keys = list(sd) if fname == "ppo-wr.pt" else ["actor.4.weight", "actor.4.bias", "critic.4.weight", "critic.4.bias"]
        for k in keys:
            out[f"{tag}.{k}"] = sd[k].numpy()
        print(f"{tag}: value {out[f'{tag}.values'].tolist()} logits[0:4] {out[f'{tag}.logits'][0, :4].tolist()} sum {out[f'{tag}.logits'].sum(1).tolist()}")
    path = os.path.join(HERE, "ppo_weights10.npz")
    np.savez_compressed(path, **out)
    print(f"-> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    out = {}
    for name, spec in CASES.items():
        out.update(run_case(name, spec))
    out["torch_version"] = torch.__version__
    path = os.path.join(HERE, "ppo_update.npz")
    np.savez_compressed(path, **out)
    print(f"-> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")
    weights_and_anchors()


if __name__ == "__main__":
    main()
