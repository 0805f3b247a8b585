"""-m gpu: the eval-mode `info` dict and the host-side logs of the VmEnv facade against golden vectors recorded from the
unmodified reference (tests/golden/make_golden_info.py): all keys of env.py:298-318 (+ action / valid, env.py:91-94),
info["timestep"] as the reference reports it (before the clock advances), vm_planned_runtime after every step and the
final vm_arrival_steps lists; plus a loop that reads exactly what Base.record_testing_step reads (base.py:131-149)."""
import json
import os

import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


def _load(name):
    z = np.load(os.path.join(gu.GOLDEN_DIR, "info.npz"))
    return {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + ".")}


@pytest.mark.parametrize("name", ["info_busy_suspend", "info_p37_v70"])
def test_facade_info_matches_reference(name):
    from vmgym import Config, VmEnv
    fx = _load(name)
    cfg = json.loads(str(fx["cfg_json"]))
    T, V = fx["actions"].shape
    env = VmEnv(Config(**cfg), trace_steps=T + 8, max_admissions=4096)
    env.eval()
    obs, info = env.reset(seed=cfg["seed"])
    want_keys = json.loads(str(fx["info_keys"]))
    assert list(info.keys()) == want_keys[:-2]                         # reset: the 18 keys of _get_info, in order
    assert info["timestep"] == fx["timestep"][0]
    assert np.array_equal(env.vm_planned_runtime, fx["planned"][0])
    rec = {k: [] for k in ("cpu", "memory", "used_pm", "vm_placements", "waiting_ratio", "actions", "rewards", "dropped_requests",
                           "total_requests", "target_cpu_mean", "target_memory_mean", "served_requests", "suspended", "placed", "rank")}
    done = False
    t = 0
    while not done:
        obs, reward, done, trunc, info = env.step(fx["actions"][t].astype(np.int64))
        assert set(info.keys()) == set(want_keys) and len(info) == 20
        # what Base.record_testing_step reads (src/agents/base.py:131-149)
        rec["cpu"].append(info["cpu"]); rec["memory"].append(info["memory"])
        rec["used_pm"].append(len(info["cpu"]) - np.count_nonzero(info["cpu"]))
        rec["vm_placements"].append(info["vm_placement"]); rec["waiting_ratio"].append(info["waiting_ratio"])
        rec["actions"].append(info["action"]); rec["rewards"].append(reward)
        rec["dropped_requests"].append(info["dropped_requests"]); rec["total_requests"].append(info["total_requests"])
        vm_arrival_steps = info["vm_arrival_steps"]
        rec["target_cpu_mean"].append(info["target_cpu_mean"]); rec["target_memory_mean"].append(info["target_memory_mean"])
        rec["served_requests"].append(int(info["served_requests"]))
        _ = info["total_cpu_requested"], info["total_memory_requested"]
        rec["suspended"].append(info["suspend_actions"]); rec["placed"].append(info["place_actions"]); rec["rank"].append(info["rank"])
        assert info["timestep"] == fx["timestep"][t + 1], f"info timestep @ {t}"
        assert np.array_equal(env.vm_planned_runtime, fx["planned"][t + 1]), f"vm_planned_runtime @ {t}"
        t += 1
    assert t == T
    assert np.array_equal(rec["used_pm"], fx["used_pm"]) and np.array_equal(rec["served_requests"], fx["served"])
    assert np.array_equal(rec["rank"], fx["rank"])
    assert np.array_equal(np.asarray(rec["waiting_ratio"]), fx["waiting_ratio"])
    off = fx["arrival_off"]
    want = [fx["arrival_flat"][off[v]:off[v + 1]].tolist() for v in range(V)]
    assert vm_arrival_steps is env.vm_arrival_steps                    # live list, as in the reference (env.py:307)
    assert env.vm_arrival_steps == want
    # a reset re-creates both logs (env.py:203-205)
    env.reset(seed=cfg["seed"])
    assert env.vm_arrival_steps == [[] for _ in range(V)] and not env.vm_planned_runtime.any()


def test_vec_arrival_step_view():
    import torch
    from vmgym import Config, VecVmEnv
    vec = VecVmEnv(Config(pms=10, vms=30, arrival_rate=0.5, service_length=30, training_steps=1000, eval_steps=1000), 5, rng="philox")
    with pytest.raises(RuntimeError):
        vec.vm_arrival_step()
    vec.enable_vm_stats()
    vec.reset(seed=3)
    vec.agent_step("firstfit", n_steps=200, want_action=False, want_valid=False)
    a = vec.vm_arrival_step().cpu().numpy()
    occupied = vec.vm_placement.cpu().numpy() <= 10
    assert (a[occupied] >= 2).all() and (a[occupied] <= 201).all()
    rem = vec.vm_remaining_runtime
    assert rem.dtype == torch.int32 and int(rem.min()) >= 0
