"""Shared replay of the golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the
unmodified reference).  The same function checks the CPU oracle (CPU suite) and the CUDA env (-m gpu suite):
anything with the reference-shaped surface reset/step/get_invalid_action_mask/state works."""
import glob
import hashlib
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixture_names():
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if not n.startswith(("drlvmp", "record", "info", "ppo"))]      # own layouts (make_golden_drlvmp.py, make_golden_record.py)


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["cfg"] = json.loads(str(d["cfg_json"]))
    d["name"] = name
    return d


def replay(fx, make_env, reward_rtol=1e-9, max_steps=None, check_masks=True):
    """Drive `make_env(cfg_dict, trace_steps, trace_adm)` with the fixture's recorded actions and compare
    everything the reference exposed.  Integer / byte / fp64-accumulator state must match bit-for-bit; rewards
    match to `reward_rtol` (the only non-bit-exact quantity: kl goes through log() and BLAS in the reference)."""
    cfg = fx["cfg"]
    P, V = cfg["pms"], cfg["vms"]
    T = fx["action"].shape[0] if max_steps is None else min(max_steps, fx["action"].shape[0])
    env = make_env(cfg, T + 8, int(fx["counters"][:, 1].max()) + 8)
    if int(fx["eval_mode"]):
        env.eval()
    obs, _ = env.reset(seed=int(fx["seed"]))
    sha = hashlib.sha256()
    sha.update(np.asarray(obs, np.float32).tobytes())
    resets = {int(t): int(s) for t, s in zip(fx["reset_at"], fx["reset_seed"])}
    mask_at = {int(t): i for i, t in enumerate(fx["mask_steps"])}

    def check_state(t):
        s = env.state()
        assert np.array_equal(s["vm_placement"], fx["placement"][t].astype(np.int64)), f"placement @ {t}"
        assert np.array_equal(np.rint(s["vm_cpu"] * 100).astype(np.int64), fx["vm_cpu_code"][t]), f"vm_cpu @ {t}"
        assert np.array_equal(s["vm_cpu"], fx["vm_cpu_code"][t] / 100.0), f"vm_cpu value @ {t}"
        assert np.array_equal(s["vm_memory"], fx["vm_mem_code"][t] / 100.0), f"vm_memory @ {t}"
        assert s["cpu"].tobytes() == fx["cpu"][t].tobytes(), f"fp64 cpu accumulators @ {t}"
        assert s["memory"].tobytes() == fx["memory"][t].tobytes(), f"fp64 memory accumulators @ {t}"
        assert int(np.dot(s["vm_remaining_runtime"], np.arange(1, V + 1))) == int(fx["remaining_sum"][t]), f"remaining @ {t}"
        assert np.array_equal(s["vm_suspended"], fx["suspended"][t]), f"suspended @ {t}"
        return s

    check_state(0)
    for t in range(T):
        if check_masks and t in mask_at:
            m = np.asarray(env.get_invalid_action_mask(True)).astype(bool)
            want = np.unpackbits(fx["mask_bits"][mask_at[t]])[: m.size].reshape(m.shape).astype(bool)
            assert np.array_equal(m, want), f"invalid-action mask @ {t}"
        action = fx["action"][t].astype(np.int64)
        obs, reward, term, trunc, info = env.step(action)
        sha.update(np.asarray(obs, np.float32).tobytes())
        assert np.array_equal(np.asarray(info["valid"]).astype(np.uint8), fx["valid"][t]), f"valid @ {t}"
        want_r = fx["reward"][t]
        assert abs(reward - want_r) <= reward_rtol * max(abs(want_r), 1e-300) or reward == want_r, \
            f"reward @ {t}: {reward!r} vs {want_r!r}"
        assert bool(term) == bool(fx["terminated"][t]), f"terminated @ {t}"
        assert trunc is False or trunc == 0
        if (t + 1) in resets and t + 1 < fx["action"].shape[0]:
            s = env.state()
            c = fx["counters"][t]
            got = [s[k] for k in ("timestep", "total_requests", "served_requests", "suspend_action", "place_action",
                                  "dropped_requests")]
            assert got == c.tolist(), f"counters @ {t}"
            seed = resets[t + 1]
            obs, _ = env.reset(seed=None if seed < 0 else seed)
            sha.update(np.asarray(obs, np.float32).tobytes())
            check_state(t + 1)
            continue
        s = check_state(t + 1)
        got = [s[k] for k in ("timestep", "total_requests", "served_requests", "suspend_action", "place_action",
                              "dropped_requests")]
        assert got == fx["counters"][t].tolist(), f"counters @ {t}: {got} vs {fx['counters'][t].tolist()}"
        sc = fx["scalars"][t]
        assert abs(s["total_cpu_requested"] - sc[0]) <= 1e-9 * max(1.0, abs(sc[0]))
        assert abs(s["total_memory_requested"] - sc[1]) <= 1e-9 * max(1.0, abs(sc[1]))
        assert abs(s["waiting_ratio"] - sc[2]) <= 1e-12
        assert abs(s["target_cpu_mean"] - sc[3]) <= 1e-12 and abs(s["target_memory_mean"] - sc[4]) <= 1e-12
    if T == fx["action"].shape[0]:
        assert sha.hexdigest() == str(fx["obs_sha256"]), "float32 observation stream digest"
        assert np.array_equal(env.state()["vm_remaining_runtime"], fx["remaining_final"])
    return env
