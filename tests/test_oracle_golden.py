"""CPU suite: pins the oracle (oracle/vmenv_oracle.c) against the reference.

 * per-step fixtures recorded from the unmodified reference classes (tests/golden/make_golden.py);
 * the reference's own published result rows (data/exp_performance_small/summary.csv:3-4,
   data/exp_performance/summary.csv:3-4; per-seed values from SURVEY §8c KAT-1..4);
 * numpy's reduction / sort orders the reference silently depends on.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import golden_util as gu
import vmoracle as vo


def _make_oracle(cfg, trace_steps, trace_adm):
    return vo.OracleVmEnv(vo.OracleConfig(**cfg), trace_steps=trace_steps, trace_adm=trace_adm)


@pytest.mark.parametrize("name", gu.fixture_names())
def test_oracle_replays_reference_fixture(name):
    gu.replay(gu.load(name), _make_oracle)


@pytest.mark.parametrize("name", [n for n in gu.fixture_names() if "firstfit" in n or "bestfit" in n])
def test_oracle_agents_reproduce_reference_actions(name):
    """firstfit.py:21-38 / bestfit.py:21-40: same float32 observation -> same action vector, every step."""
    fx = gu.load(name)
    cfg = fx["cfg"]
    P, V = cfg["pms"], cfg["vms"]
    tie = vo.TIE_NUMPY_INTROSORT if str(fx["tiebreak"]) == "numpy_introsort" else vo.TIE_STABLE
    T = fx["action"].shape[0]
    for t in range(T):
        obs = np.concatenate([fx["placement"][t].astype(np.float64), fx["vm_cpu_code"][t] / 100.0,
                              fx["vm_mem_code"][t] / 100.0, fx["cpu"][t], fx["memory"][t]]).astype(np.float32)
        a = vo.firstfit_act(P, V, obs) if str(fx["agent"]) == "firstfit" else vo.bestfit_act(P, V, obs, tie)
        assert np.array_equal(a, fx["action"][t].astype(np.int64)), f"{name} step {t}"


def test_np_sum_order():
    """np.sum of float64 == 0 + pairwise(a) with 8 accumulators / 128-blocks (drives `ut`, `kl`, target means)."""
    rng = np.random.default_rng(0)
    for n in list(range(0, 40)) + [100, 127, 128, 129, 130, 136, 137, 200, 299, 300, 301, 1000, 3000]:
        for _ in range(10):
            a = rng.uniform(0, 1, n)
            assert vo.np_sum(a) == float(np.sum(a)), n


def test_introsort_matches_numpy_scalar():
    """the oracle's aquicksort restatement == np.argsort with the SIMD sort kernels disabled."""
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import vmoracle as vo
rng = np.random.default_rng(5)
bad = 0
for n in (2, 10, 16, 17, 18, 33, 100, 128, 300, 1000):
    for rep in range(60):
        k = rng.integers(1, max(2, n // 3))
        v = (rng.integers(0, k + 1, n) / 64.0).astype(np.float32)
        if not np.array_equal(np.argsort(v), vo.argsort_introsort_f32(v)): bad += 1
print(bad)
""" % os.path.dirname(os.path.abspath(vo.__file__))
    env = dict(os.environ, NPY_DISABLE_CPU_FEATURES="AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL "
                                                    "AVX512_SPR AVX2 FMA3")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True)
    assert out.stdout.strip().splitlines()[-1] == "0", out.stdout + out.stderr


def test_trace_exhaustion_flag():
    cfg = vo.OracleConfig(pms=3, vms=5, arrival_rate=2.0, service_length=3, training_steps=50, eval_steps=50)
    env = vo.OracleVmEnv(cfg, trace_steps=10, trace_adm=4)
    for _ in range(12):
        env.step(env.state()["vm_placement"])
    assert env.state()["trace_exhausted"] == 1


# ---- known-answer rows published by the reference (100 000-step eval episodes, 5 seeds) ---------------------
KATS = {
    # arrival_rate = round(pms/0.55/service_length, 4) (exp_performance.py:26, exp_performance_small.py:23)
    # name: (config base, arrival_rate, reward, agent, tiebreak, seeds, per-seed (return, served, place), csv row)
    "KAT1_firstfit_s10": ("10", 0.0182, "ut", vo.AGENT_FIRSTFIT, vo.TIE_STABLE, [1, 2, 3, 4, 5],
                          [(702661.695, 1260, 1275), (690616.030, 1248, 1261), (695187.965, 1250, 1262),
                           (695977.240, 1260, 1272), (700256.885, 1272, 1283)],
                          dict(ret=696939.963, drop=0.241, served=1258, cpu=0.697, var=0.051, mem=0.697, wait=0.539)),
    "KAT2_firstfit_s100": ("100", 0.1818, "wr", vo.AGENT_FIRSTFIT, vo.TIE_STABLE, [0, 1, 2, 3, 4],
                           [(-53615.794, 13360, 13494), (-53208.140, 13409, 13541), (-53318.037, 13411, 13542),
                            (-52963.827, 13466, 13601), (-53735.479, 13324, 13459)],
                           dict(ret=-53368.255, drop=0.203, served=13394, cpu=0.737, var=0.052, mem=0.736, wait=0.534)),
    "KAT3_bestfit_s10": ("10", 0.0182, "ut", vo.AGENT_BESTFIT, vo.TIE_STABLE, [1, 2, 3, 4, 5],
                         [(699859.415, None, None), (694840.720, None, None), (698802.250, None, None),
                          (697559.640, None, None), (703166.045, None, None)],
                         dict(ret=698845.614, drop=0.242, served=1260, cpu=0.699, var=0.053, mem=0.699, wait=0.537)),
    "KAT4_bestfit_s100": ("100", 0.1818, "wr", vo.AGENT_BESTFIT, vo.TIE_NUMPY_INTROSORT, [0, 1, 2, 3, 4],
                          [(-52064.658, 13794, 13935), (-51690.685, 13811, 13957), (-51570.904, 13892, 14029),
                           (-51202.317, 13950, 14090), (-51729.035, 13863, 14001)],
                          dict(ret=-51651.520, drop=0.182, served=13862, cpu=0.763, var=0.057, mem=0.762, wait=0.517)),
}
_BASE = {
    "10": dict(pms=10, vms=30, service_length=1000, training_steps=10000, eval_steps=100000, cap_target_util=True,
               sequence="uniform", beta=0.5, allow_null_action=True),          # config/10.yml:1-13
    "100": dict(pms=100, vms=300, service_length=1000, training_steps=10000, eval_steps=100000, cap_target_util=True,
                sequence="uniform", beta=0.5, allow_null_action=True),         # config/100.yml:1-13
}


@pytest.mark.slow
@pytest.mark.parametrize("kat", sorted(KATS))
def test_published_rows(kat):
    base, lam, reward, agent, tie, seeds, per_seed, row = KATS[kat]
    stats = []
    for seed, want in zip(seeds, per_seed):
        cfg = vo.OracleConfig(arrival_rate=lam, reward_function=reward, seed=seed, **_BASE[base])
        env = vo.OracleVmEnv(cfg)           # full-length traces exactly like env.reset (2*max_steps sizes)
        env.eval()
        n, st = env.rollout(agent, 10**9, tie)
        assert n == 100000
        assert round(st["return"], 3) == pytest.approx(want[0], abs=2e-3), (kat, seed, st["return"])
        if want[1] is not None:
            assert (int(st["served"]), int(st["place"])) == want[1:], (kat, seed)
        stats.append(st)
    m = {k: float(np.mean([s[k] for s in stats])) for k in stats[0]}
    assert "%.3f" % m["return"] == "%.3f" % row["ret"]
    assert "%.3f" % m["drop_rate_mean"] == "%.3f" % row["drop"]
    assert "%d" % m["served"] == "%d" % row["served"]
    assert "%.3f" % m["cpu_mean"] == "%.3f" % row["cpu"]
    assert "%.3f" % m["cpu_var"] == "%.3f" % row["var"]
    assert "%.3f" % m["mem_mean"] == "%.3f" % row["mem"]
    assert "%.3f" % m["waiting_ratio_mean"] == "%.3f" % row["wait"]


RECORD_CASES = ["rec_busy_firstfit", "rec_busy_suspend", "rec_p37_v70", "rec_s10_sparse", "rec_s100_bestfit"]


def load_record_case(name):
    import json
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "record.npz"))
    return dict(cfg=json.loads(str(z[f"{name}.cfg_json"])), actions=z[f"{name}.actions"], pending=z[f"{name}.pending"],
                slowdown=z[f"{name}.slowdown"], lifetime=z[f"{name}.lifetime"], summary=json.loads(str(z[f"{name}.summary_json"])),
                agent=str(z[f"{name}.agent"]), perturb=float(z[f"{name}.perturb"]))


@pytest.mark.parametrize("name", RECORD_CASES)
def test_oracle_record_lists_match_reference_record(name):
    """The oracle's restatement of Record (record.py:34-96) on its own episode log == the lists the reference's Record
    produced for the same action stream (tests/golden/make_golden_record.py): order and values identical."""
    import vmoracle as vo
    g = load_record_case(name)
    env = vo.OracleVmEnv(vo.OracleConfig(**g["cfg"]))
    env.eval()
    T = g["actions"].shape[0]
    env.enable_record(T)
    env.reset(seed=g["cfg"]["seed"])
    done = False
    for t in range(T):
        assert not done
        _, _, done, _, _ = env.step(g["actions"][t].astype(np.int64))
    assert done
    pending, slowdown, life = env.record_lists()
    assert np.array_equal(np.asarray(pending, np.float64), g["pending"])
    assert np.array_equal(np.asarray(slowdown, np.float64), g["slowdown"])
    assert np.array_equal(np.asarray(life, np.int64), g["lifetime"])
    s = env.record_summary()
    for k, v in s.items():
        assert float(v) == g["summary"][k], k


S100_BASE = dict(pms=100, vms=300, service_length=1000, arrival_rate=1.8182, training_steps=10000, eval_steps=100000, seed=0,
                 reward_function="kl", cap_target_util=True, sequence="uniform", beta=0.5, allow_null_action=True)   # config/100.yml


@pytest.mark.parametrize("row", ["firstfit,1.0,2100,6404,0,6533,2028,0.095,0.000,0.000",
                                 "bestfit,0.8,1000,13808,0,13943,983,0.073,0.000,0.000"])
def test_oracle_reproduces_exp_suspension_rows(row):
    """data/exp_suspension/data.csv (exp_suspension.py:12-60: reward wr, arrival = round(pms/0.55/sl*load, 3), seed 0, 100 000
    eval steps): served, valid actions, mean VM life, mean pending, mean / max slowdown — every printed digit; the best-fit
    row only on numpy's scalar introsort tie order (SURVEY §8c ruling ii)."""
    import vmoracle as vo
    agent, load, sl, *_ = row.split(",")
    c = dict(S100_BASE, reward_function="wr", service_length=int(sl),
             arrival_rate=float(np.round(100 / 0.55 / int(sl) * float(load), 3)))
    env = vo.OracleVmEnv(vo.OracleConfig(**c))
    env.eval()
    env.enable_record(100000, 400000)
    env.reset(seed=c["seed"])
    n, st = env.rollout(vo.AGENT_FIRSTFIT if agent == "firstfit" else vo.AGENT_BESTFIT, 100000,
                        vo.TIE_STABLE if agent == "firstfit" else vo.TIE_NUMPY_INTROSORT)
    assert n == 100000
    p, s, life = env.record_lists()
    got = "%s,%.1f,%d,%d,%d,%d,%d,%.3f,%.3f,%.3f" % (agent, float(load), int(sl), st["served"], st["suspend"],
                                                     st["suspend"] + st["place"], np.mean(life), np.mean(p), np.mean(s), np.max(s))
    assert got == row


def test_oracle_reproduces_exp_vm_size_row():
    """data/exp_vm_size/summary.csv:3 (first-fit, lowuniform sizes at arrival pms/0.375/sl, reward kl, seeds 0..4)."""
    import vmoracle as vo
    rows = []
    for seed in range(5):
        c = dict(S100_BASE, sequence="lowuniform", arrival_rate=100 / 0.375 / 1000, seed=seed)
        env = vo.OracleVmEnv(vo.OracleConfig(**c))
        env.eval()
        env.reset(seed=seed)
        n, st = env.rollout(vo.AGENT_FIRSTFIT, 100000, vo.TIE_STABLE)
        rows.append(st)
    m = lambda k: np.mean([r[k] for r in rows])                                          # noqa: E731
    got = "firstfit,%.4f,%.4f,%d,%d,%.4f,%.4f,%.4f,%.4f,%.4f" % (
        np.mean([np.round(r["return"], 3) for r in rows]), m("drop_rate_mean"), m("served"), m("suspend"), m("cpu_mean"),
        m("cpu_var"), m("mem_mean"), m("mem_var"), m("waiting_ratio_mean"))
    assert got == "firstfit,22539.4184,0.1232,22602,0,0.8504,0.0179,0.8497,0.0180,0.2226"


def test_oracle_record_on_the_main_py_episode():
    """`python main.py -a firstfit -e -c config/10.yml` (BASELINE configs[0]): the oracle's Record restatement on the full
    100 000-step episode == the reference Record's per-VM keys and counters (golden: make_golden_record.py, real run)."""
    import vmoracle as vo
    g = load_record_case("main_s10_firstfit_wr")
    env = vo.OracleVmEnv(vo.OracleConfig(**g["cfg"]))
    env.eval()
    env.enable_record(100000, 8192)
    env.reset(seed=g["cfg"]["seed"])
    n, st = env.rollout(vo.AGENT_FIRSTFIT, 100000, vo.TIE_STABLE)
    assert n == 100000
    pending, slowdown, life = env.record_lists()
    assert np.array_equal(np.asarray(pending, np.float64), g["pending"])
    assert np.array_equal(np.asarray(slowdown, np.float64), g["slowdown"])
    assert np.array_equal(np.asarray(life, np.int64), g["lifetime"])
    s = g["summary"]
    assert int(st["served"]) == s["total served VMs"] and int(st["total_requests"]) == s["total requests"]
    assert int(st["place"]) == s["total place actions"] and float(np.round(st["return"], 3)) == s["total rewards"]
    assert float(np.round(st["drop_rate_mean"], 3)) == s["drop rate"] and float(np.round(st["cpu_mean"], 3)) == s["cpu mean"]


@pytest.mark.parametrize("name", ["info_busy_suspend", "info_p37_v70"])
def test_oracle_arrival_log_matches_reference_vm_arrival_steps(name):
    """env.py:205,293: the oracle's arrival log rebuilt into per-slot lists == the reference env's vm_arrival_steps
    (fixture from tests/golden/make_golden_info.py)."""
    import json
    z = np.load(os.path.join(gu.GOLDEN_DIR, "info.npz"))
    fx = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + ".")}
    cfg = json.loads(str(fx["cfg_json"]))
    T, V = fx["actions"].shape
    env = vo.OracleVmEnv(vo.OracleConfig(**cfg), trace_steps=T + 8, trace_adm=4096)
    env.eval()
    env.enable_record(T + 8)
    env.reset(seed=cfg["seed"])
    for t in range(T):
        env.step(fx["actions"][t].astype(np.int64))
    import ctypes as C
    n_steps, n_arr = C.c_int64(), C.c_int64()
    vo.lib().vmo_record_counts(env._h, C.byref(n_steps), C.byref(n_arr))
    got = [[] for _ in range(V)]
    for slot, step in env._arr_log[:n_arr.value]:
        got[int(slot)].append(int(step))
    off = fx["arrival_off"]
    assert got == [fx["arrival_flat"][off[v]:off[v + 1]].tolist() for v in range(V)]
