"""CPU suite: host-side trace logic of the product (vmgym/trace.py) — the numpy draws equal the oracle harness's
(both follow env.py:172-178,211-219,272,289), packing round-trips, Philox CDF tables are exact inversions."""
import math

import numpy as np
import pytest

import vmoracle as vo
from vmgym.config import Config
from vmgym.trace import EnvStreams, pack_trace, poisson_cdf_table, sample_numpy_traces


@pytest.mark.parametrize("seq", ["uniform", "lowuniform", "highuniform"])
def test_numpy_streams_match_oracle_harness(seq):
    cfg = Config(pms=10, vms=30, arrival_rate=0.7, service_length=40, training_steps=300, eval_steps=500, sequence=seq)
    ocfg = vo.OracleConfig(pms=10, vms=30, arrival_rate=0.7, service_length=40, training_steps=300, eval_steps=500, sequence=seq)
    arr, adm = sample_numpy_traces(cfg, [EnvStreams(5), EnvStreams(9)], 200, 150)
    for i, seed in enumerate((5, 9)):
        tr = vo.sample_trace(ocfg, seed, 200, 150)
        assert np.array_equal(arr[i], tr.arrivals)
        assert np.array_equal(adm[i] & 0xff, np.rint(tr.cpu_seq * 100).astype(np.uint32))
        assert np.array_equal((adm[i] >> 8) & 0xff, np.rint(tr.mem_seq * 100).astype(np.uint32))
        assert np.array_equal(adm[i] >> 16, tr.svc_seq.astype(np.uint32))


def test_stream_continuation_matches_reference_semantics():
    """reset() without a seed continues rng3/rng4 where the episode stopped (env.py:180-184)."""
    cfg = Config(pms=5, vms=9, arrival_rate=1.5, service_length=7, training_steps=50, eval_steps=50)
    s = EnvStreams(3)
    a1, c1, m1, v1 = s.draw(cfg, 50, None)
    s.rewind_to(cfg, 20, 11)          # the episode consumed 20 arrival draws and 11 admissions
    a2, c2, m2, v2 = s.draw(cfg, 50, None)
    ref = [np.random.default_rng(3 + i) for i in range(4)]
    ref[0].uniform(0.1, 1, 100); ref[1].uniform(0.1, 1, 100)
    ref[2].poisson(1.5, 20); ref[3].poisson(7, 11)
    assert np.array_equal(a2, ref[2].poisson(1.5, 50))
    assert np.array_equal(c2, np.around(ref[0].uniform(0.1, 1, 100), 2)[: len(c2)])
    assert np.array_equal(v2, ref[3].poisson(7, len(v2)) + 1)


def test_pack_limits():
    a = np.array([1, 2]); cpu = np.array([0.1, 1.0]); mem = np.array([0.55, 0.25]); svc = np.array([70000, 3])
    with pytest.raises(ValueError):
        pack_trace([(a, cpu, mem, svc)], 2)


@pytest.mark.parametrize("lam", [0.0182, 1.8182, 12.0, 1000.0])
def test_poisson_cdf_table(lam):
    kmin, th = poisson_cdf_table(lam)
    assert th.dtype == np.uint64 and th[-1] == np.iinfo(np.uint64).max
    assert np.all(np.diff(th.astype(np.float64)) >= 0)
    # pmf recovered from the thresholds matches the Poisson pmf
    p = np.diff(np.concatenate([[0.0], th.astype(np.float64) / 2.0 ** 64]))
    ks = kmin + np.arange(len(th))
    want = np.exp(ks * math.log(lam) - lam - np.array([math.lgamma(k + 1.0) for k in ks]))
    assert np.allclose(p[:-1], want[:-1], atol=1e-12)
    assert abs((p * ks).sum() - lam) < 1e-6 * max(1.0, lam)


def test_histogram_statistics_equal_numpy_on_the_lists():
    """vec_env._hist_stats / vm_stats_from_histograms: mean, median and max of the rate lists reconstructed from the
    1001-bin histograms the kernel keeps == np.mean / np.median / np.max of the lists themselves (record.py:118-125)."""
    import importlib.util
    import os
    import sys
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vm-placement-migration-gym_b200")
    sys.path.insert(0, root)
    from vmgym.vec_env import _hist_stats, vm_stats_from_histograms
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 10, 11, 500, 501):
        for skew in (1.0, 4.0):
            rates = np.around(rng.random(n) ** skew, 3)
            h = np.bincount(np.rint(rates * 1000).astype(np.int64), minlength=1024)
            mean, median, mx = _hist_stats(h)
            assert abs(mean - np.mean(rates)) < 1e-12 and median == np.median(rates) and mx == np.max(rates), (n, skew)
    assert _hist_stats(np.zeros(1024, np.int64)) == (0.0, 0.0, 0.0)           # record.py:83-84: empty slowdown list -> [0]
    hist = np.zeros((2, 2, 1024), np.int64)
    hist[0, 0, [100, 300, 1000]] = [2, 1, 1]; hist[0, 1, [0, 250]] = [2, 1]
    totals = np.array([[4, 3, 90, 0], [0, 0, 0, 0]], np.int64)
    s = vm_stats_from_histograms(hist, totals)
    assert s["average pending"][0] == np.mean([0.1, 0.1, 0.3, 1.0]) and s["median pending"][0] == 0.2 and s["max pending"][0] == 1.0
    assert s["average VM life"][0] == 22.5 and s["average slowdown"][0] == np.mean([0.0, 0.0, 0.25]) and np.isnan(s["average VM life"][1])


def test_suspension_grid_matches_reference_script():
    """vmgym.sweep.suspension_points: the (load, service length, arrival rate) grid of exp_suspension.py:75-85,19."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vm-placement-migration-gym_b200"))
    from vmgym.sweep import suspension_points
    pts = suspension_points(100)
    assert len(pts) == 20 + 9
    assert pts[0] == dict(service_length=100, arrival_rate=float(np.round(100 / 0.55 / 100, 3)), _load=1.0)
    assert pts[19]["service_length"] == 3900 and pts[20]["service_length"] == 1000 and abs(pts[20]["_load"] - 0.2) < 1e-12
    assert pts[-1]["arrival_rate"] == float(np.round(100 / 0.55 / 1000 * np.arange(0.2, 1.1, 0.1)[-1], 3))
