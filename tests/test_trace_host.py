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
