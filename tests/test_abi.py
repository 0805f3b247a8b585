"""CPU suite: the C-ABI library builds (nvcc cross-compiles without a GPU), loads, and exports every symbol
include/vmgym.h declares; layout queries (host-only) behave; no compute is launched here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vmgym.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vmgym_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from vmgym import _native as nv
    lib = nv.lib()
    names = _declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vmgym.h but not exported"
    assert sorted(nv.EXPORTS) == names
    assert lib.vmgym_abi_version() == 9


@pytest.mark.parametrize("P,V,place_bytes", [(10, 30, 1), (100, 300, 1), (253, 64, 1), (254, 64, 2), (1000, 3000, 2), (3, 5, 1)])
def test_layout(P, V, place_bytes):
    from vmgym import _native as nv
    lib = nv.lib()
    lay = nv.Layout()
    cfg = nv.Config(P, V, 1, 1, 1, 100, 0.5)
    assert lib.vmgym_get_layout(C.byref(cfg), C.byref(lay)) == 0
    assert lay.place_bytes == place_bytes and lay.obs_dim == 3 * V + 2 * P and lay.action_dim == P + 2
    assert lay.record_bytes % 128 == 0
    offs = [lay.off_cpu, lay.off_memory, lay.off_remaining, lay.off_placement, lay.off_cpu_code, lay.off_mem_code, lay.off_scalars]
    assert offs == sorted(offs) and all(o % 16 == 0 for o in offs)
    assert lay.off_memory - lay.off_cpu >= 8 * P and lay.off_placement - lay.off_remaining >= 2 * V
    assert lay.off_scalars + nv.SCALARS_BYTES <= lay.record_bytes
    assert lay.smem_bytes_per_env >= lay.record_bytes


def test_errors_do_not_throw():
    from vmgym import _native as nv
    lib = nv.lib()
    lay = nv.Layout()
    bad = nv.Config(0, 30, 1, 1, 1, 100, 0.5)
    assert lib.vmgym_get_layout(C.byref(bad), C.byref(lay)) == nv.EINVAL
    assert b"pms" in lib.vmgym_last_error()
    bad = nv.Config(10, 30, 1, 9, 1, 100, 0.5)          # unknown reward (the reference asserts, env.py:155-156)
    assert lib.vmgym_get_layout(C.byref(bad), C.byref(lay)) == nv.EINVAL
    assert lib.vmgym_set_tuning(99, 1) == nv.EINVAL
    assert lib.vmgym_set_tuning(0, 7) == 0


def test_env_requires_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from vmgym import Config, VecVmEnv
    from vmgym._native import VmgymError
    with pytest.raises(VmgymError):
        VecVmEnv(Config(), 4)


def test_code_to_f64_table():
    """vmgym_device.cuh code_to_f64: q0 = k*0.01; r = fma(-q0, 100, k); q = fma(r, 0.01, q0) equals the correctly rounded
    k / 100.0 (== np.around(u, 2), env.py:212-219) for every size code; fma restated with exact rationals."""
    from fractions import Fraction as F

    def fma(a, b, c):
        return float(F(a) * F(b) + F(c))
    for k in range(256):
        q0 = k * 0.01
        q = fma(fma(-q0, 100.0, float(k)), 0.01, q0)
        assert q == k / 100.0, k


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-steps", "200"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "scaling", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "env-steps/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in line["config"]
