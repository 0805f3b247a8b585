"""-m gpu, needs >= 2 GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`):
the N-GPU == 1-GPU parity tests of SURVEY §4 tier 3 under torch.distributed.run with the NCCL backend — see
tests/multi_gpu_worker.py.  The run's output at 2 GPUs is kept in profiles/r2_multi_gpu_parity.md."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2])
def test_sharded_envs_and_ppo_gradients_equal_single_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "multi_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(res.stdout[-4000:])
    sys.stderr.write(res.stderr[-4000:])
    assert res.returncode == 0, "multi-GPU parity worker failed"
    assert "MULTI_GPU_OK" in res.stdout
