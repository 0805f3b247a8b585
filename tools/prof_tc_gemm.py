"""ncu / timing target: the three big GEMM shapes of the PPO update on vmgym_tc_gemm and the fused head (forward + gradient mode).
    python tools/prof_tc_gemm.py [samples]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import torch  # noqa: E402

from vmgym import _native as nv  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
V, A, H = 300, 102, 512
T = nv.lib().vmgym_policy_fused_rows(A, H)            # rows per VM of the padded output layer
lib = nv.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
bf = torch.bfloat16
g = (torch.randn(M, V * T, device="cuda") * 0.01).to(bf)
a2 = torch.tanh(torch.randn(M, H, device="cuda")).to(bf)
w = (torch.randn(V * T, H, device="cuda") * 0.05).to(bf)
bias = torch.zeros(V * T, device="cuda")
dW = torch.zeros(V * T, H, device="cuda")
db = torch.zeros(V * T, device="cuda")
dz = torch.empty(M, H, device="cuda", dtype=bf)
lp, en = torch.empty(M, V, device="cuda"), torch.empty(M, V, device="cuda")
sm_, ss_ = torch.empty(M, V, device="cuda"), torch.empty(M, V, device="cuda")
# masks as in a saturated config/100.yml env: ~53 % of the slots run (valid: own PM + WAIT), the rest wait with nothing fitting (WAIT only)
import numpy as np  # noqa: E402
rng = np.random.default_rng(0)
valid = np.zeros((M, V, 128), bool)
valid[:, :, 100] = True
run = rng.random((M, V)) < 0.53
pm = rng.integers(0, 100, (M, V))
ii, jj = np.nonzero(run)
valid[ii, jj, pm[ii, jj]] = True
if len(sys.argv) > 2 and sys.argv[2] == "dense":       # far from saturation: a waiting VM fits on most PMs
    valid[:, :, :100] |= (rng.random((M, V, 100)) < 0.6) & ~run[:, :, None]
inv = ~valid
inv[:, :, A:] = False
words = (inv.reshape(M, V, 4, 32).astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(-1).astype(np.uint32).view(np.int32)
mask = torch.from_numpy(words).cuda()
act = torch.full((M, V), 100, dtype=torch.uint8, device="cuda")
c_lp = torch.randn(M, device="cuda")
act2 = torch.empty((M, V), dtype=torch.uint8, device="cuda")


def gemm(a, a_mn, b, b_mn, Mx, Nx, Kx, c32=None, acc=0, c16=None, mul=None, rows=None):
    nv.check(lib.vmgym_tc_gemm(a.data_ptr(), a_mn, a.stride(0), b.data_ptr(), b_mn, b.stride(0), Mx, Nx, Kx, None, 0,
                               mul.data_ptr() if mul is not None else None, mul.stride(0) if mul is not None else 0,
                               c32.data_ptr() if c32 is not None else None, c32.stride(0) if c32 is not None else 0, acc,
                               c16.data_ptr() if c16 is not None else None, c16.stride(0) if c16 is not None else 0,
                               rows.data_ptr() if rows is not None else None, st), "gemm")


cases = {
    "dW3 = g^T a2 (+ row sums)": (lambda: gemm(g, 1, a2, 1, V * T, H, M, c32=dW, acc=1, rows=db), 2.0 * V * T * H * M),
    "dz2 = (g W3)(1 - a2^2)": (lambda: gemm(g, 0, w, 1, M, H, V * T, c16=dz, mul=a2), 2.0 * V * T * H * M),
    "fused head eval forward": (lambda: nv.check(lib.vmgym_policy_fused_eval(a2.data_ptr(), w.data_ptr(), bias.data_ptr(), mask.data_ptr(), act.data_ptr(), M, V, A, H,
                                                                             lp.data_ptr(), en.data_ptr(), sm_.data_ptr(), ss_.data_ptr(), st), "f"), 2.0 * V * T * H * M),
    "fused head gradient": (lambda: nv.check(lib.vmgym_policy_fused_grad(a2.data_ptr(), w.data_ptr(), bias.data_ptr(), mask.data_ptr(), act.data_ptr(), M, V, A, H,
                                                                        c_lp.data_ptr(), -1e-7, en.data_ptr(), sm_.data_ptr(), ss_.data_ptr(), g.data_ptr(), g.stride(0), st), "fg"),
                            2.0 * V * T * H * M),
    "fused head sampling (rollout)": (lambda: nv.check(lib.vmgym_policy_fused(a2.data_ptr(), w.data_ptr(), bias.data_ptr(), mask.data_ptr(), None, M, V, A, H,
                                                                              1, 1, act2.data_ptr(), lp.data_ptr(), en.data_ptr(), st), "fs"), 2.0 * V * T * H * M),
}
for name, (fn, flops) in cases.items():
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:32s} {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s (padded tiles)")
