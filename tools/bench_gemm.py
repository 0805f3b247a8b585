#!/usr/bin/env python
"""Time the hand-written tcgen05 actor-output GEMM against cuBLAS (torch) on the PPO shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import torch
from vmgym.ppo import linear_bf16

def t(fn, it=20):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it

for M in (512, 4096, 8192):
    N, K = 30600, 512
    a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16(); b = torch.randn(N, device="cuda")
    out32 = torch.empty(M, N, device="cuda")
    ms_mine = t(lambda: linear_bf16(a, w, b))
    ms_cublas_bf16 = t(lambda: torch.addmm(b.bfloat16(), a, w.T))
    a32, w32 = a.float(), w.float()
    ms_cublas_f32out = t(lambda: torch.addmm(b, a32, w32.T, out=out32))
    fl = 2.0 * M * N * K
    print(f"M={M}: tcgen05 kernel {ms_mine:.3f} ms = {fl/ms_mine/1e9:.0f} TFLOP/s ({M*N*4/ms_mine/1e6:.0f} GB/s of fp32 output) | "
          f"cuBLAS bf16->bf16 {ms_cublas_bf16:.3f} ms = {fl/ms_cublas_bf16/1e9:.0f} TF | cuBLAS fp32(TF32?)->fp32 {ms_cublas_f32out:.3f} ms", flush=True)
