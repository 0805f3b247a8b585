"""vmgym_tc_gemm at the DRL-VMP head shapes (small M, long K): kernel time by CUDA events, back to back."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vm-placement-migration-gym_b200")]
import torch
from vmgym import _native as nv
lib = nv.lib(); st = C.c_void_p(torch.cuda.current_stream().cuda_stream); bf = torch.bfloat16
def run(M, N, K, act, split_out, f32_out, label):
    a = torch.randn(M, K, device="cuda").to(bf); b = torch.randn(N, K, device="cuda").to(bf); bias = torch.randn(N, device="cuda")
    c16 = torch.empty(M, 3 * N if split_out else N, device="cuda", dtype=bf) if (split_out or not f32_out) else None
    c32 = torch.empty(M, N, device="cuda") if f32_out else None
    def f():
        nv.check(lib.vmgym_tc_gemm(a.data_ptr(), 0, K, b.data_ptr(), 0, K, M, N, K, bias.data_ptr(), act | (8 if split_out else 0), None, 0,
                                   c32.data_ptr() if c32 is not None else None, N, 0, c16.data_ptr() if c16 is not None else None,
                                   c16.stride(0) if c16 is not None else 0, None, st), "g")
    f(); f(); torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    global st
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(50): f()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    ref = (a.float() @ b.float().t())
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.backends.cuda.matmul.allow_tf32 = True
    af, bfl = a.float(), b.float()
    torch.addmm(bias, af, bfl.t()); torch.cuda.synchronize()
    t0.record()
    for _ in range(50): torch.addmm(bias, af, bfl.t())
    t1.record(); torch.cuda.synchronize()
    print(f"{label:40s} M={M} N={N} K={K}: tc_gemm {us:7.1f} us = {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s | cuBLAS TF32 (K/3 x3 not applied) {t0.elapsed_time(t1) / 50 * 1e3:6.1f} us")
run(1024, 1024, 1536, 2, True, False, "hidden heads (relu, split out)")
run(1024, 320, 3072, 0, False, True, "output heads (fp32 out)")
run(1024, 1024, 512, 2, False, False, "hidden heads, plain bf16")
run(1024, 320, 1024, 0, False, True, "output heads, plain bf16")
run(1024, 128, 64, 0, False, True, "floor: one k-block, 8 CTAs")
run(1024, 1024, 64, 0, False, True, "floor: one k-block, 64 CTAs")
run(4096, 512, 3312, 1, False, False, "PPO layer 1 @4096")
run(32768, 512, 3312, 1, False, False, "PPO layer 1 @32768")
