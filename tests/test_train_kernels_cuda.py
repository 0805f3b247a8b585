"""-m gpu: the tensor-core kernels of the PPO update (csrc/vmgym_train.cu, the gradient mode of the fused head in csrc/vmgym_gemm.cu)
against plain torch fp32 / fp64 restatements on the SAME bf16-rounded operands.

Tolerances: the GEMMs accumulate bf16 products in fp32, so against a float64 product of the same rounded operands only the
summation order differs: |d| <= 2e-3 (|ref| + 1).  bf16 outputs add one rounding (2^-8 relative).  tanh in the epilogue uses the fast
exponential: 1e-3 absolute before the bf16 rounding."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lib():
    from vmgym import _native as nv
    return nv, nv.lib()


def _stream(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _gemm(torch, a, a_mn, b, b_mn, M, N, K, bias=None, act=0, mul_y=None, c32=None, accumulate=False, want_bf16=False, want_rows=False,
          rows_init=None):
    nv, lib = _lib()
    out32 = c32 if c32 is not None else torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    out16 = torch.full((M, N + 8), float("nan"), dtype=torch.bfloat16, device="cuda") if want_bf16 else None      # ld != N on purpose
    rows = (rows_init.clone() if rows_init is not None else torch.full((M,), float("nan"), dtype=torch.float32, device="cuda")) if want_rows else None
    nv.check(lib.vmgym_tc_gemm(a.data_ptr(), int(a_mn), a.stride(0), b.data_ptr(), int(b_mn), b.stride(0), M, N, K,
                               bias.data_ptr() if bias is not None else None, act, mul_y.data_ptr() if mul_y is not None else None,
                               mul_y.stride(0) if mul_y is not None else 0, out32.data_ptr(), out32.stride(0), int(accumulate),
                               out16.data_ptr() if want_bf16 else None, out16.stride(0) if want_bf16 else 0,
                               rows.data_ptr() if want_rows else None, _stream(torch)), "vmgym_tc_gemm")
    torch.cuda.synchronize()
    return out32, (out16[:, :N] if want_bf16 else None), rows


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 520, 200), (77, 1100, 1000), (512, 512, 4096)])
def test_tc_gemm_all_operand_layouts(a_mn, b_mn, M, N, K):
    import torch
    torch.manual_seed(M + 7 * N + 13 * K + a_mn + 2 * b_mn)
    pad = lambda n: (n + 7) // 8 * 8          # noqa: E731  operand row strides are multiples of 8 elements
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    a = torch.zeros((K, pad(M)) if a_mn else (M, pad(K)), dtype=torch.bfloat16, device="cuda")
    b = torch.zeros((K, pad(N)) if b_mn else (N, pad(K)), dtype=torch.bfloat16, device="cuda")
    if a_mn:
        a[:, :M] = A.t()
    else:
        a[:, :K] = A
    if b_mn:
        b[:, :N] = B.t()
    else:
        b[:, :K] = B
    Ar = (a[:, :M].t() if a_mn else a[:, :K]).double()
    Br = (b[:, :N].t() if b_mn else b[:, :K]).double()
    ref = Ar @ Br.t()
    got, got16, rows = _gemm(torch, a, a_mn, b, b_mn, M, N, K, want_bf16=True, want_rows=True)
    tol = 2e-3 * (ref.abs() + 1.0) * max(1.0, (K / 256) ** 0.5)
    assert bool(((got.double() - ref).abs() <= tol).all()), f"max |d| {(got.double() - ref).abs().max().item()}"
    assert bool(((got16.double() - ref).abs() <= tol + ref.abs() * 2 ** -7).all())
    rs = Ar.sum(1)
    assert bool(((rows.double() - rs).abs() <= 2e-3 * (rs.abs() + 1.0) * max(1.0, (K / 256) ** 0.5)).all()), "row sums of A"


def test_tc_gemm_epilogues():
    import torch
    torch.manual_seed(5)
    M, N, K = 200, 512, 520
    x = (torch.randn(M, K, device="cuda") * 0.2).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.2).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    z = x.double() @ w.double().t() + bias.double()
    # forward layer: tanh(x W^T + b) -> bf16 (ppo.py:91-109)
    _, a16, _ = _gemm(torch, x, 0, w, 0, M, N, K, bias=bias, act=1, want_bf16=True)
    assert bool(((a16.double() - torch.tanh(z)).abs() <= 1e-3 + 2 ** -7).all())
    # input gradient: (dz W) * (1 - a^2), W seen MN-major
    dz = (torch.randn(M, N, device="cuda") * 0.1).to(torch.bfloat16)
    a_prev = torch.tanh(torch.randn(M, K, device="cuda")).to(torch.bfloat16)
    ref = (dz.double() @ w.double()) * (1.0 - a_prev.double() ** 2)
    got, got16, _ = _gemm(torch, dz, 0, w, 1, M, K, N, mul_y=a_prev, want_bf16=True)
    assert bool(((got.double() - ref).abs() <= 2e-3 * (ref.abs() + 1.0)).all())
    assert bool(((got16.double() - ref).abs() <= 2e-3 * (ref.abs() + 1.0) + ref.abs() * 2 ** -7).all())
    # weight gradient accumulated over two sample chunks + bias gradient: dW += dz^T a_prev, db += sum dz
    dW = torch.zeros((N, K), dtype=torch.float32, device="cuda")
    db = torch.zeros(N, dtype=torch.float32, device="cuda")
    nv, lib = _lib()
    for lo, hi in ((0, 128), (128, M)):
        nv.check(lib.vmgym_tc_gemm(dz[lo:hi].data_ptr(), 1, dz.stride(0), a_prev[lo:hi].data_ptr(), 1, a_prev.stride(0), N, K, hi - lo, None, 0,
                                   None, 0, dW.data_ptr(), dW.stride(0), 1, None, 0, db.data_ptr(), _stream(torch)), "vmgym_tc_gemm")
    torch.cuda.synchronize()
    refW = dz.double().t() @ a_prev.double()
    assert bool(((dW.double() - refW).abs() <= 2e-3 * (refW.abs() + 1.0)).all())
    assert bool(((db.double() - dz.double().sum(0)).abs() <= 2e-3 * (dz.double().sum(0).abs() + 1.0)).all())


def test_tc_gemm_split_k_weight_gradient():
    """dW += dz^T a with few output tiles and a long sample axis takes the split-K path (atomic partial tiles): same result."""
    import torch
    torch.manual_seed(3)
    M, H, D = 9000, 512, 1100
    dz = (torch.randn(M, H, device="cuda") * 0.1).to(torch.bfloat16)
    x = torch.zeros((M, 3 * 1104), dtype=torch.bfloat16, device="cuda")
    x[:, :D] = torch.randn(M, D, device="cuda")
    dW = torch.full((H, D), 0.5, dtype=torch.float32, device="cuda")
    db = torch.full((H,), -1.0, dtype=torch.float32, device="cuda")
    nv, lib = _lib()
    nv.check(lib.vmgym_tc_gemm(dz.data_ptr(), 1, dz.stride(0), x.data_ptr(), 1, x.stride(0), H, D, M, None, 0, None, 0, dW.data_ptr(), dW.stride(0), 1,
                               None, 0, db.data_ptr(), _stream(torch)), "vmgym_tc_gemm")
    torch.cuda.synchronize()
    ref = 0.5 + dz.double().t() @ x[:, :D].double()
    assert bool(((dW.double() - ref).abs() <= 3e-3 * (ref.abs() + 1.0)).all()), f"max |d| {(dW.double() - ref).abs().max().item()}"
    rs = -1.0 + dz.double().sum(0)
    assert bool(((db.double() - rs).abs() <= 3e-3 * (rs.abs() + 1.0)).all())


def test_small_training_kernels():
    import torch
    nv, lib = _lib()
    torch.manual_seed(9)
    st = _stream(torch)
    # cast + pad
    src = torch.randn(37, 1100, device="cuda")
    dst = torch.full((37, 1104), float("nan"), dtype=torch.bfloat16, device="cuda")
    nv.check(lib.vmgym_cast_pad_bf16(src.data_ptr(), 37, 1100, src.stride(0), dst.data_ptr(), 1104, st), "cast")
    assert torch.equal(dst[:, :1100], src.to(torch.bfloat16)) and bool((dst[:, 1100:] == 0).all())
    # value head forward / backward
    H, M = 512, 333
    h = torch.tanh(torch.randn(M, H, device="cuda")).to(torch.bfloat16)
    w, b = torch.randn(H, device="cuda") * 0.1, torch.randn(1, device="cuda")
    out = torch.empty(M, device="cuda")
    nv.check(lib.vmgym_value_head(h.data_ptr(), M, H, w.data_ptr(), b.data_ptr(), out.data_ptr(), st), "value_head")
    ref = h.double() @ w.double() + b.double()
    assert torch.allclose(out.double(), ref, rtol=1e-5, atol=1e-4)
    dv = torch.randn(M, device="cuda")
    dz = torch.empty((M, H), dtype=torch.bfloat16, device="cuda")
    dw, db = torch.zeros(H, device="cuda"), torch.zeros(1, device="cuda")
    nv.check(lib.vmgym_value_head_backward(h.data_ptr(), M, H, w.data_ptr(), dv.data_ptr(), dz.data_ptr(), dw.data_ptr(), db.data_ptr(), st), "vhb")
    rz = dv.double()[:, None] * w.double()[None, :] * (1 - h.double() ** 2)
    assert bool(((dz.double() - rz).abs() <= rz.abs() * 2 ** -7 + 1e-6).all())
    assert torch.allclose(dw.double(), (dv.double()[:, None] * h.double()).sum(0), rtol=1e-4, atol=1e-3)
    assert torch.allclose(db.double(), dv.double().sum().reshape(1), rtol=1e-5, atol=1e-4)
    # PPO loss coefficients vs autograd of the reference formulas (ppo.py:259-282)
    n = 4097
    for vf_clip in (1, 0):
        new_lp = (torch.randn(n, device="cuda") * 0.15).requires_grad_(True)
        old_lp = torch.randn(n, device="cuda") * 0.05
        adv, ent = torch.randn(n, device="cuda"), torch.rand(n, device="cuda") * 5
        v = (torch.randn(n, device="cuda") * 0.3).requires_grad_(True)
        v_old, ret = v.detach() + torch.randn(n, device="cuda") * 0.12, torch.randn(n, device="cuda")
        eps, ec, vc = 0.1, 0.01, 0.5
        ratios = torch.exp(new_lp - old_lp)
        l_pg = torch.max(-ratios * adv, -torch.clamp(ratios, 1 - eps, 1 + eps) * adv)
        l_un = torch.square(v - ret)
        l_cl = torch.square(v_old + torch.clamp(v - v_old, -eps, eps) - ret)
        l_vf = 0.5 * (torch.max(l_un, l_cl) if vf_clip else l_un)
        loss = (l_pg - ec * ent + vc * l_vf).sum() / n
        loss.backward()
        c_lp, c_v = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
        sums = torch.zeros(2, dtype=torch.float64, device="cuda")
        nv.check(lib.vmgym_ppo_loss(new_lp.data_ptr(), old_lp.data_ptr(), adv.data_ptr(), ent.data_ptr(), v.data_ptr(), v_old.data_ptr(), ret.data_ptr(),
                                    n, eps, ec, vc, vf_clip, 1.0 / n, c_lp.data_ptr(), c_v.data_ptr(), sums.data_ptr(), st), "ppo_loss")
        torch.cuda.synchronize()
        assert torch.allclose(c_lp, new_lp.grad, rtol=1e-4, atol=1e-9) and torch.allclose(c_v, v.grad, rtol=1e-4, atol=1e-9)
        assert float(sums[1]) == pytest.approx(float(loss), rel=1e-5)
        assert float(sums[0]) == pytest.approx(float((new_lp - old_lp).double().sum()), rel=1e-6, abs=1e-6)


@pytest.mark.parametrize("V,A,M", [(30, 12, 70), (300, 102, 200)])
def test_fused_head_gradient_mode_matches_autograd(V, A, M):
    """vmgym_policy_fused_grad: d/dlogits of sum_e c_lp[e] logprob(e) + c_ent entropy(e), logits recomputed in tensor memory, vs
    torch autograd of the reference's masked multi-categorical (ppo.py:115-126) on logits = bf16(h) bf16(W)^T + b."""
    import torch
    import torch.nn as nn
    from vmgym.ppo import FusedActorHead
    nv, lib = _lib()
    torch.manual_seed(V + A)
    K = 512
    lin = nn.Linear(K, V * A).cuda()
    with torch.no_grad():
        lin.weight.mul_(3.0)
    head = FusedActorHead(lin, V, A)
    h = torch.tanh(torch.randn(M, K, device="cuda")).to(torch.bfloat16)
    mask = torch.rand(M, V, A, device="cuda") < 0.6
    mask[:, :, A - 1] = False                                            # at least one valid column per row
    act = torch.where(mask, torch.full((M, V, A), -1e9, device="cuda"), torch.rand(M, V, A, device="cuda")).argmax(-1).to(torch.uint8)
    words = torch.zeros((M, V, 4), dtype=torch.int32, device="cuda")
    bits = mask.reshape(M, V, A).cpu().numpy()
    pk = np.zeros((M, V, 128), np.uint64)
    pk[:, :, :A] = bits
    words.copy_(torch.from_numpy((pk.reshape(M, V, 4, 32) << np.arange(32, dtype=np.uint64)).sum(-1).astype(np.uint32).view(np.int32)))
    c_lp = torch.randn(M, device="cuda")
    c_ent = -0.37
    R = head.TILE                                                        # columns per VM in the gradient tensor: A rounded up to 16
    assert R == (A + 15) // 16 * 16 == lib.vmgym_policy_fused_rows(A, K)
    g = torch.full((M, V * R + 8), float("nan"), dtype=torch.bfloat16, device="cuda")
    lp_f, en_f = torch.empty((M, V), device="cuda"), torch.empty((M, V), device="cuda")
    st_m, st_s = torch.empty((M, V), device="cuda"), torch.empty((M, V), device="cuda")
    nv.check(lib.vmgym_policy_fused_eval(h.data_ptr(), head.w_pad.data_ptr(), head.b_pad.data_ptr(), words.data_ptr(), act.data_ptr(), M, V, A, K,
                                         lp_f.data_ptr(), en_f.data_ptr(), st_m.data_ptr(), st_s.data_ptr(), _stream(torch)), "fused_eval")
    nv.check(lib.vmgym_policy_fused_grad(h.data_ptr(), head.w_pad.data_ptr(), head.b_pad.data_ptr(), words.data_ptr(), act.data_ptr(), M, V, A, K,
                                         c_lp.data_ptr(), c_ent, en_f.data_ptr(), st_m.data_ptr(), st_s.data_ptr(), g.data_ptr(), g.stride(0),
                                         _stream(torch)), "fused_grad")
    torch.cuda.synchronize()
    logits = (h.double() @ lin.weight.detach().to(torch.bfloat16).double().t() + lin.bias.detach().double()).requires_grad_(True)
    z = logits.reshape(M, V, A).masked_fill(mask, -1e7)
    logp = torch.log_softmax(z, -1)
    lp = logp.gather(-1, act.long().unsqueeze(-1)).squeeze(-1).sum(1)
    ent = -(logp.exp() * logp).sum(-1).sum(1)
    (lp * c_lp.double()).sum().add(c_ent * ent.sum()).backward()
    # the evaluating forward (valid-column epilogue): per-(env, VM) log-prob and entropy
    rlp = logp.gather(-1, act.long().unsqueeze(-1)).squeeze(-1).detach()
    rent = -(logp.exp() * logp).sum(-1).detach()
    assert torch.allclose(lp_f.double(), rlp, rtol=1e-4, atol=2e-4) and torch.allclose(en_f.double(), rent, rtol=1e-4, atol=2e-4)
    ref = logits.grad.reshape(M, V, A)
    got = g[:, :V * R].reshape(M, V, R).double()
    assert bool((got[:, :, A:] == 0).all()), "padding columns must be zero"
    assert bool((got[:, :, :A][mask] == 0).all()), "masked columns must get no gradient"
    err = (got[:, :, :A] - ref).abs()
    assert bool((err <= 2e-3 + ref.abs() * 2 ** -6).all()), f"max |d| {err.max().item()}"
