"""Fused legacy rel-pos attention (csrc/attn_fused.cu: tcgen05 QK^T / PV with the scores on chip) against the CPU
oracle's restatement of transformer/attention.py:145-209 (matmul, rel_shift, finfo.min key fill, softmax, zero
fill, dropout, matmul) and its autograd gradients, at the BASELINE shapes (S = 1152 and 1692, dk = 192), with
ragged / fully padded utterances and the bit-identical dropout hash.  bf16 operands: tolerances are relative to the
tensor's scale and written at the comparison."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from a3t_b200 import _lib
from oracle import a3t_oracle as O


@pytest.fixture(scope="module")
def tc(cuda_lib):
    from a3t_b200.backend import CudaBackend

    return CudaBackend("cuda:0", torch.bfloat16, seed=24680, impl=_lib.IMPL_TC)


def gb(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).to(torch.bfloat16).float()


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-6)


def oracle_attention(qkv4, p, keymask, H, scale, drop, seed):
    ac, bd = O.attn_scores_fwd(qkv4, p, H)
    d = None if drop is None else (drop[0], seed, drop[1])
    P, Pd = O.relpos_softmax_fwd(ac, bd, keymask, scale, drop=d)
    return O.attn_pv_fwd(Pd, qkv4, H)


def _case(B, H, S, dk, lens, seed=1, grow=0.0):
    D = H * dk
    qkv4, p = gb(B, S, 4 * D, seed=seed), gb(S, D, seed=seed + 1)
    if grow:  # later keys score higher: the running maximum keeps moving (exercises the accumulator rescale)
        ramp = 1.0 + grow * torch.arange(S).float() / S
        qkv4[..., 2 * D:3 * D] = (qkv4[..., 2 * D:3 * D] * ramp[None, :, None]).to(torch.bfloat16).float()
    keymask = torch.arange(S)[None, :] < torch.tensor(lens)[:, None]
    return qkv4, p, keymask


CASES = [
    (2, 2, 1152, 192, [1152, 1000], 0.0),     # cfg2 shape, one ragged utterance
    (1, 2, 1692, 192, [1692], 0.0),           # cfg4 sequence length (S % 128 != 0, S % 8 != 0)
    (3, 2, 200, 64, [200, 77, 0], 0.0),       # short, a FULLY padded utterance (no valid key at all), dk = 64
    (2, 1, 300, 128, [300, 129], 6.0),        # dk = 128, growing scores -> lazy rescale path
]


@pytest.mark.parametrize("B,H,S,dk,lens,grow", CASES)
@pytest.mark.parametrize("drop", [None, (0.2, 5)])
def test_fused_attention_forward_vs_oracle(tc, B, H, S, dk, lens, grow, drop):
    qkv4, p, keymask = _case(B, H, S, dk, lens, grow=grow)
    assert tc.attn_fused_ok(B, H, S, H * dk)
    scale = 1.0 / math.sqrt(dk)
    ref = oracle_attention(qkv4, p, keymask, H, scale, drop, 24680)
    ctx, bd, lse = tc.attn_fwd_fused(qkv4.cuda().to(torch.bfloat16), p.cuda().to(torch.bfloat16), keymask.cuda(), H, scale,
                                     drop=drop)
    assert ctx.dtype == torch.bfloat16 and ctx.shape == (B, S, H * dk)
    assert torch.isfinite(ctx.float()).all()
    # bf16 P and bf16 BD_raw against the fp32 reference: 1.5e-2 of the output scale
    assert rel_err(ctx, ref) < 1.5e-2, rel_err(ctx, ref)
    for b, n in enumerate(lens):   # an utterance without a valid key attends to nothing: exact zeros (attention.py:79-86)
        if n == 0:
            assert float(ctx[b].float().abs().max()) == 0.0
    # lse is the log2-domain log-sum-exp of the scaled, masked scores
    ac, bdr = O.attn_scores_fwd(qkv4, p, H)
    sc = (ac + O.rel_shift(bdr)) * scale
    sc = sc.masked_fill(~keymask.view(B, 1, 1, S), float("-inf"))
    want = torch.logsumexp(sc, -1) / math.log(2.0)
    ok = torch.isfinite(want)
    assert float((lse.cpu()[ok] - want[ok]).abs().max()) < 3e-2


@pytest.mark.parametrize("B,H,S,dk,lens,grow", CASES[:3])
@pytest.mark.parametrize("drop", [None, (0.2, 5)])
def test_fused_attention_backward_vs_oracle_autograd(tc, B, H, S, dk, lens, grow, drop):
    D = H * dk
    qkv4, p, keymask = _case(B, H, S, dk, lens, grow=grow)
    scale = 1.0 / math.sqrt(dk)
    dctx = gb(B, S, D, seed=9, scale=0.5)
    q_ = qkv4.clone().requires_grad_(True)
    p_ = p.clone().requires_grad_(True)
    ref = oracle_attention(q_, p_, keymask, H, scale, drop, 24680)
    ref.backward(dctx)
    qc, pc = qkv4.cuda().to(torch.bfloat16), p.cuda().to(torch.bfloat16)
    ctx, bd, lse = tc.attn_fwd_fused(qc, pc, keymask.cuda(), H, scale, drop=drop)
    dq = torch.full_like(qc, float("nan"))
    dp = tc.attn_bwd_fused(dctx.cuda().to(torch.bfloat16), ctx, lse, bd, qc, pc, keymask.cuda(), H, scale, dq, drop=drop)
    assert torch.isfinite(dq.float()).all() and torch.isfinite(dp).all()
    # reference gradient w.r.t. qkv4 = [d(q+u) | d(q+v) | dk | dv]; bf16 operands everywhere: 3e-2 of each block's scale
    for name, blk in (("dqu", slice(0, D)), ("dqv", slice(D, 2 * D)), ("dk", slice(2 * D, 3 * D)), ("dv", slice(3 * D, 4 * D))):
        e = rel_err(dq[..., blk], q_.grad[..., blk])
        assert e < 3e-2, (name, e)
    assert rel_err(dp, p_.grad) < 3e-2, rel_err(dp, p_.grad)


def test_fused_attention_dropout_pattern_is_the_unfused_one(tc):
    """Same seed / site: the fused kernel keeps exactly the elements the unfused softmax kernel keeps (compare the
    dropped probabilities the backward writes with the unfused kernel's Pd)."""
    B, H, S, dk = 1, 2, 333, 64
    D = H * dk
    qkv4, p, keymask = _case(B, H, S, dk, [300])
    scale = 1.0 / math.sqrt(dk)
    qc, pc, kc = qkv4.cuda().to(torch.bfloat16), p.cuda().to(torch.bfloat16), keymask.cuda()
    ac, bdr = tc.attn_scores_fwd(qc, pc, H)
    Pm, Pd = tc.relpos_softmax_fwd(ac, bdr, kc, scale, drop=(0.3, 11))
    ctx, bd, lse = tc.attn_fwd_fused(qc, pc, kc, H, scale, drop=(0.3, 11))
    pd2 = tc._like(bd)
    ds, dbd = tc._like(bd), tc._like(bd)
    dq = torch.empty_like(qc)
    pr, seed, site = tc._drop((0.3, 11))
    dctx = torch.zeros(B, S, D, dtype=torch.bfloat16, device="cuda")
    delta = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
    _lib.call("a3t_relpos_attn_bwd", qc.data_ptr(), bd.data_ptr(), bd.stride(2), kc.view(torch.uint8).data_ptr(), ctx.data_ptr(),
              dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), pd2.data_ptr(), ds.data_ptr(), dbd.data_ptr(), B, H, S, D, scale, pr,
              seed, site, torch.cuda.current_stream().cuda_stream)
    a, b = Pd.float().cpu(), pd2.float().cpu()
    big = Pm.float().cpu() > 1e-4                    # where the undropped probability is not rounding noise
    assert torch.equal((a == 0)[big], (b == 0)[big])
    assert rel_err(b, a) < 2e-2
