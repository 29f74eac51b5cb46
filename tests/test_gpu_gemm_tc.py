"""tcgen05 tensor-core GEMM (csrc/gemm_tc.cu) against the exact CUDA-core kernel on the same bf16
operands, through the C-ABI (`a3t_gemm`), for every mode / operand-major combination the A3T step
uses: implicit conv (taps 1/3/5, ragged channel counts and sequence lengths), dgrad with ReLU mask,
wgrad with split-K, and the batched attention contractions.  fp32 accumulation on both sides, so
the only difference is summation order: tolerance 2e-3 relative to the output scale.
IMPL_TC makes the library fail (not fall back) when a shape does not qualify."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from a3t_b200 import _lib


@pytest.fixture(scope="module")
def bes(cuda_lib):
    from a3t_b200.backend import CudaBackend

    tc = CudaBackend("cuda:0", torch.bfloat16, seed=1234567, impl=_lib.IMPL_TC)
    simt = CudaBackend("cuda:0", torch.bfloat16, seed=1234567, impl=_lib.IMPL_SIMT)
    return tc, simt


def g(*shape, seed=0, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).to(dtype).cuda()


def close(a, b, tol=2e-3):
    a, b = a.float(), b.float()
    assert a.shape == b.shape
    scale = max(b.abs().max().item(), 1e-6)
    err = (a - b).abs().max().item()
    assert err <= tol * scale, f"max err {err} vs scale {scale}"


@pytest.mark.parametrize("taps,C,N,S,B", [
    (1, 384, 1536, 1152, 2),    # qkv4 / pointwise
    (3, 384, 1536, 1152, 2),    # FFN w_1
    (3, 1536, 384, 1152, 1),    # FFN w_2
    (1, 80, 384, 1024, 2),      # pre-net (K = 80)
    (1, 384, 80, 1024, 2),      # mel head (N = 80)
    (5, 80, 256, 200, 3),       # postnet first layer, ragged sequence
    (5, 256, 80, 200, 3),       # postnet last layer
    (3, 128, 512, 230, 2),      # cfg1 FFN: sequence length not a multiple of any tile
    (1, 384, 384, 1692, 1),     # cfg4 sequence length
])
def test_conv_family_tc(bes, taps, C, N, S, B):
    tc, simt = bes
    x = g(B, S, C, seed=1)
    w = g(N, C, taps, seed=2, scale=1.0 / math.sqrt(C * taps), dtype=torch.float32)
    bias = g(N, seed=3, dtype=torch.float32)
    res = g(B, S, N, seed=4, dtype=torch.float32)
    pw_t, pw_s = tc.pack_weight(w), simt.pack_weight(w)
    for kw in (dict(), dict(drop=(0.25, 7)), dict(relu=True, drop=(0.25, 8)),
               dict(drop=(0.5, 9), residual=res, out_scale=0.5), dict(out_dtype=torch.float32)):
        yt = tc.conv_fwd(x, pw_t, bias, **kw)
        ys = simt.conv_fwd(x, pw_s, bias, **kw)
        assert yt.dtype == ys.dtype
        close(yt, ys, tol=1e-2 if yt.dtype == torch.bfloat16 else 2e-3)
        if "drop" in kw and "residual" not in kw and "relu" not in kw:
            assert torch.equal(yt == 0, ys == 0)  # identical dropout pattern
    dy = g(B, S, N, seed=5)
    mask = (g(B, S, C, seed=6) > 0).to(torch.bfloat16)
    close(tc.conv_dgrad(dy, pw_t, out_dtype=torch.float32), simt.conv_dgrad(dy, pw_s, out_dtype=torch.float32))
    close(tc.conv_dgrad(dy, pw_t, mask=mask, mask_scale=1.25, out_dtype=torch.float32),
          simt.conv_dgrad(dy, pw_s, mask=mask, mask_scale=1.25, out_dtype=torch.float32))
    close(tc.conv_wgrad(dy, x, taps), simt.conv_wgrad(dy, x, taps))


@pytest.mark.parametrize("B,H,S,dk", [(2, 2, 1152, 192), (1, 2, 256, 64), (2, 1, 200, 64)])
def test_attention_contractions_tc(bes, B, H, S, dk):
    tc, simt = bes
    D = H * dk
    qkv4, p = g(B, S, 4 * D, seed=1, scale=0.5), g(S, D, seed=2, scale=0.5)
    ac_t, bd_t = tc.attn_scores_fwd(qkv4, p, H)
    ac_s, bd_s = simt.attn_scores_fwd(qkv4, p, H)
    close(ac_t, ac_s, tol=1e-2)  # bf16 score tensors: one ulp of the largest score
    close(bd_t, bd_s, tol=1e-2)
    Pd = torch.softmax(g(B, H, S, S, seed=3, dtype=torch.float32), -1).to(torch.bfloat16)
    close(tc.attn_pv_fwd(Pd, qkv4, H), simt.attn_pv_fwd(Pd, qkv4, H), tol=1e-2)
    dctx = g(B, S, D, seed=4)
    dq_t, dq_s = torch.zeros_like(qkv4), torch.zeros_like(qkv4)
    close(tc.attn_pv_bwd(dctx, Pd, qkv4, H, dq_t), simt.attn_pv_bwd(dctx, Pd, qkv4, H, dq_s), tol=1e-2)
    dS, dBD = g(B, H, S, S, seed=5, scale=0.1), g(B, H, S, S, seed=6, scale=0.1)
    dp_t = tc.attn_scores_bwd(dS, dBD, qkv4, p, H, dq_t)
    dp_s = simt.attn_scores_bwd(dS, dBD, qkv4, p, H, dq_s)
    close(dp_t, dp_s)
    close(dq_t, dq_s, tol=1e-2)


@pytest.mark.parametrize("B,H,S,dk", [(2, 2, 100, 64), (1, 2, 1692, 192)])
def test_attention_contractions_tc_any_length(bes, B, H, S, dk):
    """S % 8 != 0: the backend's score tensors carry a row pitch padded to 8 elements, so every attention
    contraction still qualifies for the tcgen05 kernel (IMPL_TC raises instead of falling back)."""
    tc, simt = bes
    D = H * dk
    qkv4, p = g(B, S, 4 * D, seed=1, scale=0.5), g(S, D, seed=2, scale=0.5)
    ac_t, bd_t = tc.attn_scores_fwd(qkv4, p, H)
    ac_s, bd_s = simt.attn_scores_fwd(qkv4, p, H)
    assert ac_t.stride(2) % 8 == 0 and ac_t.shape == (B, H, S, S)
    close(ac_t, ac_s, tol=1e-2)
    close(bd_t, bd_s, tol=1e-2)
    Pd = tc._scores(B, H, S, "cuda")
    Pd.copy_(torch.softmax(g(B, H, S, S, seed=3, dtype=torch.float32), -1))
    close(tc.attn_pv_fwd(Pd, qkv4, H), simt.attn_pv_fwd(Pd, qkv4, H), tol=1e-2)
    dctx = g(B, S, D, seed=4)
    dq_t, dq_s = torch.zeros_like(qkv4), torch.zeros_like(qkv4)
    close(tc.attn_pv_bwd(dctx, Pd, qkv4, H, dq_t), simt.attn_pv_bwd(dctx, Pd, qkv4, H, dq_s), tol=1e-2)
    dS, dBD = tc._scores(B, H, S, "cuda"), tc._scores(B, H, S, "cuda")
    dS.copy_(g(B, H, S, S, seed=5, scale=0.1))
    dBD.copy_(g(B, H, S, S, seed=6, scale=0.1))
    close(tc.attn_scores_bwd(dS, dBD, qkv4, p, H, dq_t), simt.attn_scores_bwd(dS, dBD, qkv4, p, H, dq_s))
    close(dq_t, dq_s, tol=1e-2)


def test_tc_is_what_auto_picks(cuda_lib):
    """AUTO dispatch must select the tensor-core kernel for the FFN shapes of the paper config."""
    from a3t_b200.backend import CudaBackend

    be = CudaBackend("cuda:0", torch.bfloat16)
    x = g(1, 1152, 384, seed=1)
    w = g(1536, 384, 3, seed=2, dtype=torch.float32)
    pw = be.pack_weight(w)
    d = be._desc(1152, 1536, 3 * 384, _lib.GEMM_CONV, taps=3, pad=1, seq=1152, cin=384, dtype_a=1, dtype_b=1, dtype_c=1,
                 sa_m=384, sa_k=1, sb_n=1152, sb_tap=384, sb_k=1, sc_m=1536, sc_n=1)
    y = torch.empty(1, 1152, 1536, dtype=torch.bfloat16, device="cuda")
    assert _lib.call("a3t_gemm_tc_supported", d, x.data_ptr(), pw.fwd.data_ptr(), y.data_ptr()) == 1


@pytest.mark.parametrize("taps,C,N,S,B", [(3, 384, 1536, 1152, 2), (1, 384, 384, 1152, 3), (5, 80, 256, 200, 3)])
def test_cta_pair_mode(bes, taps, C, N, S, B):
    """cta_group::2 (256 x N tiles on a CTA pair, forced with impl = IMPL_TC_PAIR): same results as the
    CUDA-core kernel, including an odd number of M tiles (phantom second tile) and split-K wgrad."""
    from a3t_b200.backend import CudaBackend

    _, simt = bes
    tc = CudaBackend("cuda:0", torch.bfloat16, seed=1234567, impl=_lib.IMPL_TC_PAIR)
    x = g(B, S, C, seed=11)
    w = g(N, C, taps, seed=12, scale=1.0 / math.sqrt(C * taps), dtype=torch.float32)
    bias = g(N, seed=13, dtype=torch.float32)
    res = g(B, S, N, seed=14, dtype=torch.float32)
    pw_t, pw_s = tc.pack_weight(w), simt.pack_weight(w)
    close(tc.conv_fwd(x, pw_t, bias, drop=(0.5, 9), residual=res, out_scale=0.5),
          simt.conv_fwd(x, pw_s, bias, drop=(0.5, 9), residual=res, out_scale=0.5))
    dy = g(B, S, N, seed=15)
    close(tc.conv_dgrad(dy, pw_t, out_dtype=torch.float32), simt.conv_dgrad(dy, pw_s, out_dtype=torch.float32))
    close(tc.conv_wgrad(dy, x, taps), simt.conv_wgrad(dy, x, taps))
