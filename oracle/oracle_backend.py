"""CPU oracle with the interface of `a3t_b200.backend.CudaBackend`.  TEST INFRASTRUCTURE ONLY.

Forward ops delegate to `oracle/a3t_oracle.py` (each cites the reference file:line it restates);
backward ops are obtained by differentiating those forward ops with torch autograd, so they are
the reference's own gradients by construction.  Used by tests to (1) check the hand-written
backward graph in `a3t_b200/graph.py` without a GPU and (2) check the CUDA ops one by one.
Nothing under `a3t_b200/` imports this file.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import a3t_oracle as O


class _PW:
    def __init__(self, w):
        self.w = w
        self.N, self.C, self.taps = w.shape


def _vjp(fn, inputs, gout):
    ins = [t.detach().clone().requires_grad_(True) for t in inputs]
    with torch.enable_grad():
        out = fn(*ins)
    return torch.autograd.grad(out, ins, gout, allow_unused=True)


class OracleBackend:
    def __init__(self, act_dtype=torch.float32, seed: int = 0, autograd: bool = False):
        """autograd=True keeps the torch graph through the forward ops so that `loss.backward()` gives the
        gradients (the CPU-baseline arm of bench.py times exactly that: forward ops + torch autograd)."""
        assert act_dtype == torch.float32, "the oracle computes in fp32"
        self.act_dtype = act_dtype
        self.seed = seed
        self.autograd = autograd

    def set_seed(self, seed):
        self.seed = seed

    def _d(self, drop):
        return None if drop is None or drop[0] <= 0 else (drop[0], self.seed, drop[1])

    # conv family -----------------------------------------------------------------------------
    def pack_weight(self, w):
        w3 = w if w.dim() == 3 else w.unsqueeze(-1)
        return _PW(w3 if self.autograd else w3.detach())

    def conv_fwd(self, x, pw, bias=None, *, relu=False, drop=None, residual=None, out_scale=1.0, out_dtype=None):
        return O.conv_fwd(x, pw.w, bias, relu=relu, drop=self._d(drop), residual=residual, out_scale=out_scale)

    def conv_dgrad(self, dy, pw, *, mask=None, mask_scale=1.0, out_dtype=None):
        B, S, _ = dy.shape
        x0 = torch.zeros(B, S, pw.C)
        (dx,) = _vjp(lambda x: O.conv_fwd(x, pw.w), [x0], dy)
        if mask is not None:
            dx = dx * (mask != 0).float() * mask_scale
        return dx.contiguous()

    def conv_wgrad(self, dy, x, taps, out=None, out_zeroed=False):
        N, C = dy.shape[-1], x.shape[-1]
        w0 = torch.zeros(N, C, taps)
        (dw,) = _vjp(lambda w: O.conv_fwd(x, w), [w0], dy)
        if out is not None:
            out.view(N, C, taps).copy_(dw)
            return out.view(N, C, taps)
        return dw

    def colsum(self, x):
        return O.colsum(x.float())

    # layernorm ------------------------------------------------------------------------------
    def ln_fwd(self, x, gamma, beta, eps, *, relu=False, out_scale=1.0, drop=None, out_dtype=None):
        y = O.ln_fwd(x, gamma, beta, eps, relu=relu, out_scale=out_scale, drop=self._d(drop))
        mean = x.mean(-1).reshape(-1)
        rstd = torch.rsqrt(x.var(-1, unbiased=False) + eps).reshape(-1)
        return y, mean, rstd

    def ln_bwd(self, dy, x, mean, rstd, gamma, beta, *, relu=False, out_scale=1.0, drop=None, dres=None, eps=None,
               nxt=None):
        dx, dg, db = _vjp(lambda a, g, b: O.ln_fwd(a, g, b, eps, relu=relu, out_scale=out_scale, drop=self._d(drop)),
                          [x, gamma, beta], dy)
        if dres is not None:
            dx = dx + dres
        if nxt is not None:  # the next backward section's grad prep (fused into the CUDA kernel)
            g = O.scale_dropout(dx, nxt[0], self._d(nxt[1]))
            return dx, dg, db, g, g.reshape(-1, g.shape[-1]).sum(0)
        return dx, dg, db

    def scale_dropout(self, x, scale, drop=None, out_dtype=None):
        return O.scale_dropout(x, scale, self._d(drop))

    def cast_act(self, x):
        return x

    # embedding ------------------------------------------------------------------------------
    def mask_input_fwd(self, speech, masked, mask_feature, out_dtype=None):
        return O.mask_input_fwd(speech, masked.bool(), mask_feature)

    def mask_input_bwd(self, dx, masked):
        return (dx * masked.bool().unsqueeze(-1).float()).reshape(-1, dx.shape[-1]).sum(0)

    def embed_assemble_fwd(self, speech_y, text, sseg, tseg, emb, seg, xscale, *, drop_speech=None, drop_text=None):
        if sseg is None:
            seg = torch.zeros(1, speech_y.shape[-1])
            sseg = torch.zeros(speech_y.shape[:2], dtype=torch.long)
            tseg = torch.zeros(text.shape, dtype=torch.long)
        return O.embed_assemble_fwd(speech_y, text, sseg, tseg, emb, seg, xscale, drop_speech=self._d(drop_speech),
                                    drop_text=self._d(drop_text))

    def embed_assemble_bwd(self, dxs, text, sseg, tseg, V, nseg, xscale, emb_pad, seg_pad, *, drop_speech=None,
                           drop_text=None):
        B, S, D = dxs.shape
        Tt = text.shape[1]
        Ts = S - Tt
        has_seg = sseg is not None
        if not has_seg:
            sseg = torch.zeros(B, Ts, dtype=torch.long)
            tseg = torch.zeros(B, Tt, dtype=torch.long)

        def f(sy, emb, seg):
            e = F.embedding(text, emb, padding_idx=emb_pad)
            sp = O._drop(sy.contiguous(), self._d(drop_speech)) + F.embedding(sseg, seg, padding_idx=seg_pad)
            tx = O._drop((e * xscale).contiguous(), self._d(drop_text)) + F.embedding(tseg, seg, padding_idx=seg_pad)
            return torch.cat([sp, tx], 1)

        dsy, demb, dseg = _vjp(f, [torch.zeros(B, Ts, D), torch.zeros(V, D), torch.zeros(nseg, D)], dxs)
        return dsy, demb, (dseg if has_seg else None)

    # attention ------------------------------------------------------------------------------
    def attn_scores_fwd(self, qkv4, p, H):
        return O.attn_scores_fwd(qkv4, p, H)

    def relpos_softmax_fwd(self, ac, bd_raw, keymask, scale, *, drop=None):
        return O.relpos_softmax_fwd(ac, bd_raw, keymask.bool(), scale, drop=self._d(drop))

    def attn_pv_fwd(self, pd, qkv4, H):
        return O.attn_pv_fwd(pd, qkv4, H)

    def attn_pv_bwd(self, dctx, pd, qkv4, H, dqkv4):
        dpd, dq = _vjp(lambda a, q: O.attn_pv_fwd(a, q, H), [pd, qkv4], dctx)
        D = qkv4.shape[-1] // 4
        dqkv4[..., 3 * D:] = dq[..., 3 * D:]
        return dpd

    def relpos_softmax_bwd(self, dPd, P, scale, *, drop=None):
        # differentiate through softmax using P itself: dS = P*(dPu - sum dPu*P)*scale
        if drop is not None and drop[0] > 0:
            keep = O.keep_mask(P.numel(), drop[0], self.seed, drop[1]).view(P.shape).float()
            dPu = dPd * keep * (1.0 / (1.0 - drop[0]))
        else:
            dPu = dPd
        dS = P * (dPu - (dPu * P).sum(-1, keepdim=True)) * scale
        (dBD,) = _vjp(lambda b: O.rel_shift(b), [torch.zeros_like(dS)], dS)
        return dS.contiguous(), dBD.contiguous()

    def attn_scores_bwd(self, dS, dBD, qkv4, p, H, dqkv4):
        def f(q, pp):
            ac, bd = O.attn_scores_fwd(q, pp, H)
            return (ac * dS).sum() + (bd * dBD).sum()

        ins = [qkv4.detach().clone().requires_grad_(True), p.detach().clone().requires_grad_(True)]
        with torch.enable_grad():
            out = f(*ins)
        dq, dp = torch.autograd.grad(out, ins)
        D = qkv4.shape[-1] // 4
        dqkv4[..., : 3 * D] = dq[..., : 3 * D]
        return dp

    # conv module ----------------------------------------------------------------------------
    def glu_dwconv_fwd(self, u, w, bias):
        return O.glu_dwconv_fwd(u, w, bias)

    def glu_dwconv_bwd(self, dz, u, w):
        C = w.shape[0]
        du, dw, db = _vjp(lambda a, ww, b: O.glu_dwconv_fwd(a, ww, b), [u, w, torch.zeros(C)], dz)
        return du, dw.reshape(C, 1, -1), db

    def bn_stats(self, z, running_mean, running_var, nbt, momentum, eps, training):
        return O.bn_stats(z, running_mean, running_var, nbt, momentum, eps, training)

    def bn_act_fwd(self, z, mean, rstd, gamma, beta, act, *, drop=None, residual=None, out_dtype=None):
        return O.bn_act_fwd(z, mean, rstd, gamma, beta, act, drop=self._d(drop), residual=residual)

    def bn_act_bwd(self, dy, z, mean, rstd, gamma, beta, act, training, *, drop=None, eps=1e-5):

        def f(zz, g, b):
            if training:
                zf = zz.reshape(-1, zz.shape[-1])
                m = zf.mean(0)
                v = zf.var(0, unbiased=False)
                r = torch.rsqrt(v + eps)
            else:
                m, r = mean, rstd
            return O.bn_act_fwd(zz, m, r, g, b, act, drop=self._d(drop))

        dz, dg, db = _vjp(f, [z, gamma, beta], dy)
        return dz, dg, db

    # loss -----------------------------------------------------------------------------------
    def masked_l1_fwd(self, before, after, y, mask):
        if after is None:
            l = (before - y).abs().sum(-1)
            m = mask.float()
            den = m.sum() + 1e-10
            return ((l * m).sum() / den).view(1), den.view(1)
        return O.masked_l1_fwd(before, after, y, mask.bool())

    def masked_l1_bwd(self, gloss, before, after, y, mask, den):
        if after is None:
            (db,) = _vjp(lambda b: self.masked_l1_fwd(b, None, y, mask)[0], [before], gloss.reshape(1))
            return db, None
        db, da = _vjp(lambda b, a: O.masked_l1_fwd(b, a, y, mask.bool())[0], [before, after], gloss.reshape(1))
        return db, da
