"""Tensor-level wrapper of the C-ABI (`include/a3t_b200.h`): torch tensors in, torch tensors out.

torch is used only for device memory, streams and shapes; every arithmetic op below is a call
into `liba3t_b200.so`.  The op names and argument meaning mirror `oracle/a3t_oracle.py` (the CPU
checker used by the tests) so the two can be compared call by call.

`act_dtype` selects the precision mode:
  * torch.float32  - every tensor fp32, GEMMs on the exact-fp32 CUDA-core kernel (parity mode);
  * torch.bfloat16 - GEMM operands bf16 (tcgen05 tensor-core kernel, fp32 accumulate in TMEM),
                     residual stream / LayerNorm / softmax / BatchNorm / loss math in fp32.
"""
from __future__ import annotations

import contextlib
import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import A3T_BF16, A3T_F32, GEMM_CONV, GEMM_PLAIN, GEMM_WGRAD, GemmDesc, call

ACT_NONE, ACT_SWISH, ACT_TANH = _lib.ACT_NONE, _lib.ACT_SWISH, _lib.ACT_TANH


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return A3T_F32
    if t.dtype == torch.bfloat16:
        return A3T_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _u8(t: torch.Tensor) -> torch.Tensor:
    """Mask tensor -> contiguous uint8 (no copy for bool / uint8).  Any other dtype (int64 / float masks built by
    hand) is converted with `!= 0`, which is what the reference's `masked_fill` / `eq(0)` see."""
    if t.dtype == torch.bool:
        return t.contiguous().view(torch.uint8)
    if t.dtype == torch.uint8:
        return t.contiguous()
    return (t != 0).to(torch.uint8).contiguous()


class PackedWeight:
    """A conv/linear weight (N, C, taps) in the layouts the GEMM kernels read.

    fp32 mode: the parameter itself (strided reads).  bf16 mode: two K-major bf16 packs,
    `fwd[n, tap*C + c]` and `dgrad[c, tap'*N + n]` (flipped taps), rebuilt after each optimizer step.
    """

    __slots__ = ("w", "N", "C", "taps", "fwd", "dgrad")

    def __init__(self, w, N, C, taps, fwd=None, dgrad=None):
        self.w, self.N, self.C, self.taps, self.fwd, self.dgrad = w, N, C, taps, fwd, dgrad


class CudaBackend:
    def __init__(self, device, act_dtype=torch.float32, seed: int = 0, impl: int = _lib.IMPL_AUTO):
        _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.A3TError("a3t_b200 runs on CUDA devices only (no CPU fallback)")
        self.act_dtype = act_dtype
        self.impl = impl if act_dtype == torch.bfloat16 else _lib.IMPL_SIMT
        self.seed = torch.tensor([seed], dtype=torch.int64, device=self.device)
        # zero-initialised fp32 arena for the small outputs kernels ACCUMULATE into with atomics (bias / LayerNorm
        # parameter gradients): one memset per backward pass instead of a second reduction kernel per output.
        # Slices are valid until the next `begin_backward()`.
        self._arena = torch.zeros(1 << 20, dtype=torch.float32, device=self.device)
        self._arena_off = 0
        # bit 0: token id outside the embedding table, bit 1: segment id outside it (set by the embedding kernel,
        # read by `check_ids()`; torch raises IndexError at the same place)
        self.err_flag = torch.zeros(1, dtype=torch.int32, device=self.device)

    # ------------------------------------------------------------------ helpers
    def begin_backward(self):
        """Zero the accumulation arena (called once at the start of graph.backward)."""
        self._arena.zero_()
        self._arena_off = 0

    def _zeros(self, n: int) -> torch.Tensor:
        n4 = (n + 3) & ~3
        if self._arena_off + n4 > self._arena.numel():
            return torch.zeros(n, dtype=torch.float32, device=self.device)
        t = self._arena[self._arena_off:self._arena_off + n]
        self._arena_off += n4
        return t

    def owns(self, t: torch.Tensor) -> bool:
        """True when `t` aliases the accumulation arena (callers that keep it past the step must clone)."""
        a = self._arena
        return a.data_ptr() <= t.data_ptr() < a.data_ptr() + a.numel() * 4

    def set_seed(self, seed: int):
        self.seed.fill_(seed)

    def advance_seed(self):
        call("a3t_seed_advance", _p(self.seed), _stream(self.seed))

    def seed_snapshot(self) -> torch.Tensor:
        """Device copy of the current dropout seed.  Dropout masks are never stored: every kernel re-derives them
        from the seed it is handed, so a backward pass that may run after the master seed has advanced
        (`model(**batch)` ... `loss.backward()`) must be given the seed its forward used."""
        return self.seed.clone()

    @contextlib.contextmanager
    def using_seed(self, seed: torch.Tensor):
        """Run the enclosed kernels with `seed` (a device int64[1]) instead of the master seed."""
        prev, self.seed = self.seed, seed
        try:
            yield
        finally:
            self.seed = prev

    def check_ids(self):
        """Raise IndexError if an embedding kernel saw a token / segment id outside its table since the last
        check (synchronises; call it where the host synchronises anyway, e.g. inference)."""
        f = int(self.err_flag.item())
        if f:
            self.err_flag.zero_()
            what = [w for b, w in ((1, "token id outside the text embedding table"),
                                   (2, "segment id outside the segment embedding table (more than 499 phones?)")) if f & b]
            raise IndexError("a3t_b200: " + "; ".join(what))

    def _drop(self, drop):
        if drop is None or drop[0] <= 0.0:
            return 0.0, None, 0
        return float(drop[0]), _p(self.seed), int(drop[1])

    def _gemm(self, d: GemmDesc, A, B, Cout, bias=None, res=None, mask=None, a_off=0, b_off=0, c_off=0):
        d.impl = self.impl
        call("a3t_gemm", d, A.data_ptr() + a_off * A.element_size(), B.data_ptr() + b_off * B.element_size(),
             Cout.data_ptr() + c_off * Cout.element_size(), _p(bias), _p(res), _p(mask),
             _p(self.seed) if d.drop_p > 0 else None, _stream(Cout))

    @staticmethod
    def _desc(M, N, K, mode=GEMM_PLAIN, **kw) -> GemmDesc:
        d = GemmDesc()
        d.M, d.N, d.K, d.mode = M, N, K, mode
        d.taps, d.pad, d.seq, d.cin = 1, 0, max(M, 1), K
        d.batch1 = d.batch2 = 1
        d.alpha = d.out_scale = d.mask_scale = 1.0
        for k, v in kw.items():
            setattr(d, k, v)
        return d

    # ------------------------------------------------------------------ weights
    def pack_weight(self, w: torch.Tensor) -> PackedWeight:
        w3 = w if w.dim() == 3 else w.unsqueeze(-1)
        N, Cc, taps = w3.shape
        w3 = w3.contiguous()
        if self.act_dtype == torch.float32:
            return PackedWeight(w3, N, Cc, taps)
        fwd = torch.empty(N, taps * Cc, dtype=torch.bfloat16, device=w.device)
        dg = torch.empty(Cc, taps * N, dtype=torch.bfloat16, device=w.device)
        call("a3t_pack_conv_weight", _p(w3), N, Cc, taps, _p(fwd), _p(dg), _stream(w3))
        return PackedWeight(w3, N, Cc, taps, fwd, dg)

    def build_pack_plan(self, entries):
        """entries: list of (source weights stacked along N, PackedWeight with persistent bf16 packs).
        Returns the device table `a3t_pack_conv_weights` consumes, or None when an entry cannot be batched."""
        if self.act_dtype != torch.bfloat16 or not entries:
            return None
        items = (_lib.PackItem * len(entries))()
        tile, max_taps = 0, 1
        for i, (srcs, pw) in enumerate(entries):
            if pw.fwd is None or pw.N % len(srcs) or pw.taps > 11:
                return None
            it = items[i]
            for j in range(4):
                it.w[j] = srcs[j].data_ptr() if j < len(srcs) else None
            it.fwd, it.dgrad = pw.fwd.data_ptr(), pw.dgrad.data_ptr()
            it.N, it.C, it.taps, it.seg_rows = pw.N, pw.C, pw.taps, pw.N // len(srcs)
            it.tile_start, it.tiles_c = tile, (pw.C + 63) // 64
            tile += ((pw.N + 63) // 64) * it.tiles_c
            max_taps = max(max_taps, pw.taps)
        table = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8).to(self.device)
        return dict(table=table, n=len(entries), tiles=tile, max_taps=max_taps, keep=entries)

    def repack(self, plan):
        call("a3t_pack_conv_weights", plan["table"].data_ptr(), plan["n"], plan["tiles"], plan["max_taps"],
             _stream(plan["table"]))

    def qkv4_bias(self, bq, bk, bv, u, v, out):
        call("a3t_qkv4_bias", _p(bq), _p(bk), _p(bv), _p(u), _p(v), _p(out), bq.numel(), _stream(out))

    # ------------------------------------------------------------------ conv / linear family
    def conv_fwd(self, x, pw: PackedWeight, bias=None, *, relu=False, drop=None, residual=None, out_scale=1.0,
                 out_dtype=None):
        """y = residual + out_scale * dropout(relu(conv(x, w) + bias)); x (B,S,C) channels-last."""
        Bn, S, Cc = x.shape
        assert Cc == pw.C and x.is_contiguous()
        out_dtype = out_dtype or (torch.float32 if residual is not None else self.act_dtype)
        y = torch.empty(Bn, S, pw.N, dtype=out_dtype, device=x.device)
        p, _, site = self._drop(drop)
        d = self._desc(Bn * S, pw.N, pw.taps * Cc, GEMM_CONV if pw.taps > 1 else GEMM_PLAIN, taps=pw.taps,
                       pad=(pw.taps - 1) // 2, seq=S, cin=Cc, relu=int(relu), out_scale=out_scale, drop_p=p,
                       drop_site=site, dtype_a=_dt(x), dtype_c=_dt(y), sa_m=Cc, sa_k=1, sc_m=pw.N, sc_n=1,
                       sr_m=pw.N, sr_n=1)
        if pw.fwd is not None and x.dtype == torch.bfloat16:
            Bm = pw.fwd
            d.dtype_b, d.sb_n, d.sb_tap, d.sb_k = A3T_BF16, pw.taps * Cc, Cc, 1
        else:
            Bm = pw.w
            d.dtype_b, d.sb_n, d.sb_tap, d.sb_k = A3T_F32, Cc * pw.taps, 1, pw.taps
        if residual is not None:
            assert residual.dtype == torch.float32 and residual.is_contiguous()
        self._gemm(d, x, Bm, y, bias, residual)
        return y

    def conv_dgrad(self, dy, pw: PackedWeight, *, mask=None, mask_scale=1.0, out_dtype=None):
        """dx[b,t,c] = sum_{tap,n} dy[b,t-(tap-pad),n] w[n,c,tap]; optional (mask != 0) * mask_scale."""
        Bn, S, N = dy.shape
        assert N == pw.N and dy.is_contiguous()
        out_dtype = out_dtype or self.act_dtype
        dx = torch.empty(Bn, S, pw.C, dtype=out_dtype, device=dy.device)
        d = self._desc(Bn * S, pw.C, pw.taps * N, GEMM_CONV if pw.taps > 1 else GEMM_PLAIN, taps=pw.taps,
                       pad=(pw.taps - 1) // 2, seq=S, cin=N, mask_scale=mask_scale, dtype_a=_dt(dy), dtype_c=_dt(dx),
                       sa_m=N, sa_k=1, sc_m=pw.C, sc_n=1)
        b_off = 0
        if pw.dgrad is not None and dy.dtype == torch.bfloat16:
            Bm = pw.dgrad
            d.dtype_b, d.sb_n, d.sb_tap, d.sb_k = A3T_BF16, pw.taps * N, N, 1
        else:
            Bm = pw.w
            d.dtype_b, d.sb_n, d.sb_tap, d.sb_k = A3T_F32, pw.taps, -1, pw.C * pw.taps
            b_off = pw.taps - 1
        if mask is not None:
            assert mask.shape == dx.shape and mask.is_contiguous()
            d.dtype_mask = _dt(mask)
        self._gemm(d, dy, Bm, dx, None, None, mask, b_off=b_off)
        return dx

    def conv_wgrad(self, dy, x, taps: int, out=None, out_zeroed=False):
        """dW[n,c,tap] = sum_{b,t} dy[b,t,n] x[b,t+tap-pad,c]  -> (N, C, taps) fp32.  `out`: caller-owned
        destination (e.g. a view of the trainer's flat gradient buffer); `out_zeroed`: it already holds zeros, so a
        split-K launch needs no clear."""
        Bn, S, N = dy.shape
        Cc = x.shape[-1]
        assert x.shape[:2] == dy.shape[:2] and dy.is_contiguous() and x.is_contiguous()
        if out is None:
            dW = torch.empty(N, Cc, taps, dtype=torch.float32, device=dy.device)
        else:
            assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == N * Cc * taps
            dW = out.view(N, Cc, taps)
        d = self._desc(N, taps * Cc, Bn * S, GEMM_WGRAD, taps=taps, pad=(taps - 1) // 2, seq=S, cin=Cc,
                       dtype_a=_dt(dy), dtype_b=_dt(x), dtype_c=A3T_F32, sa_k=N, sa_m=1, sb_k=Cc, sb_n=1,
                       sc_m=Cc * taps, sc_n=taps, sc_tap=1, c_zeroed=int(out is not None and out_zeroed))
        self._gemm(d, dy, x, dW)
        return dW

    def colsum(self, x):
        C_ = x.shape[-1]
        rows = x.numel() // C_
        vn = 8 if x.dtype == torch.bfloat16 else 4
        if C_ % vn == 0 and x.data_ptr() % 16 == 0:
            out = self._zeros(C_)  # single kernel, atomics into the zeroed arena
            call("a3t_colsum", _p(x), _dt(x), _p(out), None, rows, C_, C_, _stream(x))
            return out
        out = torch.empty(C_, dtype=torch.float32, device=x.device)
        nblk = call("a3t_colsum_blocks", rows)
        partial = torch.empty(nblk * C_, dtype=torch.float32, device=x.device)
        call("a3t_colsum", _p(x), _dt(x), _p(out), _p(partial), rows, C_, C_, _stream(x))
        return out

    # ------------------------------------------------------------------ LayerNorm
    def ln_fwd(self, x, gamma, beta, eps, *, relu=False, out_scale=1.0, drop=None, out_dtype=None):
        C_ = x.shape[-1]
        rows = x.numel() // C_
        assert x.dtype == torch.float32 and x.is_contiguous()
        y = torch.empty(x.shape, dtype=out_dtype or self.act_dtype, device=x.device)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        p, seed, site = self._drop(drop)
        call("a3t_layernorm_fwd", _p(x), _p(gamma), _p(beta), _p(y), _dt(y), _p(mean), _p(rstd), rows, C_, eps,
             int(relu), out_scale, p, seed, site, _stream(x))
        return y, mean, rstd

    def ln_bwd(self, dy, x, mean, rstd, gamma, beta, *, relu=False, out_scale=1.0, drop=None, dres=None, eps=None,
               nxt=None):
        """Returns (dx, dgamma, dbeta).  With nxt = (scale, drop) also the next backward section's grad prep,
        fused into the same pass: (dx, dgamma, dbeta, g, gsum) with g = dropout'(dx * scale) in the GEMM dtype
        and gsum = column sums of g."""
        C_ = x.shape[-1]
        rows = x.numel() // C_
        assert dy.is_contiguous() and x.is_contiguous()
        dx = torch.empty_like(x)
        acc = self._zeros((3 if nxt is not None else 2) * C_)
        dgamma, dbeta = acc[:C_], acc[C_:2 * C_]
        p, seed, site = self._drop(drop)
        g = gsum = None
        gscale, gp, gsite = 1.0, 0.0, 0
        if nxt is not None:
            gscale = float(nxt[0])
            gp, gseed, gsite = self._drop(nxt[1])
            seed = seed or gseed
            g = torch.empty(x.shape, dtype=self.act_dtype, device=x.device)
            gsum = acc[2 * C_:]
        call("a3t_layernorm_bwd", _p(dy), _dt(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(dres), _p(dx),
             _p(dgamma), _p(dbeta), None, rows, C_, int(relu), out_scale, p, seed, site, _p(g),
             _dt(g) if g is not None else A3T_F32, gscale, gp, gsite, _p(gsum), _stream(x))
        if nxt is not None:
            return dx, dgamma, dbeta, g, gsum
        return dx, dgamma, dbeta

    def scale_dropout(self, x, scale, drop=None, out_dtype=None):
        assert x.dtype == torch.float32 and x.is_contiguous()
        y = torch.empty(x.shape, dtype=out_dtype or self.act_dtype, device=x.device)
        p, seed, site = self._drop(drop)
        call("a3t_scale_dropout", _p(x), _p(y), _dt(y), x.numel(), scale, p, seed, site, _stream(x))
        return y

    def cast_act(self, x):
        """fp32 -> activation dtype (identity in fp32 mode)."""
        if x.dtype == self.act_dtype:
            return x
        return self.scale_dropout(x, 1.0, None, out_dtype=self.act_dtype)

    # ------------------------------------------------------------------ embedding
    def mask_input_fwd(self, speech, masked, mask_feature, out_dtype=None):
        Bn, Ts, C_ = speech.shape
        assert speech.dtype == torch.float32 and speech.is_contiguous()
        y = torch.empty(speech.shape, dtype=out_dtype or self.act_dtype, device=speech.device)
        call("a3t_mask_input_fwd", _p(speech), _p(_u8(masked)), _p(mask_feature), _p(y), _dt(y), Bn * Ts, C_,
             _stream(speech))
        return y

    def mask_input_bwd(self, dx, masked):
        C_ = dx.shape[-1]
        rows = dx.numel() // C_
        out = torch.empty(C_, dtype=torch.float32, device=dx.device)
        nblk = call("a3t_colsum_blocks", rows)
        partial = torch.empty(nblk * C_, dtype=torch.float32, device=dx.device)
        call("a3t_mask_input_bwd", _p(dx), _p(_u8(masked)), _p(out), _p(partial), rows, C_, _stream(dx))
        return out

    def embed_assemble_fwd(self, speech_y, text, sseg, tseg, emb, seg, xscale, *, drop_speech=None, drop_text=None):
        Bn, Ts, D = speech_y.shape
        Tt = text.shape[1]
        xs = torch.empty(Bn, Ts + Tt, D, dtype=torch.float32, device=speech_y.device)
        p, seed, s1 = self._drop(drop_speech)
        _, _, s2 = self._drop(drop_text)
        call("a3t_embed_assemble_fwd", _p(speech_y), _p(text), _p(sseg), _p(tseg), _p(emb), _p(seg), _p(xs), Bn, Ts,
             Tt, D, xscale, p, seed, s1, s2, emb.shape[0], seg.shape[0] if seg is not None else 0, _p(self.err_flag),
             _stream(xs))
        return xs

    def embed_assemble_bwd(self, dxs, text, sseg, tseg, V, nseg, xscale, emb_pad, seg_pad, *, drop_speech=None,
                           drop_text=None):
        Bn, S, D = dxs.shape
        Tt = text.shape[1]
        Ts = S - Tt
        dsy = torch.empty(Bn, Ts, D, dtype=torch.float32, device=dxs.device)
        demb = torch.zeros(V, D, dtype=torch.float32, device=dxs.device)
        dseg = torch.zeros(nseg, D, dtype=torch.float32, device=dxs.device) if sseg is not None else None
        p, seed, s1 = self._drop(drop_speech)
        _, _, s2 = self._drop(drop_text)
        call("a3t_embed_assemble_bwd", _p(dxs), _p(text), _p(sseg), _p(tseg), _p(dsy), _p(demb), _p(dseg), Bn, Ts, Tt,
             D, xscale, emb_pad, seg_pad, p, seed, s1, s2, V, nseg, _stream(dxs))
        return dsy, demb, dseg

    # ------------------------------------------------------------------ attention
    def _scores(self, Bn, H, S, device):
        """(B,H,S,S) score tensor in the activation dtype.  In the tensor-core mode its rows are padded to a
        multiple of 8 elements (a strided view of a (B,H,S,Sp) buffer): TMA tensor maps need 16-byte aligned
        rows, and real batches have arbitrary S."""
        Sp = S if self.act_dtype == torch.float32 else (S + 7) // 8 * 8
        return torch.empty(Bn, H, S, Sp, dtype=self.act_dtype, device=device)[..., :S]

    @staticmethod
    def _like(t):
        """Uninitialised tensor with t's shape AND strides (keeps the row pitch of a score tensor)."""
        return torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, device=t.device)

    @classmethod
    def _as(cls, x, ref):
        """x laid out with ref's strides (copy only when a caller hands in a differently pitched tensor)."""
        if x.stride() == ref.stride():
            return x
        y = cls._like(ref)
        y.copy_(x)
        return y

    @staticmethod
    def _sstr(t):
        """(row pitch, batch stride, head stride) of a score tensor."""
        return t.stride(2), t.stride(0), t.stride(1)

    def attn_scores_fwd(self, qkv4, p, H):
        """qkv4 (B,S,4D) = [q+u | q+v | k | v]; p (S,D).  AC = (q+u)k^T, BDraw = (q+v)p^T, (B,H,S,S)."""
        Bn, S, D4 = qkv4.shape
        D = D4 // 4
        dk = D // H
        # score tensors in the activation dtype: fp32 in the parity mode; bf16 in the tensor-core mode, where P
        # itself is stored in bf16 (halves the bytes the memory-bound softmax kernel reads)
        ac, bd = self._scores(Bn, H, S, qkv4.device), self._scores(Bn, H, S, qkv4.device)
        ld, sb1, sb2 = self._sstr(ac)
        dt = _dt(qkv4)
        common = dict(batch1=Bn, batch2=H, dtype_a=dt, dtype_b=dt, dtype_c=_dt(ac), sa_m=D4, sa_k=1, sa_b1=S * D4,
                      sa_b2=dk, sc_m=ld, sc_n=1, sc_b1=sb1, sc_b2=sb2)
        d = self._desc(S, S, dk, sb_n=D4, sb_k=1, sb_b1=S * D4, sb_b2=dk, **common)
        self._gemm(d, qkv4, qkv4, ac, a_off=0, b_off=2 * D)
        d = self._desc(S, S, dk, sb_n=p.stride(0), sb_k=1, sb_b1=0, sb_b2=dk, **common)
        self._gemm(d, qkv4, p, bd, a_off=D)
        return ac, bd

    def relpos_softmax_fwd(self, ac, bd_raw, keymask, scale, *, drop=None):
        Bn, H, S, _ = ac.shape
        assert ac.stride(3) == 1
        bd_raw = self._as(bd_raw, ac)
        P = self._like(ac)
        p, seed, site = self._drop(drop)
        Pd = self._like(ac) if p > 0 else P
        ld = ac.stride(2)
        call("a3t_relpos_softmax_fwd", _p(ac), _p(bd_raw), _dt(ac), _p(_u8(keymask)), _p(P), _p(Pd), _dt(P), Bn, H, S, ld,
             scale, p, seed, site, _stream(ac))
        return P, Pd

    def attn_pv_fwd(self, pd, qkv4, H):
        Bn, S, D4 = qkv4.shape
        D = D4 // 4
        dk = D // H
        ld, sb1, sb2 = self._sstr(pd)
        ctx = torch.empty(Bn, S, D, dtype=self.act_dtype, device=qkv4.device)
        d = self._desc(S, dk, S, batch1=Bn, batch2=H, dtype_a=_dt(pd), dtype_b=_dt(qkv4), dtype_c=_dt(ctx), sa_m=ld,
                       sa_k=1, sa_b1=sb1, sa_b2=sb2, sb_n=1, sb_k=D4, sb_b1=S * D4, sb_b2=dk, sc_m=D, sc_n=1,
                       sc_b1=S * D, sc_b2=dk)
        self._gemm(d, pd, qkv4, ctx, b_off=3 * D)
        return ctx

    def attn_pv_bwd(self, dctx, pd, qkv4, H, dqkv4):
        """dPd = dctx v^T;  dV = Pd^T dctx -> dqkv4[..., 3D:]."""
        Bn, S, D4 = qkv4.shape
        D = D4 // 4
        dk = D // H
        dPd = self._scores(Bn, H, S, qkv4.device)
        ld, sb1, sb2 = self._sstr(dPd)
        d = self._desc(S, S, dk, batch1=Bn, batch2=H, dtype_a=_dt(dctx), dtype_b=_dt(qkv4), dtype_c=_dt(dPd), sa_m=D,
                       sa_k=1, sa_b1=S * D, sa_b2=dk, sb_n=D4, sb_k=1, sb_b1=S * D4, sb_b2=dk, sc_m=ld, sc_n=1,
                       sc_b1=sb1, sc_b2=sb2)
        self._gemm(d, dctx, qkv4, dPd, b_off=3 * D)
        ld, sb1, sb2 = self._sstr(pd)
        d = self._desc(S, dk, S, batch1=Bn, batch2=H, dtype_a=_dt(pd), dtype_b=_dt(dctx), dtype_c=_dt(dqkv4), sa_m=1,
                       sa_k=ld, sa_b1=sb1, sa_b2=sb2, sb_n=1, sb_k=D, sb_b1=S * D, sb_b2=dk, sc_m=D4, sc_n=1,
                       sc_b1=S * D4, sc_b2=dk)
        self._gemm(d, pd, dctx, dqkv4, c_off=3 * D)
        return dPd

    def relpos_softmax_bwd(self, dPd, P, scale, *, drop=None):
        Bn, H, S, _ = P.shape
        assert P.stride(3) == 1
        dPd = self._as(dPd, P)
        dS, dBD = self._like(P), self._like(P)
        p, seed, site = self._drop(drop)
        ld = P.stride(2)
        call("a3t_relpos_softmax_bwd", _p(dPd), _dt(dPd), _p(P), _dt(P), _p(dS), _p(dBD), _dt(dS), Bn, H, S, ld, scale, p,
             seed, site, _stream(P))
        return dS, dBD

    def attn_scores_bwd(self, dS, dBD, qkv4, p, H, dqkv4):
        """dqu = dS k -> [0:D]; dqv = dBD p -> [D:2D]; dk = dS^T qu -> [2D:3D]; returns dp (S,D) fp32."""
        Bn, S, D4 = qkv4.shape
        D = D4 // 4
        dk = D // H
        dts, dtq = _dt(dS), _dt(qkv4)
        dBD = self._as(dBD, dS)
        ld, sb1, sb2 = self._sstr(dS)
        a_row = dict(sa_m=ld, sa_k=1, sa_b1=sb1, sa_b2=sb2)   # A[i, j]
        a_col = dict(sa_m=1, sa_k=ld, sa_b1=sb1, sa_b2=sb2)   # A^T
        c_qkv = dict(sc_m=D4, sc_n=1, sc_b1=S * D4, sc_b2=dk)
        b_qkv = dict(sb_n=1, sb_k=D4, sb_b1=S * D4, sb_b2=dk)
        base = dict(batch1=Bn, batch2=H, dtype_a=dts, dtype_b=dtq, dtype_c=_dt(dqkv4))
        self._gemm(self._desc(S, dk, S, **base, **a_row, **b_qkv, **c_qkv), dS, qkv4, dqkv4, b_off=2 * D, c_off=0)
        self._gemm(self._desc(S, dk, S, **base, **a_col, **b_qkv, **c_qkv), dS, qkv4, dqkv4, b_off=0, c_off=2 * D)
        d = self._desc(S, dk, S, batch1=Bn, batch2=H, dtype_a=dts, dtype_b=_dt(p), dtype_c=_dt(dqkv4), **a_row,
                       sb_n=1, sb_k=p.stride(0), sb_b1=0, sb_b2=dk, **c_qkv)
        self._gemm(d, dBD, p, dqkv4, c_off=D)
        tmp = torch.empty(Bn, S, D, dtype=torch.float32, device=qkv4.device)
        d = self._desc(S, dk, S, batch1=Bn, batch2=H, dtype_a=dts, dtype_b=dtq, dtype_c=A3T_F32, **a_col, **b_qkv,
                       sc_m=D, sc_n=1, sc_b1=S * D, sc_b2=dk)
        self._gemm(d, dBD, qkv4, tmp, b_off=D)
        if Bn == 1:
            return tmp.view(S, D)
        return self.colsum(tmp.view(Bn, S * D)).view(S, D)

    # ------------------------------------------------------------------ fused attention (bf16 / tcgen05)
    def attn_fused_ok(self, Bn, H, S, D) -> bool:
        """True when the fused tcgen05 attention kernels take this shape (bf16 mode, head width 64 / 128 / 192)."""
        return (self.act_dtype == torch.bfloat16 and self.impl != _lib.IMPL_SIMT
                and call("a3t_attn_fused_supported", Bn, H, S, D) == 1)

    def attn_fwd_fused(self, qkv4, p, keymask, H, scale, *, drop=None):
        """ctx = attention(qkv4, p) with the scores kept on chip.  Returns (ctx (B,S,D), bd_raw, lse): bd_raw =
        (q+v) p^T is the one (B,H,S,S) tensor still materialised (bf16, by the batched GEMM); both it and lse are
        what the fused backward recomputes the probabilities from."""
        Bn, S, D4 = qkv4.shape
        D = D4 // 4
        dk = D // H
        bd = self._scores(Bn, H, S, qkv4.device)
        ld, sb1, sb2 = self._sstr(bd)
        dt = _dt(qkv4)
        d = self._desc(S, S, dk, batch1=Bn, batch2=H, dtype_a=dt, dtype_b=dt, dtype_c=_dt(bd), sa_m=D4, sa_k=1,
                       sa_b1=S * D4, sa_b2=dk, sc_m=ld, sc_n=1, sc_b1=sb1, sc_b2=sb2, sb_n=p.stride(0), sb_k=1, sb_b1=0,
                       sb_b2=dk)
        self._gemm(d, qkv4, p, bd, a_off=D)
        ctx = torch.empty(Bn, S, D, dtype=torch.bfloat16, device=qkv4.device)
        lse = torch.empty(Bn, H, S, dtype=torch.float32, device=qkv4.device)
        pr, seed, site = self._drop(drop)
        call("a3t_relpos_attn_fwd", _p(qkv4), _p(bd), ld, _p(_u8(keymask)), _p(ctx), _p(lse), Bn, H, S, D, scale, pr, seed,
             site, _stream(qkv4))
        return ctx, bd, lse

    def attn_bwd_fused(self, dctx, ctx, lse, bd, qkv4, p, keymask, H, scale, dqkv4, *, drop=None):
        """Backward of attn_fwd_fused: fills dqkv4 = [d(q+u) | d(q+v) | dk | dv] and returns dp (S,D) fp32."""
        Bn, S, D4 = qkv4.shape
        D = D4 // 4
        dk = D // H
        dev = qkv4.device
        pd, ds, dbd = self._like(bd), self._like(bd), self._like(bd)
        ld, sb1, sb2 = self._sstr(bd)
        pr, seed, site = self._drop(drop)
        delta = torch.empty(Bn, H, S, dtype=torch.float32, device=dev)
        call("a3t_relpos_attn_bwd", _p(qkv4), _p(bd), ld, _p(_u8(keymask)), _p(ctx), _p(dctx), _p(lse), _p(delta), _p(dqkv4),
             _p(pd), _p(ds), _p(dbd), Bn, H, S, D, scale, pr, seed, site, _stream(qkv4))
        dts, dtq = _dt(ds), _dt(qkv4)
        a_row = dict(sa_m=ld, sa_k=1, sa_b1=sb1, sa_b2=sb2)   # A[i, j]
        a_col = dict(sa_m=1, sa_k=ld, sa_b1=sb1, sa_b2=sb2)   # A^T
        c_qkv = dict(sc_m=D4, sc_n=1, sc_b1=S * D4, sc_b2=dk)
        b_qkv = dict(sb_n=1, sb_k=D4, sb_b1=S * D4, sb_b2=dk)
        base = dict(batch1=Bn, batch2=H, dtype_a=dts, dtype_b=dtq, dtype_c=_dt(dqkv4))
        # dv = Pd^T dctx
        d = self._desc(S, dk, S, batch1=Bn, batch2=H, dtype_a=dts, dtype_b=_dt(dctx), dtype_c=_dt(dqkv4), **a_col, sb_n=1,
                       sb_k=D, sb_b1=S * D, sb_b2=dk, **c_qkv)
        self._gemm(d, pd, dctx, dqkv4, c_off=3 * D)
        # dk = dS^T (q+u);  d(q+v) = dBD_raw p;  dp = sum_b dBD_raw^T (q+v)
        self._gemm(self._desc(S, dk, S, **base, **a_col, **b_qkv, **c_qkv), ds, qkv4, dqkv4, b_off=0, c_off=2 * D)
        d = self._desc(S, dk, S, batch1=Bn, batch2=H, dtype_a=dts, dtype_b=_dt(p), dtype_c=_dt(dqkv4), **a_row, sb_n=1,
                       sb_k=p.stride(0), sb_b1=0, sb_b2=dk, **c_qkv)
        self._gemm(d, dbd, p, dqkv4, c_off=D)
        tmp = torch.empty(Bn, S, D, dtype=torch.float32, device=dev)
        d = self._desc(S, dk, S, batch1=Bn, batch2=H, dtype_a=dts, dtype_b=dtq, dtype_c=A3T_F32, **a_col, **b_qkv, sc_m=D,
                       sc_n=1, sc_b1=S * D, sc_b2=dk)
        self._gemm(d, dbd, qkv4, tmp, b_off=D)
        if Bn == 1:
            return tmp.view(S, D)
        return self.colsum(tmp.view(Bn, S * D)).view(S, D)

    # ------------------------------------------------------------------ conv module
    def glu_dwconv_fwd(self, u, w, bias):
        Bn, S, C2 = u.shape
        C_ = C2 // 2
        k = w.shape[-1]
        z = torch.empty(Bn, S, C_, dtype=torch.float32, device=u.device)
        call("a3t_glu_dwconv_fwd", _p(u), _dt(u), _p(w), _p(bias), _p(z), Bn, S, C_, k, _stream(u))
        return z

    def glu_dwconv_bwd(self, dz, u, w):
        Bn, S, C2 = u.shape
        C_ = C2 // 2
        k = w.shape[-1]
        du = torch.empty_like(u)
        dw = torch.empty(C_, 1, k, dtype=torch.float32, device=u.device)
        db = torch.empty(C_, dtype=torch.float32, device=u.device)
        nblk = call("a3t_dwconv_bwd_blocks", Bn, S)
        partial = torch.empty(nblk * (k + 1) * C_, dtype=torch.float32, device=u.device)
        call("a3t_glu_dwconv_bwd", _p(dz), _p(u), _dt(u), _p(w), _p(du), _dt(du), _p(dw), _p(db), _p(partial), Bn, S,
             C_, k, _stream(u))
        return du, dw, db

    def bn_stats(self, z, running_mean, running_var, nbt, momentum, eps, training):
        C_ = z.shape[-1]
        rows = z.numel() // C_
        mean = torch.empty(C_, dtype=torch.float32, device=z.device)
        rstd = torch.empty(C_, dtype=torch.float32, device=z.device)
        nblk = call("a3t_colsum_blocks", rows)
        partial = torch.empty(nblk * 2 * C_, dtype=torch.float64, device=z.device) if training else None
        call("a3t_bn_stats", _p(z), _p(mean), _p(rstd), _p(running_mean), _p(running_var), _p(nbt), _p(partial), rows,
             C_, momentum, eps, int(training), _stream(z))
        return mean, rstd

    def bn_act_fwd(self, z, mean, rstd, gamma, beta, act, *, drop=None, residual=None, out_dtype=None):
        C_ = z.shape[-1]
        rows = z.numel() // C_
        y = torch.empty(z.shape, dtype=out_dtype or self.act_dtype, device=z.device)
        p, seed, site = self._drop(drop)
        call("a3t_bn_act_fwd", _p(z), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(residual), _p(y), _dt(y), rows, C_,
             act, p, seed, site, _stream(z))
        return y

    def bn_act_bwd(self, dy, z, mean, rstd, gamma, beta, act, training, *, drop=None, eps=None):
        C_ = z.shape[-1]
        rows = z.numel() // C_
        assert dy.dtype == torch.float32 and dy.is_contiguous()
        dz = torch.empty_like(z)
        dgamma = torch.empty(C_, dtype=torch.float32, device=z.device)
        dbeta = torch.empty(C_, dtype=torch.float32, device=z.device)
        nblk = call("a3t_colsum_blocks", rows)
        partial = torch.empty(nblk * 2 * C_, dtype=torch.float64, device=z.device)
        coef = torch.empty(2 * C_, dtype=torch.float32, device=z.device)
        p, seed, site = self._drop(drop)
        call("a3t_bn_act_bwd", _p(dy), _p(z), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(dz), _p(dgamma), _p(dbeta),
             _p(partial), _p(coef), rows, C_, act, int(training), p, seed, site, _stream(z))
        return dz, dgamma, dbeta

    # ------------------------------------------------------------------ loss
    def masked_l1_fwd(self, before, after, y, mask):
        C_ = y.shape[-1]
        rows = y.numel() // C_
        out = torch.empty(2, dtype=torch.float32, device=y.device)
        nblk = call("a3t_colsum_blocks", rows)
        partial = torch.empty(nblk * 2, dtype=torch.float64, device=y.device)
        call("a3t_masked_l1_fwd", _p(before), _p(after), _p(y), _p(_u8(mask)), _p(out), _p(partial), rows, C_,
             _stream(y))
        return out[0:1], out[1:2]

    def masked_l1_bwd(self, gloss, before, after, y, mask, loss_den):
        C_ = y.shape[-1]
        rows = y.numel() // C_
        dbefore = torch.empty_like(before)
        dafter = torch.empty_like(after) if after is not None else None
        call("a3t_masked_l1_bwd", _p(gloss.reshape(1).contiguous()), _p(before), _p(after), _p(y), _p(_u8(mask)), _p(loss_den),
             _p(dbefore), _p(dafter), rows, C_, _stream(y))
        return dbefore, dafter
