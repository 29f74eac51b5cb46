"""Forward and hand-written backward of the A3T masked-mel model as a static op graph.

The graph is written against an `ops` object (`a3t_b200.backend.CudaBackend` in the product;
the tests substitute the CPU oracle to check this host logic without a GPU).  It follows
`ESPnetMLMEncAsDecoderModel._forward` (espnet2/tts/sedit/sedit_model.py:350-375),
`MLMEncoder.forward` / `MLMDecoder.forward` (espnet/nets/pytorch_backend/conformer/encoder.py:522-614),
`EncoderLayer.forward` (conformer/encoder_layer.py:80-180), `Postnet.forward`
(tacotron2/decoder.py:254-268) and `_calc_mlm_loss` (sedit_model.py:320-340).

Parameters are addressed by the reference's own state_dict names.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

ACT_NONE, ACT_SWISH, ACT_TANH = 0, 1, 2


@dataclass
class A3TConfig:
    idim: int = 80
    odim: int = 80
    vocab_size: int = 73
    D: int = 384
    H: int = 2
    FF: int = 1536
    ffn_kernel: int = 3
    enc_blocks: int = 4
    dec_blocks: int = 4
    enc_dw_kernel: int = 7
    dec_dw_kernel: int = 31
    dropout: float = 0.2
    pos_dropout: float = 0.2
    att_dropout: float = 0.2
    dec_dropout: float = 0.2
    dec_pos_dropout: float = 0.2
    dec_att_dropout: float = 0.2
    postnet_layers: int = 5
    postnet_chans: int = 256
    postnet_filts: int = 5
    postnet_dropout: float = 0.5
    n_segments: int = 500
    sega: bool = True
    max_len: int = 5000


_POS_CACHE: Dict[tuple, torch.Tensor] = {}


def legacy_rel_pos_table(T: int, D: int, device, max_len: int = 5000) -> torch.Tensor:
    """Rows [:T] of the reversed sinusoid table (transformer/embedding.py:56-80,147-170):
    row t encodes position max(T, max_len)-1-t.  Built in fp32 exactly as the reference does."""
    key = (T, D, str(device), max_len)
    t = _POS_CACHE.get(key)
    if t is None:
        L = max(T, max_len)
        position = torch.arange(L - 1, -1, -1.0, dtype=torch.float32).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, D, 2, dtype=torch.float32) * -(math.log(10000.0) / D))
        pe = torch.zeros(L, D)
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        t = pe[:T].contiguous().to(device)
        _POS_CACHE[key] = t
    return t


class WeightCache:
    """Packed GEMM weights keyed by parameter name; refreshed when the parameter changes."""

    def __init__(self):
        self._c = {}
        # bumped whenever an entry is (re)built or the cache is dropped: holders of derived state (the trainer's
        # in-place repack plan) compare it to know their pointers are stale
        self.generation = 0

    def packed(self, ops, key, tensors, make):
        sig = tuple((t.data_ptr(), t._version) for t in tensors) + (str(ops.act_dtype),)
        hit = self._c.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        if make is None:
            raise RuntimeError(f"weight cache miss for {key} during backward (parameter changed since forward?)")
        val = make()
        self._c[key] = (sig, val)
        self.generation += 1
        return val

    def clear(self):
        self._c.clear()
        self.generation += 1


class _Sites:
    """Deterministic numbering of the dropout sites of one step."""

    def __init__(self):
        self.n = 0

    def next(self):
        self.n += 1
        return self.n


def _drop(p, site, training):
    return (p, site) if (training and p > 0.0) else None


# ----------------------------------------------------------------------------------------------
# one Conformer block
# ----------------------------------------------------------------------------------------------

def _ffn_fwd(ops, P, wc, pre, norm, x, p_drop, sites, training, ctx):
    g, b = P[f"{pre}.{norm}.weight"], P[f"{pre}.{norm}.bias"]
    ff = "feed_forward_macaron" if norm == "norm_ff_macaron" else "feed_forward"
    w1 = wc.packed(ops, f"{pre}.{ff}.w_1", [P[f"{pre}.{ff}.w_1.weight"]], lambda: ops.pack_weight(P[f"{pre}.{ff}.w_1.weight"]))
    w2 = wc.packed(ops, f"{pre}.{ff}.w_2", [P[f"{pre}.{ff}.w_2.weight"]], lambda: ops.pack_weight(P[f"{pre}.{ff}.w_2.weight"]))
    s1, s2 = sites.next(), sites.next()
    h, mean, rstd = ops.ln_fwd(x, g, b, 1e-12)
    u = ops.conv_fwd(h, w1, P[f"{pre}.{ff}.w_1.bias"], relu=True, drop=_drop(p_drop, s1, training))
    y = ops.conv_fwd(u, w2, P[f"{pre}.{ff}.w_2.bias"], drop=_drop(p_drop, s2, training), residual=x, out_scale=0.5)
    ctx.update({f"{ff}.x": x, f"{ff}.mean": mean, f"{ff}.rstd": rstd, f"{ff}.h": h, f"{ff}.u": u, f"{ff}.s1": s1,
                f"{ff}.s2": s2})
    return y


class _GradOut(dict):
    """Parameter-gradient dictionary.  `dest` (optional) maps parameter names to caller-owned, ZEROED fp32
    buffers (the trainer's flat gradient views): weight-gradient GEMMs write straight into them."""

    def __init__(self, P, dest=None):
        super().__init__()
        self.P = P
        self.dest = dest or {}

    def wgrad(self, ops, name, dy, x, taps):
        """dW of a conv / linear layer from its output gradient dy and input x, in the parameter's own shape."""
        out = self.dest.get(name)
        if out is not None:
            r = ops.conv_wgrad(dy, x, taps, out=out, out_zeroed=True)
        else:
            r = ops.conv_wgrad(dy, x, taps)
        self[name] = r.view(self.P[name].shape)
        return self[name]


def _grad_prep(ops, dy, scale, drop, gpre):
    """g = dropout'(dy * scale) in the GEMM dtype and its column sums: taken from the LayerNorm backward that
    produced dy when it fused them (`gpre`), else computed here."""
    if gpre is not None:
        return gpre
    g = ops.scale_dropout(dy, scale, drop)
    return g, ops.colsum(g)


def _ln_bwd_next(ops, nxt, *args, **kw):
    """LayerNorm backward that also emits the NEXT backward section's grad prep when `nxt` = (scale, drop)."""
    out = ops.ln_bwd(*args, nxt=nxt, **kw)
    if nxt is None:
        return out[0], out[1], out[2], None
    return out[0], out[1], out[2], (out[3], out[4])


def _ffn_bwd(ops, P, wc, pre, norm, dy, p_drop, training, ctx, G, gpre=None, nxt=None):
    ff = "feed_forward_macaron" if norm == "norm_ff_macaron" else "feed_forward"
    w1 = wc.packed(ops, f"{pre}.{ff}.w_1", [P[f"{pre}.{ff}.w_1.weight"]], None)
    w2 = wc.packed(ops, f"{pre}.{ff}.w_2", [P[f"{pre}.{ff}.w_2.weight"]], None)
    x, h, u = ctx[f"{ff}.x"], ctx[f"{ff}.h"], ctx[f"{ff}.u"]
    d1, d2 = _drop(p_drop, ctx[f"{ff}.s1"], training), _drop(p_drop, ctx[f"{ff}.s2"], training)
    g, gsum = _grad_prep(ops, dy, 0.5, d2, gpre)
    G[f"{pre}.{ff}.w_2.bias"] = gsum
    G.wgrad(ops, f"{pre}.{ff}.w_2.weight", g, u, w2.taps)
    inv_keep = 1.0 / (1.0 - p_drop) if d1 is not None else 1.0
    du = ops.conv_dgrad(g, w2, mask=u, mask_scale=inv_keep)
    G[f"{pre}.{ff}.w_1.bias"] = ops.colsum(du)
    G.wgrad(ops, f"{pre}.{ff}.w_1.weight", du, h, w1.taps)
    dh = ops.conv_dgrad(du, w1)
    dx, dg, db, gn = _ln_bwd_next(ops, nxt, dh, x, ctx[f"{ff}.mean"], ctx[f"{ff}.rstd"], P[f"{pre}.{norm}.weight"],
                                  P[f"{pre}.{norm}.bias"], dres=dy, eps=1e-12)
    G[f"{pre}.{norm}.weight"], G[f"{pre}.{norm}.bias"] = dg, db
    return dx, gn


def _qkv4_weight(P, pre):
    a = f"{pre}.self_attn"
    wq, wk, wv = P[f"{a}.linear_q.weight"], P[f"{a}.linear_k.weight"], P[f"{a}.linear_v.weight"]
    bq, bk, bv = P[f"{a}.linear_q.bias"], P[f"{a}.linear_k.bias"], P[f"{a}.linear_v.bias"]
    u, v = P[f"{a}.pos_bias_u"].reshape(-1), P[f"{a}.pos_bias_v"].reshape(-1)
    w4 = torch.cat([wq, wq, wk, wv], dim=0).contiguous()
    b4 = torch.cat([bq + u, bq + v, bk, bv], dim=0).contiguous()
    return w4, b4


def _mha_fwd(ops, P, wc, pre, x, pos_d, keymask, cfg_H, p_drop, p_att, sites, training, ctx):
    a = f"{pre}.self_attn"
    D = x.shape[-1]
    names = [f"{a}.linear_q.weight", f"{a}.linear_k.weight", f"{a}.linear_v.weight", f"{a}.linear_q.bias",
             f"{a}.linear_k.bias", f"{a}.linear_v.bias", f"{a}.pos_bias_u", f"{a}.pos_bias_v"]

    def make():
        w4, b4 = _qkv4_weight(P, pre)
        return ops.pack_weight(w4), b4

    w4, b4 = wc.packed(ops, f"{a}.qkv4", [P[n] for n in names], make)
    wpos = wc.packed(ops, f"{a}.linear_pos", [P[f"{a}.linear_pos.weight"]], lambda: ops.pack_weight(P[f"{a}.linear_pos.weight"]))
    wo = wc.packed(ops, f"{a}.linear_out", [P[f"{a}.linear_out.weight"]], lambda: ops.pack_weight(P[f"{a}.linear_out.weight"]))
    s_att, s_out = sites.next(), sites.next()
    h, mean, rstd = ops.ln_fwd(x, P[f"{pre}.norm_mha.weight"], P[f"{pre}.norm_mha.bias"], 1e-12)
    qkv4 = ops.conv_fwd(h, w4, b4)
    pp = ops.conv_fwd(pos_d.unsqueeze(0), wpos, None).squeeze(0)
    scale = 1.0 / math.sqrt(D // cfg_H)
    fused = hasattr(ops, "attn_fused_ok") and ops.attn_fused_ok(x.shape[0], cfg_H, x.shape[1], D)
    if fused:
        # tcgen05 kernels that keep the scores on chip; BD_raw and the row log-sum-exp are what the backward needs
        cx, bd, lse = ops.attn_fwd_fused(qkv4, pp, keymask, cfg_H, scale, drop=_drop(p_att, s_att, training))
        ctx.update({"mha.bd": bd, "mha.lse": lse, "mha.keymask": keymask})
    else:
        ac, bd = ops.attn_scores_fwd(qkv4, pp, cfg_H)
        Pm, Pd = ops.relpos_softmax_fwd(ac, bd, keymask, scale, drop=_drop(p_att, s_att, training))
        del ac, bd
        cx = ops.attn_pv_fwd(Pd, qkv4, cfg_H)
        ctx.update({"mha.P": Pm, "mha.Pd": Pd})
    y = ops.conv_fwd(cx, wo, P[f"{a}.linear_out.bias"], drop=_drop(p_drop, s_out, training), residual=x)
    ctx.update({"mha.x": x, "mha.mean": mean, "mha.rstd": rstd, "mha.h": h, "mha.qkv4": qkv4, "mha.pp": pp,
                "mha.cx": cx, "mha.s_att": s_att, "mha.s_out": s_out, "mha.fused": fused})
    return y


def _mha_bwd(ops, P, wc, pre, dy, pos_d, cfg_H, p_drop, p_att, training, ctx, G, gpre=None, nxt=None):
    a = f"{pre}.self_attn"
    x, h, qkv4, pp = ctx["mha.x"], ctx["mha.h"], ctx["mha.qkv4"], ctx["mha.pp"]
    cx = ctx["mha.cx"]
    D = x.shape[-1]
    H = cfg_H
    w4, _ = wc.packed(ops, f"{a}.qkv4", [P[n] for n in [f"{a}.linear_q.weight", f"{a}.linear_k.weight",
                                                        f"{a}.linear_v.weight", f"{a}.linear_q.bias",
                                                        f"{a}.linear_k.bias", f"{a}.linear_v.bias",
                                                        f"{a}.pos_bias_u", f"{a}.pos_bias_v"]], None)
    wpos = wc.packed(ops, f"{a}.linear_pos", [P[f"{a}.linear_pos.weight"]], None)
    wo = wc.packed(ops, f"{a}.linear_out", [P[f"{a}.linear_out.weight"]], None)
    g, gsum = _grad_prep(ops, dy, 1.0, _drop(p_drop, ctx["mha.s_out"], training), gpre)
    G[f"{a}.linear_out.bias"] = gsum
    G.wgrad(ops, f"{a}.linear_out.weight", g, cx, 1)
    dcx = ops.conv_dgrad(g, wo)
    dqkv4 = torch.empty_like(qkv4)
    if ctx["mha.fused"]:
        dpp = ops.attn_bwd_fused(dcx, cx, ctx["mha.lse"], ctx["mha.bd"], qkv4, pp, ctx["mha.keymask"], H,
                                 1.0 / math.sqrt(D // H), dqkv4, drop=_drop(p_att, ctx["mha.s_att"], training))
    else:
        Pm, Pd = ctx["mha.P"], ctx["mha.Pd"]
        dPd = ops.attn_pv_bwd(dcx, Pd, qkv4, H, dqkv4)
        dS, dBD = ops.relpos_softmax_bwd(dPd, Pm, 1.0 / math.sqrt(D // H), drop=_drop(p_att, ctx["mha.s_att"], training))
        del dPd, Pm, Pd
        dpp = ops.attn_scores_bwd(dS, dBD, qkv4, pp, H, dqkv4)
        del dS, dBD
    G.wgrad(ops, f"{a}.linear_pos.weight", ops.cast_act(dpp).unsqueeze(0), pos_d.unsqueeze(0), 1)
    db4 = ops.colsum(dqkv4)
    dw4 = ops.conv_wgrad(dqkv4, h, 1).squeeze(-1)
    G[f"{a}.linear_q.weight"] = dw4[0:D] + dw4[D:2 * D]
    G[f"{a}.linear_k.weight"] = dw4[2 * D:3 * D]
    G[f"{a}.linear_v.weight"] = dw4[3 * D:4 * D]
    G[f"{a}.linear_q.bias"] = db4[0:D] + db4[D:2 * D]
    G[f"{a}.linear_k.bias"] = db4[2 * D:3 * D]
    G[f"{a}.linear_v.bias"] = db4[3 * D:4 * D]
    G[f"{a}.pos_bias_u"] = db4[0:D].reshape(H, D // H)
    G[f"{a}.pos_bias_v"] = db4[D:2 * D].reshape(H, D // H)
    dh = ops.conv_dgrad(dqkv4, w4)
    dx, dg, db, gn = _ln_bwd_next(ops, nxt, dh, x, ctx["mha.mean"], ctx["mha.rstd"], P[f"{pre}.norm_mha.weight"],
                                  P[f"{pre}.norm_mha.bias"], dres=dy, eps=1e-12)
    G[f"{pre}.norm_mha.weight"], G[f"{pre}.norm_mha.bias"] = dg, db
    return dx, gn


def _convmod_fwd(ops, P, wc, pre, x, p_drop, sites, training, ctx):
    c = f"{pre}.conv_module"
    w1 = wc.packed(ops, f"{c}.pw1", [P[f"{c}.pointwise_conv1.weight"]], lambda: ops.pack_weight(P[f"{c}.pointwise_conv1.weight"]))
    w2 = wc.packed(ops, f"{c}.pw2", [P[f"{c}.pointwise_conv2.weight"]], lambda: ops.pack_weight(P[f"{c}.pointwise_conv2.weight"]))
    s = sites.next()
    h, mean, rstd = ops.ln_fwd(x, P[f"{pre}.norm_conv.weight"], P[f"{pre}.norm_conv.bias"], 1e-12)
    u = ops.conv_fwd(h, w1, P[f"{c}.pointwise_conv1.bias"])
    wdw = P[f"{c}.depthwise_conv.weight"]
    z = ops.glu_dwconv_fwd(u, wdw.reshape(wdw.shape[0], wdw.shape[-1]), P[f"{c}.depthwise_conv.bias"])
    bm, br = ops.bn_stats(z, P[f"{c}.norm.running_mean"], P[f"{c}.norm.running_var"],
                          P[f"{c}.norm.num_batches_tracked"], 0.1, 1e-5, training)
    act = ops.bn_act_fwd(z, bm, br, P[f"{c}.norm.weight"], P[f"{c}.norm.bias"], ACT_SWISH)
    y = ops.conv_fwd(act, w2, P[f"{c}.pointwise_conv2.bias"], drop=_drop(p_drop, s, training), residual=x)
    ctx.update({"cm.x": x, "cm.mean": mean, "cm.rstd": rstd, "cm.h": h, "cm.u": u, "cm.z": z, "cm.bm": bm, "cm.br": br,
                "cm.act": act, "cm.s": s})
    return y


def _convmod_bwd(ops, P, wc, pre, dy, p_drop, training, ctx, G, gpre=None, nxt=None):
    c = f"{pre}.conv_module"
    w1 = wc.packed(ops, f"{c}.pw1", [P[f"{c}.pointwise_conv1.weight"]], None)
    w2 = wc.packed(ops, f"{c}.pw2", [P[f"{c}.pointwise_conv2.weight"]], None)
    x, h, u, z, act = ctx["cm.x"], ctx["cm.h"], ctx["cm.u"], ctx["cm.z"], ctx["cm.act"]
    g, gsum = _grad_prep(ops, dy, 1.0, _drop(p_drop, ctx["cm.s"], training), gpre)
    G[f"{c}.pointwise_conv2.bias"] = gsum
    G.wgrad(ops, f"{c}.pointwise_conv2.weight", g, act, 1)
    dact = ops.conv_dgrad(g, w2, out_dtype=torch.float32)
    dz, dgam, dbet = ops.bn_act_bwd(dact, z, ctx["cm.bm"], ctx["cm.br"], P[f"{c}.norm.weight"], P[f"{c}.norm.bias"],
                                    ACT_SWISH, training, eps=1e-5)
    G[f"{c}.norm.weight"], G[f"{c}.norm.bias"] = dgam, dbet
    wdw = P[f"{c}.depthwise_conv.weight"]
    du, dw, db = ops.glu_dwconv_bwd(dz, u, wdw.reshape(wdw.shape[0], wdw.shape[-1]))
    G[f"{c}.depthwise_conv.weight"], G[f"{c}.depthwise_conv.bias"] = dw.reshape(wdw.shape), db
    G[f"{c}.pointwise_conv1.bias"] = ops.colsum(du)
    G.wgrad(ops, f"{c}.pointwise_conv1.weight", du, h, 1)
    dh = ops.conv_dgrad(du, w1)
    dx, dg, dbb, gn = _ln_bwd_next(ops, nxt, dh, x, ctx["cm.mean"], ctx["cm.rstd"], P[f"{pre}.norm_conv.weight"],
                                   P[f"{pre}.norm_conv.bias"], dres=dy, eps=1e-12)
    G[f"{pre}.norm_conv.weight"], G[f"{pre}.norm_conv.bias"] = dg, dbb
    return dx, gn


def _layer_fwd(ops, P, wc, pre, x, pos_d, keymask, H, p_drop, p_att, sites, training, ctx):
    x = _ffn_fwd(ops, P, wc, pre, "norm_ff_macaron", x, p_drop, sites, training, ctx)
    x = _mha_fwd(ops, P, wc, pre, x, pos_d, keymask, H, p_drop, p_att, sites, training, ctx)
    x = _convmod_fwd(ops, P, wc, pre, x, p_drop, sites, training, ctx)
    x = _ffn_fwd(ops, P, wc, pre, "norm_ff", x, p_drop, sites, training, ctx)
    y, mean, rstd = ops.ln_fwd(x, P[f"{pre}.norm_final.weight"], P[f"{pre}.norm_final.bias"], 1e-12,
                               out_dtype=torch.float32)
    ctx.update({"fin.x": x, "fin.mean": mean, "fin.rstd": rstd})
    return y


def _layer_bwd(ops, P, wc, pre, dy, pos_d, H, p_drop, p_att, training, ctx, G):
    # every LayerNorm backward also produces the grad prep (scaled / dropped bf16 copy + bias-gradient column
    # sums) of the section that consumes its dx next, so that section starts directly with its GEMMs
    nxt = lambda scale, site: (scale, _drop(p_drop, site, training))
    dx, dg, db, g = _ln_bwd_next(ops, nxt(0.5, ctx["feed_forward.s2"]), dy, ctx["fin.x"], ctx["fin.mean"],
                                 ctx["fin.rstd"], P[f"{pre}.norm_final.weight"], P[f"{pre}.norm_final.bias"], eps=1e-12)
    G[f"{pre}.norm_final.weight"], G[f"{pre}.norm_final.bias"] = dg, db
    dx, g = _ffn_bwd(ops, P, wc, pre, "norm_ff", dx, p_drop, training, ctx, G, gpre=g, nxt=nxt(1.0, ctx["cm.s"]))
    dx, g = _convmod_bwd(ops, P, wc, pre, dx, p_drop, training, ctx, G, gpre=g, nxt=nxt(1.0, ctx["mha.s_out"]))
    dx, g = _mha_bwd(ops, P, wc, pre, dx, pos_d, H, p_drop, p_att, training, ctx, G, gpre=g,
                     nxt=nxt(0.5, ctx["feed_forward_macaron.s2"]))
    dx, _ = _ffn_bwd(ops, P, wc, pre, "norm_ff_macaron", dx, p_drop, training, ctx, G, gpre=g)
    return dx


# ----------------------------------------------------------------------------------------------
# whole model
# ----------------------------------------------------------------------------------------------

@dataclass
class StepContext:
    saved: dict = field(default_factory=dict)
    layers: List[dict] = field(default_factory=list)
    training: bool = True


def forward(ops, P: Dict[str, torch.Tensor], wc: WeightCache, cfg: A3TConfig, batch: dict, training: bool,
            need_loss: bool = True):
    """Returns (loss[1] or None, before (B,Ts,odim), after (B,Ts,odim), ctx)."""
    speech = batch["speech"].contiguous()
    text = batch["text"].long().contiguous()   # the embedding kernels read int64 ids
    masked = batch["masked_position"].contiguous()
    Bn, Ts, _ = speech.shape
    Tt = text.shape[1]
    S = Ts + Tt
    D = cfg.D
    dev = speech.device
    keymask = torch.cat([batch["speech_mask"].reshape(Bn, Ts), batch["text_mask"].reshape(Bn, Tt)], dim=1).contiguous()
    sseg = batch["speech_segment_pos"].long().contiguous() if cfg.sega else None
    tseg = batch["text_segment_pos"].long().contiguous() if cfg.sega else None
    xscale = math.sqrt(D)
    sites = _Sites()
    ctx = StepContext(training=training)
    sv = ctx.saved

    # ---- encoder embedding (conformer/encoder.py:526-553)
    w_in = wc.packed(ops, "enc.prenet", [P["encoder.speech_embed.1.weight"]],
                     lambda: ops.pack_weight(P["encoder.speech_embed.1.weight"]))
    xm = ops.mask_input_fwd(speech, masked, P["encoder.speech_embed.0.mask_feature"].reshape(-1))
    h0 = ops.conv_fwd(xm, w_in, P["encoder.speech_embed.1.bias"], out_dtype=torch.float32)
    sy, mean0, rstd0 = ops.ln_fwd(h0, P["encoder.speech_embed.2.weight"], P["encoder.speech_embed.2.bias"], 1e-5,
                                  relu=True, out_scale=xscale, out_dtype=torch.float32)
    s_sp, s_tx, s_pos = sites.next(), sites.next(), sites.next()
    xs = ops.embed_assemble_fwd(sy, text, sseg, tseg, P["encoder.text_embed.0.weight"],
                                P["encoder.segment_emb.weight"] if cfg.sega else None, xscale,
                                drop_speech=_drop(cfg.pos_dropout, s_sp, training),
                                drop_text=_drop(cfg.pos_dropout, s_tx, training))
    pos = torch.cat([legacy_rel_pos_table(Ts, D, dev, cfg.max_len), legacy_rel_pos_table(Tt, D, dev, cfg.max_len)], 0)
    pos_e = ops.scale_dropout(pos.contiguous(), 1.0, _drop(cfg.pos_dropout, s_pos, training))
    sv.update(dict(xm=xm, h0=h0, mean0=mean0, rstd0=rstd0, s_sp=s_sp, s_tx=s_tx, pos_e=pos_e, keymask=keymask,
                   text=text, sseg=sseg, tseg=tseg, masked=masked, speech=speech, Ts=Ts, Tt=Tt))

    x = xs
    for l in range(cfg.enc_blocks):
        lc = {}
        x = _layer_fwd(ops, P, wc, f"encoder.encoders.{l}", x, pos_e, keymask, cfg.H, cfg.dropout, cfg.att_dropout,
                       sites, training, lc)
        ctx.layers.append(lc)
    xe, mean_e, rstd_e = ops.ln_fwd(x, P["encoder.after_norm.weight"], P["encoder.after_norm.bias"], 1e-12,
                                    out_dtype=torch.float32)
    sv.update(dict(enc_x=x, enc_mean=mean_e, enc_rstd=rstd_e))

    # ---- decoder (conformer/encoder.py:570-614): embed = positional encoding only
    s_dx, s_dpos = sites.next(), sites.next()
    xd = ops.scale_dropout(xe, xscale, _drop(cfg.dec_pos_dropout, s_dx, training), out_dtype=torch.float32)
    pos_d = ops.scale_dropout(legacy_rel_pos_table(S, D, dev, cfg.max_len), 1.0,
                              _drop(cfg.dec_pos_dropout, s_dpos, training))
    sv.update(dict(s_dx=s_dx, pos_d=pos_d))
    x = xd
    for l in range(cfg.dec_blocks):
        lc = {}
        x = _layer_fwd(ops, P, wc, f"decoder.encoders.{l}", x, pos_d, keymask, cfg.H, cfg.dec_dropout,
                       cfg.dec_att_dropout, sites, training, lc)
        ctx.layers.append(lc)
    z, mean_d, rstd_d = ops.ln_fwd(x, P["decoder.after_norm.weight"], P["decoder.after_norm.bias"], 1e-12)
    sv.update(dict(dec_x=x, dec_mean=mean_d, dec_rstd=rstd_d))

    # ---- head + postnet (sedit_model.py:363-372, tacotron2/decoder.py:254-268)
    zs = z[:, :Ts].contiguous()
    w_sfc = wc.packed(ops, "sfc", [P["sfc.weight"]], lambda: ops.pack_weight(P["sfc.weight"]))
    before = ops.conv_fwd(zs, w_sfc, P["sfc.bias"], out_dtype=torch.float32)
    sv.update(dict(zs=zs, before=before))
    after = None
    if cfg.postnet_layers > 0:
        a_in = ops.cast_act(before)
        pn = []
        for i in range(cfg.postnet_layers):
            wi = wc.packed(ops, f"postnet.{i}", [P[f"postnet.postnet.{i}.0.weight"]],
                           lambda i=i: ops.pack_weight(P[f"postnet.postnet.{i}.0.weight"]))
            zi = ops.conv_fwd(a_in, wi, None, out_dtype=torch.float32)
            bn = f"postnet.postnet.{i}.1"
            bm, br = ops.bn_stats(zi, P[f"{bn}.running_mean"], P[f"{bn}.running_var"], P[f"{bn}.num_batches_tracked"],
                                  0.1, 1e-5, training)
            last = i == cfg.postnet_layers - 1
            si = sites.next()
            ai = ops.bn_act_fwd(zi, bm, br, P[f"{bn}.weight"], P[f"{bn}.bias"], ACT_NONE if last else ACT_TANH,
                                drop=_drop(cfg.postnet_dropout, si, training), residual=before if last else None,
                                out_dtype=torch.float32 if last else None)
            pn.append(dict(a_in=a_in, z=zi, bm=bm, br=br, s=si))
            a_in = ai
        after = a_in
        sv["pn"] = pn
    sv["after"] = after

    loss = None
    if need_loss:
        loss, den = ops.masked_l1_fwd(before, after, speech, masked)
        sv["den"] = den
    return loss, before, after, ctx


def backward(ops, P, wc: WeightCache, cfg: A3TConfig, ctx: StepContext, gloss: torch.Tensor,
             dbefore_ext=None, dafter_ext=None, gout=None, on_ready=None) -> Dict[str, torch.Tensor]:
    """Gradients of all parameters given d loss (and optionally extra grads on before/after).
    gout: optional {parameter name: zeroed fp32 buffer of the parameter's shape}; weight gradients are then
    written in place (G[name] aliases it) instead of into fresh tensors.
    on_ready: optional callback(G) invoked after every section of the backward sweep (postnet + head, each
    decoder block, each encoder block) with the gradients finished so far -- the data-parallel trainer uses it
    to start the gradient exchange of finished parameter ranges while the sweep continues."""
    sv = ctx.saved
    training = ctx.training
    G = _GradOut(P, gout)
    if hasattr(ops, "begin_backward"):
        ops.begin_backward()  # zeroes the arena the small gradient outputs are accumulated into
    speech, masked = sv["speech"], sv["masked"]
    before, after = sv["before"], sv["after"]
    Ts, Tt = sv["Ts"], sv["Tt"]
    Bn = speech.shape[0]
    D = cfg.D
    xscale = math.sqrt(D)

    dbefore, dafter = ops.masked_l1_bwd(gloss, before, after, speech, masked, sv["den"])
    if dbefore_ext is not None:
        dbefore = dbefore + dbefore_ext
    if after is not None:
        if dafter_ext is not None:
            dafter = dafter + dafter_ext
        d = dafter
        n = cfg.postnet_layers
        for i in reversed(range(n)):
            e = sv["pn"][i]
            bn = f"postnet.postnet.{i}.1"
            last = i == n - 1
            dz, dgam, dbet = ops.bn_act_bwd(d, e["z"], e["bm"], e["br"], P[f"{bn}.weight"], P[f"{bn}.bias"],
                                            ACT_NONE if last else ACT_TANH, training,
                                            drop=_drop(cfg.postnet_dropout, e["s"], training), eps=1e-5)
            G[f"{bn}.weight"], G[f"{bn}.bias"] = dgam, dbet
            wi = wc.packed(ops, f"postnet.{i}", [P[f"postnet.postnet.{i}.0.weight"]], None)
            dza = ops.cast_act(dz)
            G.wgrad(ops, f"postnet.postnet.{i}.0.weight", dza, e["a_in"], wi.taps)
            d = ops.conv_dgrad(dza, wi, out_dtype=torch.float32)
        dbefore = dbefore + dafter + d

    w_sfc = wc.packed(ops, "sfc", [P["sfc.weight"]], None)
    dba = ops.cast_act(dbefore.contiguous())
    G["sfc.bias"] = ops.colsum(dbefore)
    G.wgrad(ops, "sfc.weight", dba, sv["zs"], 1)
    dzs = ops.conv_dgrad(dba, w_sfc)
    dz = torch.zeros(Bn, Ts + Tt, D, dtype=dzs.dtype, device=dzs.device)
    dz[:, :Ts] = dzs
    dx, dg, db = ops.ln_bwd(dz, sv["dec_x"], sv["dec_mean"], sv["dec_rstd"], P["decoder.after_norm.weight"],
                            P["decoder.after_norm.bias"], eps=1e-12)
    G["decoder.after_norm.weight"], G["decoder.after_norm.bias"] = dg, db
    ready = on_ready if on_ready is not None else (lambda g: None)
    ready(G)
    for l in reversed(range(cfg.dec_blocks)):
        dx = _layer_bwd(ops, P, wc, f"decoder.encoders.{l}", dx, sv["pos_d"], cfg.H, cfg.dec_dropout,
                        cfg.dec_att_dropout, training, ctx.layers[cfg.enc_blocks + l], G)
        ctx.layers[cfg.enc_blocks + l] = None
        ready(G)
    dxe = ops.scale_dropout(dx, xscale, _drop(cfg.dec_pos_dropout, sv["s_dx"], training), out_dtype=torch.float32)
    dx, dg, db = ops.ln_bwd(dxe, sv["enc_x"], sv["enc_mean"], sv["enc_rstd"], P["encoder.after_norm.weight"],
                            P["encoder.after_norm.bias"], eps=1e-12)
    G["encoder.after_norm.weight"], G["encoder.after_norm.bias"] = dg, db
    for l in reversed(range(cfg.enc_blocks)):
        dx = _layer_bwd(ops, P, wc, f"encoder.encoders.{l}", dx, sv["pos_e"], cfg.H, cfg.dropout, cfg.att_dropout,
                        training, ctx.layers[l], G)
        ctx.layers[l] = None
        ready(G)

    V = P["encoder.text_embed.0.weight"].shape[0]
    dsy, demb, dseg = ops.embed_assemble_bwd(dx, sv["text"], sv["sseg"], sv["tseg"], V, cfg.n_segments, xscale, V - 1,
                                             cfg.n_segments - 1,
                                             drop_speech=_drop(cfg.pos_dropout, sv["s_sp"], training),
                                             drop_text=_drop(cfg.pos_dropout, sv["s_tx"], training))
    G["encoder.text_embed.0.weight"] = demb
    if dseg is not None:
        G["encoder.segment_emb.weight"] = dseg
    dh0, dg, db = ops.ln_bwd(dsy, sv["h0"], sv["mean0"], sv["rstd0"], P["encoder.speech_embed.2.weight"],
                             P["encoder.speech_embed.2.bias"], relu=True, out_scale=xscale, eps=1e-5)
    G["encoder.speech_embed.2.weight"], G["encoder.speech_embed.2.bias"] = dg, db
    w_in = wc.packed(ops, "enc.prenet", [P["encoder.speech_embed.1.weight"]], None)
    G["encoder.speech_embed.1.bias"] = ops.colsum(dh0)
    dh0a = ops.cast_act(dh0)
    G.wgrad(ops, "encoder.speech_embed.1.weight", dh0a, sv["xm"], 1)
    dxm = ops.conv_dgrad(dh0a, w_in, out_dtype=torch.float32)
    G["encoder.speech_embed.0.mask_feature"] = ops.mask_input_bwd(dxm, masked).reshape(1, 1, -1)
    return G
