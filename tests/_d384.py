"""Loader of tests/golden/model_d384.pt: the paper-width (D=384, H=2, FF=1536, dw 7/31, postnet 5x256, 1+1 blocks)
fixture produced by running the REFERENCE on B=2, Ts=1024, Tt=128 ragged inputs (oracle/make_golden.py
::model_d384_fixture).  Weights and the speech tensor are regenerated from their seeds."""
import os

import torch

from oracle.fixtures import fill_params


def load(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "model_d384.pt"), weights_only=False)
    b = dict(fx["batch_small"])
    B, Ts = b["masked_position"].shape
    torch.manual_seed(fx["batch_seed"])
    speech = torch.randn(B, Ts, 80)
    for i, n in enumerate(fx["speech_lengths"].tolist()):
        speech[i, n:] = 0.0
    assert torch.equal(speech[:, ::97, ::7], fx["speech_probe"]), "speech regeneration drifted (torch RNG changed?)"
    b["speech"] = speech
    return fx, b


def build(fx, act_dtype=torch.float32, dropout_zero=True):
    from a3t_b200.model import build_model

    conf = fx["conf"]
    enc, dec = dict(conf["encoder_conf"]), dict(conf["decoder_conf"])
    if dropout_zero:
        for c in (enc, dec):
            c.update(dropout_rate=0.0, positional_dropout_rate=0.0, attention_dropout_rate=0.0)
    m = build_model(enc, dec, conf["model_conf"], vocab_size=fx["vocab"], act_dtype=act_dtype, init=None)
    fill_params(m, fx["weight_seed"])
    if dropout_zero:
        m.postnet.dropout_rate = 0.0
    return m


def pwg30_state_dict(f):
    """Regenerate the reference generator's weights: same names / shapes / sorted order as fill_params saw."""
    class _Holder(torch.nn.Module):
        def __init__(self, shapes):
            super().__init__()
            self.names = list(shapes)
            for i, (k, s) in enumerate(shapes.items()):
                self.register_parameter(f"p{i}", torch.nn.Parameter(torch.zeros(s)))

        def named_parameters(self, *a, **k):  # the reference's own names decide the visiting order
            return [(n, getattr(self, f"p{i}")) for i, n in enumerate(self.names)]

    h = _Holder(f["param_shapes"])
    fill_params(h, f["weight_seed"], scale=f["weight_scale"])
    return {n: p.detach().clone() for n, p in h.named_parameters()}


def pwg30_z(f):
    g = torch.Generator().manual_seed(f["z_seed"])
    c = torch.randn(1, 80, 200, generator=g)
    assert torch.equal(c, f["c"])
    z = torch.randn(1, 1, 200 * 300, generator=g)
    assert torch.equal(z[0, 0, ::601], f["z_probe"])
    return z
