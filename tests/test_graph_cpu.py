"""Host logic of the hand-written forward/backward graph (a3t_b200/graph.py) checked on the CPU:
the graph is run over the oracle ops and compared with fixtures produced by the reference model
(loss, outputs, every parameter gradient, BatchNorm running statistics, inference stitching)."""
import os

import pytest
import torch

from a3t_b200 import graph
from a3t_b200.model import build_model
from oracle.oracle_backend import OracleBackend


@pytest.fixture(scope="module")
def fx(golden_dir):
    return torch.load(os.path.join(golden_dir, "model_tiny.pt"), weights_only=False)


def _model(fx, dropout_zero=True):
    conf = fx["conf"]
    enc, dec = dict(conf["encoder_conf"]), dict(conf["decoder_conf"])
    if dropout_zero:
        for c in (enc, dec):
            c.update(dropout_rate=0.0, positional_dropout_rate=0.0, attention_dropout_rate=0.0)
    m = build_model(enc, dec, conf["model_conf"], vocab_size=fx["vocab"])
    m.load_state_dict(fx["state_dict"], strict=True)
    if dropout_zero:
        m.postnet.dropout_rate = 0.0
    return m


def _P(m):
    P = {n: p.detach().clone() for n, p in m.named_parameters()}
    P.update({n: b.clone() for n, b in m.named_buffers()})
    return P


def test_state_dict_layout_matches_reference(fx):
    m = _model(fx)
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    ref = {k: tuple(v.shape) for k, v in fx["state_dict"].items()}
    assert mine == ref


def test_train_step_matches_reference(fx):
    m = _model(fx)
    P, ops, wc = _P(m), OracleBackend(), graph.WeightCache()
    loss, before, after, ctx = graph.forward(ops, P, wc, m.cfg, fx["batch"], training=True)
    assert abs(float(loss) - float(fx["loss_train"])) < 1e-4
    G = graph.backward(ops, P, wc, m.cfg, ctx, torch.ones(1))
    assert set(G) == set(fx["grads"])
    for n, g in fx["grads"].items():
        tol = 2e-4 * float(g.abs().max()) + 2e-5
        assert float((G[n] - g).abs().max()) <= tol, n
    for k, v in fx["bn_after"].items():
        assert torch.allclose(P[k].float(), v.float(), atol=1e-5), k


def test_eval_and_inference_match_reference(fx):
    m = _model(fx)
    P, ops, wc = _P(m), OracleBackend(), graph.WeightCache()
    loss, before, after, _ = graph.forward(ops, P, wc, m.cfg, fx["batch"], training=False)
    assert abs(float(loss) - float(fx["loss_eval"])) < 1e-4
    assert torch.allclose(before, fx["before_eval"], atol=1e-4)
    assert torch.allclose(after, fx["after_eval"], atol=1e-4)
    b1 = {k: v[:1] for k, v in fx["batch"].items()}
    _, _, after1, _ = graph.forward(ops, P, wc, m.cfg, b1, training=False, need_loss=False)
    s, e = 20, 41
    want = fx["inference"]
    assert torch.equal(b1["speech"][:, :s], want[0]) and torch.equal(b1["speech"][:, e:], want[2])
    assert torch.allclose(after1[0][s:e], want[1], atol=1e-4)


def test_dropout_backward_is_consistent():
    """With dropout on, the analytic backward must equal autograd through the same masks:
    finite-difference-free check = compare against autograd of the oracle forward graph."""
    torch.manual_seed(0)
    conf = dict(num_blocks=1, attention_dim=16, attention_heads=2, linear_units=32, input_layer="sega_mlm",
                dropout_rate=0.3, positional_dropout_rate=0.3, attention_dropout_rate=0.3, normalize_before=True,
                macaron_style=True, use_cnn_module=True, selfattention_layer_type="rel_selfattn",
                activation_type="swish", pos_enc_layer_type="rel_pos", positionwise_layer_type="conv1d",
                positionwise_conv_kernel_size=3, cnn_module_kernel=5)
    dconf = {k: v for k, v in conf.items() if k != "input_layer"}
    m = build_model(conf, dconf, dict(postnet_layers=2, postnet_filts=5, postnet_chans=8), vocab_size=11, init=None)
    B, Ts, Tt = 2, 12, 4
    batch = dict(speech=torch.randn(B, Ts, 80), text=torch.randint(1, 9, (B, Tt)),
                 masked_position=torch.rand(B, Ts) < 0.6, speech_mask=torch.ones(B, 1, Ts, dtype=torch.bool),
                 text_mask=torch.ones(B, 1, Tt, dtype=torch.bool),
                 speech_segment_pos=torch.randint(0, 5, (B, Ts)), text_segment_pos=torch.randint(0, 5, (B, Tt)))
    P, ops, wc = _P(m), OracleBackend(seed=99), graph.WeightCache()
    loss, _, _, ctx = graph.forward(ops, P, wc, m.cfg, batch, training=True)
    G = graph.backward(ops, P, wc, m.cfg, ctx, torch.ones(1))
    # central differences on a few scalar parameters
    for name in ("sfc.bias", "encoder.encoders.0.norm_mha.weight", "decoder.encoders.0.feed_forward.w_1.bias",
                 "encoder.encoders.0.self_attn.pos_bias_u", "postnet.postnet.0.1.weight"):
        p = P[name]
        idx = tuple(0 for _ in p.shape)
        orig = float(p[idx])
        vals = []
        for d in (1e-2, -1e-2):
            P2 = {k: v.clone() for k, v in _P(m).items()}
            P2[name][idx] = orig + d
            l2, _, _, _ = graph.forward(ops, P2, graph.WeightCache(), m.cfg, batch, training=True)
            vals.append(float(l2))
        fd = (vals[0] - vals[1]) / 2e-2
        assert abs(fd - float(G[name][idx])) <= 5e-2 * max(abs(fd), 1.0), (name, fd, float(G[name][idx]))
