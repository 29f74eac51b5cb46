"""The CUDA model (through the nn.Module / autograd surface) against fixtures produced by the
reference model, and the frontend / vocoder / collate kernels against reference outputs (-m gpu)."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import a3t_oracle as O


@pytest.fixture(scope="module")
def fx(golden_dir, cuda_lib):
    assert torch.cuda.is_available()
    return torch.load(os.path.join(golden_dir, "model_tiny.pt"), weights_only=False)


def _model(fx, act_dtype=torch.float32):
    from a3t_b200.model import build_model

    conf = fx["conf"]
    enc, dec = dict(conf["encoder_conf"]), dict(conf["decoder_conf"])
    for c in (enc, dec):
        c.update(dropout_rate=0.0, positional_dropout_rate=0.0, attention_dropout_rate=0.0)
    m = build_model(enc, dec, conf["model_conf"], vocab_size=fx["vocab"], act_dtype=act_dtype)
    m.load_state_dict(fx["state_dict"], strict=True)
    m.postnet.dropout_rate = 0.0
    return m.cuda()


def _cuda(b):
    return {k: v.cuda() for k, v in b.items()}


def test_train_step_fp32_matches_reference(fx):
    m = _model(fx).train()
    loss, stats, weight = m(**_cuda(fx["batch"]))
    assert abs(float(loss) - float(fx["loss_train"])) < 1e-4 * max(1.0, abs(float(fx["loss_train"]))) * 2
    assert int(weight) == int(fx["weight"]) and set(stats) == {"loss", "loss_mlm", "loss_copy"}
    loss.backward()
    worst = ("", 0.0)
    for n, p in m.named_parameters():
        gref = fx["grads"][n]
        assert p.grad is not None, n
        err = float((p.grad.cpu() - gref).abs().max())
        tol = 5e-4 * float(gref.abs().max()) + 5e-5
        if err / tol > worst[1]:
            worst = (n, err / tol)
        assert err <= tol, (n, err, tol)
    sd = m.state_dict()
    for k, v in fx["bn_after"].items():
        assert torch.allclose(sd[k].float().cpu(), v.float(), atol=1e-5), k


def test_eval_and_inference_match_reference(fx):
    m = _model(fx).eval()
    b = _cuda(fx["batch"])
    with torch.no_grad():
        loss, _, _ = m(**b)
        before, after, _, _ = m._forward(dict(speech_pad=b["speech"], text_pad=b["text"],
                                              masked_position=b["masked_position"], speech_mask=b["speech_mask"],
                                              text_mask=b["text_mask"], speech_segment_pos=b["speech_segment_pos"],
                                              text_segment_pos=b["text_segment_pos"]))
    assert abs(float(loss) - float(fx["loss_eval"])) < 2e-4
    assert torch.allclose(before.cpu(), fx["before_eval"], atol=2e-4)
    assert torch.allclose(after.cpu(), fx["after_eval"], atol=2e-4)
    b1 = {k: v[:1] for k, v in b.items() if k not in ("speech_lengths", "text_lengths")}
    out = m.inference(**b1, span_boundary=[20, 41], use_teacher_forcing=True)["feat_gen"]
    want = fx["inference"]
    assert torch.equal(out[0].cpu(), want[0]) and torch.equal(out[2].cpu(), want[2])
    assert torch.allclose(out[1].cpu(), want[1], atol=2e-4)


def test_bf16_mode_is_close(fx):
    """bf16 GEMM operands, fp32 accumulation/residuals: reported separately from the fp32 parity gate."""
    m = _model(fx, torch.bfloat16).train()
    loss, _, _ = m(**_cuda(fx["batch"]))
    rel = abs(float(loss) - float(fx["loss_train"])) / abs(float(fx["loss_train"]))
    assert rel < 2e-2, rel
    loss.backward()
    cos = {}
    for n, p in m.named_parameters():
        gref = fx["grads"][n].flatten()
        # linear_k.bias (softmax shift invariance) and depthwise_conv.bias (followed by BatchNorm) have a
        # mathematically zero gradient: only rounding noise to compare
        if n.endswith("linear_k.bias") or n.endswith("depthwise_conv.bias") or float(gref.abs().max()) < 1e-5:
            continue
        cos[n] = float(torch.nn.functional.cosine_similarity(p.grad.flatten().cpu(), gref, dim=0))
    worst = min(cos, key=cos.get)
    assert cos[worst] > 0.9 and sum(cos.values()) / len(cos) > 0.99, (worst, cos[worst], sum(cos.values()) / len(cos))


def test_dropout_training_runs_and_is_reproducible(fx):
    from a3t_b200.model import build_model

    conf = fx["conf"]
    m = build_model(conf["encoder_conf"], conf["decoder_conf"], conf["model_conf"], vocab_size=fx["vocab"])
    m.load_state_dict(fx["state_dict"])
    m = m.cuda().train()
    b = _cuda(fx["batch"])
    m._backend(b["speech"].device).set_seed(42)
    l1, _, _ = m(**b)
    l1.backward()
    g1 = m.sfc.weight.grad.clone()
    m.zero_grad()
    m.load_state_dict(fx["state_dict"])
    m._backend(b["speech"].device).set_seed(42)
    l2, _, _ = m(**b)
    l2.backward()
    assert float(l1) == float(l2) and torch.equal(g1, m.sfc.weight.grad)
    assert math.isfinite(float(l1)) and abs(float(l1) - float(fx["loss_train"])) > 1e-3  # dropout changed it
    l3, _, _ = m(**b)  # seed advanced -> different masks
    assert float(l3) != float(l2)


def test_frontend_matches_reference(golden_dir, cuda_lib):
    from a3t_b200.frontend import LogMelFbank

    fx = torch.load(os.path.join(golden_dir, "frontend.pt"), weights_only=False)
    for name, f in fx.items():
        fe = LogMelFbank(**f["kw"]).cuda()
        feats, lens = fe(f["wav"].cuda(), f["lens"].cuda())
        assert torch.equal(lens.cpu(), f["feats_lens"]), name
        err = float((feats.cpu() - f["feats"]).abs().max())
        assert torch.allclose(feats.cpu(), f["feats"], atol=1e-4, rtol=1e-4), (name, err)
        feats2, lens2 = fe(f["wav"].cuda(), None)
        assert torch.allclose(feats2.cpu(), f["feats_nolen"], atol=1e-4, rtol=1e-4), name
        assert fe.output_size() == 80 and fe.get_parameters()["n_shift"] == f["kw"]["hop_length"]


def test_frontend_full_size_properties(cuda_lib):
    """BASELINE size (B=16, T=1024): linearity in amplitude (log10 shift) and frame-shift equivariance."""
    from a3t_b200.frontend import LogMelFbank

    fe = LogMelFbank(fs=24000, n_fft=2048, win_length=1200, hop_length=300, fmin=80, fmax=7600, n_mels=80).cuda()
    torch.manual_seed(0)
    wav = 0.1 * torch.randn(16, 1023 * 300, device="cuda")
    m1, l1 = fe(wav, None)
    assert m1.shape == (16, 1024, 80) and int(l1[0]) == 1024
    m2, _ = fe(wav * 4.0, None)
    assert torch.allclose(m2, m1 + math.log10(4.0), atol=2e-4)
    m3, _ = fe(wav[:, 300:].contiguous(), None)  # shift by one hop: interior frames move by one
    assert torch.allclose(m3[:, 5:1000], m1[:, 6:1001], atol=2e-4)
    # against the oracle on a slice the CPU finishes quickly
    mo, _ = O.stft_logmel(wav[:2, :30000].cpu(), torch.tensor([30000, 30000]), fs=24000, n_fft=2048, win_length=1200,
                          hop=300, n_mels=80, fmin=80, fmax=7600)
    mc, _ = fe(wav[:2, :30000].contiguous(), None)
    assert torch.allclose(mc.cpu(), mo, atol=1e-4, rtol=1e-4)


def test_pwg_matches_reference(golden_dir, cuda_lib):
    from a3t_b200.vocoder import ParallelWaveGANGenerator, ParallelWaveGANPretrainedVocoder

    f = torch.load(os.path.join(golden_dir, "pwg.pt"), weights_only=False)
    gen = ParallelWaveGANGenerator(layers=f["layers"], stacks=f["stacks"], upsample_params={"upsample_scales": f["scales"]})
    gen.load_reference_state_dict(f["state_dict"])
    gen = gen.cuda()
    y = gen(f["c"].cuda(), f["z"].cuda())
    err = float((y.cpu() - f["wav"]).abs().max())
    assert torch.allclose(y.cpu(), f["wav"], atol=1e-4, rtol=1e-4), err
    voc = ParallelWaveGANPretrainedVocoder(gen, fs=24000)
    w1 = voc(f["c"][0].t().contiguous().cuda(), f["z"][0].t().contiguous().cuda())
    assert w1.shape == (2700,) and torch.allclose(w1.cpu(), f["wav_inference"].view(-1), atol=1e-4, rtol=1e-4)


def test_collate_kernels_bit_exact(golden_dir, cuda_lib):
    from a3t_b200 import collate as Cc

    c = torch.load(os.path.join(golden_dir, "kat.pt"), weights_only=False)["collate"]
    a_s = Cc.align_to_frames(c["t_start"].cuda(), 24000, 300)
    a_e = Cc.align_to_frames(c["t_end"].cuda(), 24000, 300)
    assert torch.equal(a_s.cpu(), c["align_start"]) and torch.equal(a_e.cpu(), c["align_end"])
    speech = torch.zeros(4, 90, 80, device="cuda")
    np.random.seed(c["seed"])
    mp, _ = Cc.phones_masking(speech, c["speech_mask"].cuda(), a_s, a_e, c["lens"], 0.8, 8)
    assert torch.equal(mp.cpu(), c["masked_position"])
    mp, _ = Cc.phones_masking(speech, c["speech_mask"].cuda(), a_s, a_e, c["lens"], 0.8, 8, span_boundary=c["span_boundary"])
    assert torch.equal(mp.cpu(), c["masked_position_span_boundary"])
    sp, tp = Cc.get_segment_pos(speech, torch.zeros(4, 14, dtype=torch.long, device="cuda"), a_s, a_e, c["lens"], True)
    assert torch.equal(sp.cpu(), c["sseg"]) and torch.equal(tp.cpu(), c["tseg"])
    # align floor at scale: random seconds, compare with the fp32 torch expression of the reference
    t = torch.rand(1 << 16, generator=torch.Generator().manual_seed(1)) * 20.0
    assert torch.equal(Cc.align_to_frames(t.cuda(), 24000, 300).cpu(), torch.floor(24000 * t / 300).int())


def test_cfg2_full_size_properties(cuda_lib):
    """BASELINE cfg2 shapes (B=16, Ts=1024, Tt=128): size-independent properties of the bf16 path:
    batch-permutation equivariance of per-utterance outputs in eval mode, padded-key invariance."""
    from a3t_b200.model import build_model
    from bench import paper_conf, synthetic_batch

    torch.manual_seed(0)
    enc, dec, mc = paper_conf()
    m = build_model(enc, dec, mc, act_dtype=torch.bfloat16).cuda().eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() == 1 and n.endswith("weight"):
                p.fill_(1.0)
    b = synthetic_batch(4, 1024, 128, device="cuda", seed=0)
    with torch.no_grad():
        l1, before1, after1 = m._run(b, need_loss=True)
        perm = torch.tensor([2, 0, 3, 1], device="cuda")
        b2 = {k: v[perm] for k, v in b.items()}
        l2, before2, after2 = m._run(b2, need_loss=True)
    assert torch.isfinite(l1).all()
    assert torch.allclose(after2, after1[perm], atol=2e-2, rtol=2e-2)
    assert abs(float(l1) - float(l2)) < 1e-3 * abs(float(l1))


def test_cfg2_full_size_backward_properties(cuda_lib):
    """BASELINE cfg2 at full size (B=16, Ts=1024, Tt=128, bf16 GEMMs, dropout ON), through the op graph the
    trainer runs: (1) with the same seed the forward is reproducible (the dropout masks are a pure function of
    seed / site / element index), (2) the hand-written backward is LINEAR in the incoming loss gradient:
    grads(2 g) == 2 grads(g) up to the summation-order noise of the atomically accumulated / split-K
    reductions, for every one of the 363 parameters, (3) all gradients are finite and non-trivial."""
    from a3t_b200 import graph
    from a3t_b200.model import build_model
    from bench import paper_conf, synthetic_batch

    torch.manual_seed(0)
    enc, dec, mc = paper_conf()
    m = build_model(enc, dec, mc, act_dtype=torch.bfloat16).cuda().train()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() == 1 and n.endswith("weight"):
                p.fill_(1.0)
    b = synthetic_batch(16, 1024, 128, device="cuda", seed=0)
    ops, P, wc, cfg = m._backend(torch.device("cuda", 0)), m._param_dict(), m._wcache, m.cfg
    res = []
    for scale in (1.0, 2.0):
        loss, before, after, ctx = graph.forward(ops, P, wc, cfg, b, True, True)   # same seed: same masks
        G = graph.backward(ops, P, wc, cfg, ctx, torch.full((1,), scale, device="cuda"))
        res.append((float(loss), {n: G[n].float().clone() for n in m._param_names}))
    (l1, g1), (l2, g2) = res
    assert math.isfinite(l1) and l1 == l2
    worst = 0.0
    for n in m._param_names:
        a, c = g1[n], g2[n]
        assert torch.isfinite(a).all() and torch.isfinite(c).all(), n
        sc = float(a.abs().max())
        if sc > 0:
            worst = max(worst, float((c - 2.0 * a).abs().max()) / sc)
    assert worst < 2e-3, worst
    assert sum(float(v.abs().sum()) > 0 for v in g1.values()) > 300  # (a few gradients are structurally zero)


def test_trainer_fp32_matches_oracle_update(fx):
    """DataParallelTrainer (single rank, fp32 kernels): reduced gradient, clip + Adam + Noam update and the
    statistics tail against the oracle's restatement of espnet2/train/trainer.py:583-675."""
    from a3t_b200.trainer import DataParallelTrainer

    m = _model(fx).train()
    tr = DataParallelTrainer(m)
    p0 = tr.flat_p.clone()
    stats = tr.step(_cuda(fx["batch"]))
    B = fx["batch"]["speech"].shape[0]
    assert abs(float(stats[0]) / B - float(fx["loss_train"])) < 2e-4 * abs(float(fx["loss_train"]))
    assert float(stats[2]) == B
    names = [n for n, _ in m.named_parameters()]
    gref = torch.cat([fx["grads"][n].reshape(-1) for n in names])
    g = tr.flat_g[:tr.n].cpu() / B
    assert float((g - gref).abs().max()) <= 5e-4 * float(gref.abs().max()) + 5e-5
    pe, mb, vb = p0.cpu().clone(), torch.zeros(tr.n), torch.zeros(tr.n)
    O.clip_adam_step(pe, gref.clone(), mb, vb, 1, O.noam_lr(1.0, m.encoder.attention_dim, 4000.0, 1))
    assert int(tr.step_count) == 1
    # Adam's first step is lr * sign(g): compare where the gradient is not rounding noise
    big = gref.abs() > 1e-4 * float(gref.abs().max())
    assert torch.allclose(tr.flat_p.cpu()[big], pe[big], atol=1e-9 + 1e-3 * float((pe - p0.cpu()).abs().max()))


def test_trainer_bf16_in_place_repack_equals_lazy_repack(fx):
    """The batched in-place weight repack (one kernel per step) must feed the GEMMs the same bf16 operands
    as the lazy per-weight packing it replaces: same losses over several optimizer steps."""
    from a3t_b200.trainer import DataParallelTrainer

    losses = []
    for lazy in (False, True):
        m = _model(fx, torch.bfloat16).train()
        tr = DataParallelTrainer(m, lr=1e-3, warmup=0.0)
        tr.in_place_repack = not lazy  # lazy: never build the plan -> cache dropped every step
        b = _cuda(fx["batch"])
        ls = []
        for _ in range(4):
            st = tr.step(b)
            ls.append(float(st[0] / st[2]))
        assert (tr._plan is None) == lazy
        losses.append(ls)
    assert losses[0][0] != losses[0][3]  # the parameters did move
    for a, bb in zip(*losses):
        assert abs(a - bb) <= 2e-3 * abs(bb), losses
