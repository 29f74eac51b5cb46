"""The CPU oracle (oracle/a3t_oracle.py) against fixtures produced by the REFERENCE itself
(oracle/make_golden.py) and the known-answer values of SURVEY.md 8c."""
import os

import numpy as np
import pytest
import torch

from oracle import a3t_oracle as O


@pytest.fixture(scope="module")
def kat(golden_dir):
    return torch.load(os.path.join(golden_dir, "kat.pt"), weights_only=False)


def test_span_sampler_known_answers(kat):
    np.random.seed(0)
    m = O.random_spans_noise_mask(30, 0.8, 8)
    assert "".join(str(int(v)) for v in m) == "000011111101111111111111110111"
    assert np.array_equal(m.astype(np.uint8), kat["span_30_0.8_8"])
    np.random.seed(0)
    m = O.random_spans_noise_mask(20, 0.15, 3)
    assert "".join(str(int(v)) for v in m) == "0" * 17 + "111"
    assert np.array_equal(m.astype(np.uint8), kat["span_20_0.15_3"])
    np.random.seed(7)
    for L, want in zip((2, 3, 17, 64, 128, 200), kat["span_seq"]):
        assert np.array_equal(O.random_spans_noise_mask(L, 0.8, 8).astype(np.uint8), want)


def test_product_span_sampler_matches_reference_rng_stream(kat):
    from a3t_b200.collate import random_spans_noise_mask

    np.random.seed(7)
    for L, want in zip((2, 3, 17, 64, 128, 200), kat["span_seq"]):
        assert np.array_equal(random_spans_noise_mask(L, 0.8, 8).astype(np.uint8), want)


def test_positional_table(kat):
    t = O.legacy_rel_pos_table(4, 8)
    assert torch.equal(t, kat["pos_4_8"][0])
    assert abs(float(t[0, 0]) - (-0.6639)) < 1e-4 and abs(float(t[0, 1]) - (-0.7478)) < 1e-4  # sin/cos(4999)
    assert torch.equal(O.legacy_rel_pos_table(12, 16, max_len=10), kat["pos_12_16_maxlen10"][0])
    from a3t_b200.graph import legacy_rel_pos_table

    assert torch.equal(legacy_rel_pos_table(4, 8, "cpu"), kat["pos_4_8"][0])
    assert torch.equal(legacy_rel_pos_table(12, 16, "cpu", max_len=10), kat["pos_12_16_maxlen10"][0])


def test_rel_shift(kat):
    want = torch.tensor([[3, 0, 4, 5], [6, 7, 0, 8], [9, 10, 11, 0], [12, 13, 14, 15.0]])
    assert torch.equal(O.rel_shift(torch.arange(16.0).view(1, 1, 4, 4))[0, 0], want)
    assert torch.equal(kat["rel_shift_4"][0, 0], want)
    assert torch.equal(O.rel_shift(kat["rel_shift_in"]), kat["rel_shift_out"])


def test_collate_integer_math(kat):
    c = kat["collate"]
    a_s = O.align_to_frames(c["t_start"], 24000, 300)
    a_e = O.align_to_frames(c["t_end"], 24000, 300)
    assert torch.equal(a_s, c["align_start"]) and torch.equal(a_e, c["align_end"])
    np.random.seed(c["seed"])
    pm = O.draw_phone_masks(c["lens"], 0.8, 8, a_s.shape[1])
    valid = c["speech_mask"].reshape(4, -1)
    mp = O.expand_phone_mask(pm, a_s, a_e, c["lens"], valid)
    assert torch.equal(mp, c["masked_position"])
    mp = O.expand_phone_mask(None, a_s, a_e, c["lens"], valid, span_boundary=c["span_boundary"])
    assert torch.equal(mp, c["masked_position_span_boundary"])
    sp, tp = O.segment_pos(a_s, a_e, c["lens"], valid.shape[1], a_s.shape[1])
    assert torch.equal(sp, c["sseg"]) and torch.equal(tp, c["tseg"])


def test_frontend_against_reference(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "frontend.pt"), weights_only=False)
    for name, f in fx.items():
        kw = f["kw"]
        mel, olens = O.stft_logmel(f["wav"], f["lens"], fs=kw["fs"], n_fft=kw["n_fft"], win_length=kw["win_length"],
                                   hop=kw["hop_length"], n_mels=kw["n_mels"], fmin=kw["fmin"], fmax=kw["fmax"])
        assert torch.equal(olens, f["feats_lens"])
        assert torch.allclose(mel, f["feats"], atol=1e-4, rtol=1e-4), name
        # mel matrix: restated Slaney formula == the matrix the reference run used, and == torchaudio
        assert torch.equal(O.slaney_mel_matrix(kw["fs"], kw["n_fft"], kw["n_mels"], kw["fmin"], kw["fmax"]), f["melmat"])
        torchaudio = pytest.importorskip("torchaudio")
        ta = torchaudio.functional.melscale_fbanks(kw["n_fft"] // 2 + 1, float(kw["fmin"]), float(kw["fmax"]),
                                                   kw["n_mels"], kw["fs"], norm="slaney", mel_scale="slaney")
        assert torch.allclose(ta, f["melmat"], atol=1e-6)
        from a3t_b200.frontend import slaney_mel_filterbank

        assert np.array_equal(slaney_mel_filterbank(kw["fs"], kw["n_fft"], kw["n_mels"], kw["fmin"], kw["fmax"]),
                              f["melmat"].numpy())


def test_pwg_against_reference(golden_dir):
    f = torch.load(os.path.join(golden_dir, "pwg.pt"), weights_only=False)
    y = O.pwg_generate(f["c"], f["z"], f["state_dict"], upsample_scales=f["scales"], layers=f["layers"],
                       stacks=f["stacks"])
    assert torch.allclose(y, f["wav"], atol=1e-5, rtol=1e-5)


def test_dropout_hash_statistics():
    k = O.keep_mask(1 << 20, 0.2, seed=1234567890123, site=5)
    assert abs(float(k.float().mean()) - 0.8) < 2e-3
    k2 = O.keep_mask(1 << 20, 0.2, seed=1234567890123, site=6)
    assert abs(float((k & k2).float().mean()) - 0.64) < 3e-3  # sites are independent


def test_clip_adam_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(1000)
    g = torch.randn(1000) * 3
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref_p], lr=1.0)
    sched_lr = O.noam_lr(1.0, 384, 4000, 1)
    for grp in opt.param_groups:
        grp["lr"] = sched_lr
    ref_p.grad = g.clone()
    torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
    opt.step()
    m, v = torch.zeros(1000), torch.zeros(1000)
    pp = p.clone()
    O.clip_adam_step(pp, g.clone(), m, v, 1, sched_lr)
    assert torch.allclose(pp, ref_p.detach(), atol=1e-7)


def test_dropout_pair_hash_independence():
    """The stateless dropout mask (one 32-bit hash per element PAIR, 16 bits each; device twin in
    csrc/common.cuh) keeps with probability 1-p, the two elements of a pair are uncorrelated, and different
    sites / seeds give unrelated masks."""
    n = 1 << 20
    for p in (0.1, 0.2, 0.5):
        m = O.keep_mask(n, p, seed=1234567, site=3).numpy().astype(np.float64)
        q = 1.0 - np.floor(p * 65536.0) / 65536.0           # 16-bit threshold
        assert abs(m.mean() - q) < 4.0 * np.sqrt(q * (1 - q) / n)
        a, b = m[0::2] - q, m[1::2] - q                      # the two halves of one hash
        assert abs((a * b).mean()) < 4.0 * q * (1 - q) / np.sqrt(n / 2)
        c = m[1::2][:-1] - q                                 # neighbours from different hashes
        d = m[2::2] - q
        assert abs((c * d[: c.size]).mean()) < 4.0 * q * (1 - q) / np.sqrt(n / 2)
        m2 = O.keep_mask(n, p, seed=1234567, site=4).numpy().astype(np.float64)
        m3 = O.keep_mask(n, p, seed=1234568, site=3).numpy().astype(np.float64)
        for other in (m2, m3):
            assert abs(((m - q) * (other - q)).mean()) < 4.0 * q * (1 - q) / np.sqrt(n)
    assert bool(O.keep_mask(17, 0.0, 1, 1).all())



# ---- round 2: fixtures at the paper width / published vocoder depth / whole collate -----------------
def test_oracle_graph_matches_reference_at_paper_width(golden_dir):
    """D=384, H=2, FF=1536, dw 7/31, postnet 5x256, 1+1 blocks, B=2, Ts=1024, Tt=128 ragged: the op graph
    (a3t_b200/graph.py) over the oracle ops against the numbers the reference produced."""
    import _d384
    from a3t_b200 import graph
    from oracle.fixtures import grad_probe
    from oracle.oracle_backend import OracleBackend

    fx, b = _d384.load(golden_dir)
    m = _d384.build(fx)
    P = {n: p.detach() for n, p in m.named_parameters()}
    P.update({n: v.clone() for n, v in m.named_buffers()})
    ops, wc = OracleBackend(), graph.WeightCache()
    loss, before, after, ctx = graph.forward(ops, P, wc, m.cfg, b, True, True)
    assert abs(float(loss) - float(fx["loss_train"])) <= 1e-5 * abs(float(fx["loss_train"]))
    G = graph.backward(ops, P, wc, m.cfg, ctx, torch.ones(1))
    for n, _ in m.named_parameters():
        ref = fx["grad_probe"][n]
        assert float((grad_probe(G[n]) - ref).abs().max()) <= 5e-4 * float(ref.abs().max()) + 5e-5, n
        assert abs(float(G[n].norm()) - fx["grad_norm"][n]) <= 1e-3 * fx["grad_norm"][n] + 1e-5, n
    P = {n: p.detach() for n, p in m.named_parameters()}
    P.update({n: v.clone() for n, v in _d384.build(fx).named_buffers()})
    le, before, after, _ = graph.forward(ops, P, graph.WeightCache(), m.cfg, b, False, True)
    assert abs(float(le) - float(fx["loss_eval"])) <= 1e-5 * abs(float(fx["loss_eval"]))
    assert torch.allclose(before[:, ::8], fx["before_eval"], atol=2e-4, rtol=1e-5)
    assert torch.allclose(after[:, ::8], fx["after_eval"], atol=2e-4, rtol=1e-5)


def test_pwg30_against_reference(golden_dir):
    """30 layers / 3 stacks / dilations 1..512 / 200 frames (the published generator shape)."""
    import _d384

    f = torch.load(os.path.join(golden_dir, "pwg30.pt"), weights_only=False)
    y = O.pwg_generate(f["c"], _d384.pwg30_z(f), _d384.pwg30_state_dict(f), upsample_scales=f["scales"],
                       layers=f["layers"], stacks=f["stacks"])
    assert torch.allclose(y, f["wav"], atol=1e-5, rtol=1e-5)


def test_oracle_collate_against_reference_mlm_collate_fn(golden_dir):
    """The oracle's pieces composed as espnet2/train/collate_fn.py:158-287 composes them, on raw utterances,
    against the reference functor's output dict: text+alignment, span_boundary and speech-only batches."""
    fx = torch.load(os.path.join(golden_dir, "collate.pt"), weights_only=False)
    kw = fx["kw"]
    data = fx["data"]

    def collate(dd, span_boundary=None):
        B = len(dd)
        n = max(d["speech"].shape[0] for _, d in dd)
        wav = torch.zeros(B, n)
        wl = torch.tensor([d["speech"].shape[0] for _, d in dd])
        for i, (_, d) in enumerate(dd):
            wav[i, : wl[i]] = torch.from_numpy(d["speech"])
        mel, ol = O.stft_logmel(wav, wl, fs=kw["fs"], n_fft=kw["n_fft"], win_length=kw["win_length"], hop=kw["hop_length"],
                                n_mels=kw["n_mels"], fmin=kw["fmin"], fmax=kw["fmax"])
        Ts = int(ol.max())
        mel = mel[:, :Ts]
        valid = torch.arange(Ts)[None] < ol[:, None]
        if "text" not in dd[0][1]:
            m = O.random_spans_noise_mask(Ts, 0.15, min(Ts * 0.15 // 3, 50))
            mp = torch.from_numpy(m)[None].expand(B, Ts) & valid
            return dict(speech=mel, text=torch.zeros(B, 1, dtype=torch.long) - 2, masked_position=mp,
                        speech_segment_pos=torch.zeros(B, Ts, dtype=torch.long),
                        text_segment_pos=torch.zeros(B, 1, dtype=torch.long))
        Tt = max(d["text"].shape[0] for _, d in dd)
        lens = torch.tensor([d["text"].shape[0] for _, d in dd])
        text = torch.zeros(B, Tt, dtype=torch.long)
        ts, te = torch.zeros(B, Tt), torch.zeros(B, Tt)
        for i, (_, d) in enumerate(dd):
            text[i, : lens[i]] = torch.from_numpy(d["text"])
            ts[i, : lens[i]] = torch.from_numpy(d["align_start"])
            te[i, : lens[i]] = torch.from_numpy(d["align_end"])
        a_s, a_e = O.align_to_frames(ts, kw["fs"], kw["hop_length"]), O.align_to_frames(te, kw["fs"], kw["hop_length"])
        if span_boundary is not None:
            mp = O.expand_phone_mask(None, a_s, a_e, lens, valid, span_boundary=span_boundary)
        else:
            mp = O.expand_phone_mask(O.draw_phone_masks(lens, 0.8, 8, Tt), a_s, a_e, lens, valid)
        sp, tp = O.segment_pos(a_s, a_e, lens, Ts, Tt)
        return dict(speech=mel, text=text, masked_position=mp, speech_segment_pos=sp, text_segment_pos=tp)

    sbs = ([3, 9], [0, 15], [10, 20], [5, 5])
    for name, dd, sb in (("train", data, None), ("span_boundary", data, sbs),
                         ("speech_only", [(u, dict(speech=d["speech"])) for u, d in data], None)):
        np.random.seed(fx["seeds"][name])
        got = collate(dd, sb)
        ref = fx[name][1]
        for k, v in got.items():
            if k == "speech":
                assert torch.allclose(v, ref[k], atol=1e-4, rtol=1e-4), (name, k)
            else:
                assert torch.equal(v.to(ref[k].dtype), ref[k]), (name, k)
