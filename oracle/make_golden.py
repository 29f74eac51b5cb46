"""Generate tests/golden/*.pt by running the REFERENCE itself (imported from /root/reference).

Run in the build container only:  python -m oracle.make_golden
The fixtures travel with the repo; nothing at test time on the GPU box reads /root/reference.
Each fixture stores inputs, the reference's outputs, and (for the model) the full state_dict.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

TINY = dict(num_blocks=1, attention_dim=32, attention_heads=2, linear_units=64)


def tiny_conf():
    conf = R.model_conf("paper")
    conf["encoder_conf"].update(TINY)
    conf["decoder_conf"].update(TINY)
    conf["model_conf"].update(postnet_chans=32)
    return conf


def model_fixture():
    conf = tiny_conf()
    vocab = 20
    ref = R.build_reference_model(conf, vocab=vocab, dropout_zero=True)
    R.randomize_degenerate_params(ref)
    batch, aux = R.synthetic_batch(3, 70, 12, vocab=vocab, seed=3, ragged=True)
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    # train mode, dropout 0 (BatchNorm in batch-stat mode)
    ref.train()
    ref.zero_grad(set_to_none=True)
    loss, stats, weight = ref(**batch)
    loss.backward()
    grads = {n: p.grad.clone() for n, p in ref.named_parameters()}
    sd_after = {k: v.clone() for k, v in ref.state_dict().items() if "running" in k or "num_batches" in k}
    before_t, after_t, _, _ = None, None, None, None
    ref.load_state_dict(sd0)
    ref.eval()
    with torch.no_grad():
        before_e, after_e, _, _ = ref._forward(
            dict(speech_pad=batch["speech"], text_pad=batch["text"], masked_position=batch["masked_position"],
                 speech_mask=batch["speech_mask"], text_mask=batch["text_mask"],
                 speech_segment_pos=batch["speech_segment_pos"], text_segment_pos=batch["text_segment_pos"]),
            batch["speech_segment_pos"])
        loss_e, _, _ = ref(**batch)
        inf = ref.inference(**{k: v[:1] for k, v in batch.items() if k not in ("speech_lengths", "text_lengths")},
                            span_boundary=[20, 41], use_teacher_forcing=True)
    torch.save(dict(conf=conf, vocab=vocab, state_dict=sd0, batch=batch, aux=aux, loss_train=loss.detach(),
                    weight=weight, grads=grads, bn_after=sd_after, loss_eval=loss_e, before_eval=before_e,
                    after_eval=after_e, inference=[t.clone() for t in inf["feat_gen"]]),
               os.path.join(OUT, "model_tiny.pt"))
    print("model_tiny: loss_train", float(loss), "loss_eval", float(loss_e), "params",
          sum(p.numel() for p in ref.parameters()))


def kat_fixture():
    R._activate()
    from espnet2.train.collate_fn import get_segment_pos, phones_masking, random_spans_noise_mask
    from espnet.nets.pytorch_backend.transformer.attention import LegacyRelPositionMultiHeadedAttention
    from espnet.nets.pytorch_backend.transformer.embedding import LegacyRelPositionalEncoding

    out = {}
    np.random.seed(0)
    out["span_30_0.8_8"] = random_spans_noise_mask(30, 0.8, 8).astype(np.uint8)
    np.random.seed(0)
    out["span_20_0.15_3"] = random_spans_noise_mask(20, 0.15, 3).astype(np.uint8)
    np.random.seed(7)
    out["span_seq"] = [random_spans_noise_mask(L, 0.8, 8).astype(np.uint8) for L in (2, 3, 17, 64, 128, 200)]
    pe = LegacyRelPositionalEncoding(8, 0.0)
    out["pos_4_8"] = pe(torch.zeros(1, 4, 8))[1].clone()
    pe = LegacyRelPositionalEncoding(16, 0.0, max_len=10)
    out["pos_12_16_maxlen10"] = pe(torch.zeros(1, 12, 16))[1].clone()
    att = LegacyRelPositionMultiHeadedAttention(2, 8, 0.0)
    out["rel_shift_4"] = att.rel_shift(torch.arange(16.0).view(1, 1, 4, 4)).clone()
    x = torch.randn(2, 3, 7, 7, generator=torch.Generator().manual_seed(0))
    out["rel_shift_in"], out["rel_shift_out"] = x, att.rel_shift(x).clone()
    # alignment floor, span expansion and segment ids on a ragged batch
    g = torch.Generator().manual_seed(5)
    B, Ts, Tt = 4, 90, 14
    lens = torch.tensor([14, 9, 2, 1])
    t_start = torch.zeros(B, Tt)
    t_end = torch.zeros(B, Tt)
    for b in range(B):
        cuts = torch.sort(torch.rand(int(lens[b]) + 1, generator=g) * (Ts - 1) * 300 / 24000)[0]
        t_start[b, : lens[b]] = cuts[:-1]
        t_end[b, : lens[b]] = cuts[1:]
    a_s = torch.floor(24000 * t_start / 300).int()
    a_e = torch.floor(24000 * t_end / 300).int()
    speech = torch.randn(B, Ts, 80, generator=g)
    slens = torch.tensor([90, 77, 60, 33])
    smask = (torch.arange(Ts)[None] < slens[:, None]).unsqueeze(-2)
    np.random.seed(11)
    mp, _ = phones_masking(speech, smask, a_s, a_e, lens, 0.8, 8)
    sseg, tseg = get_segment_pos(speech, torch.zeros(B, Tt, dtype=torch.long), a_s, a_e, lens, True)
    mp_sb, _ = phones_masking(speech, smask, a_s, a_e, lens, 0.8, 8, span_boundary=[[3, 9], [0, 80], [50, 70, 10, 20], [5, 5]])
    out["collate"] = dict(t_start=t_start, t_end=t_end, align_start=a_s, align_end=a_e, lens=lens, slens=slens,
                          speech_mask=smask, masked_position=mp, sseg=sseg, tseg=tseg, seed=11,
                          masked_position_span_boundary=mp_sb,
                          span_boundary=[[3, 9], [0, 80], [50, 70, 10, 20], [5, 5]])
    torch.save(out, os.path.join(OUT, "kat.pt"))
    print("kat:", "".join(map(str, out["span_30_0.8_8"])))


def frontend_fixture():
    R._activate()
    from espnet2.tts.feats_extract.log_mel_fbank import LogMelFbank

    g = torch.Generator().manual_seed(2)
    res = {}
    for name, kw, N in (("vctk24k", dict(fs=24000, n_fft=2048, win_length=1200, hop_length=300, fmin=80, fmax=7600, n_mels=80), 6100),
                        ("16k", dict(fs=16000, n_fft=1024, win_length=800, hop_length=200, fmin=80, fmax=7600, n_mels=80), 3333)):
        fe = LogMelFbank(**kw)
        wav = 0.1 * torch.randn(3, N, generator=g)
        wav[1] += 0.3 * torch.sin(torch.arange(N) * 0.05)
        lens = torch.tensor([N, N - 700, N // 2])
        with torch.no_grad():
            feats, flens = fe(wav, lens)
            feats_full, flens_full = fe(wav, None)
        res[name] = dict(kw=kw, wav=wav, lens=lens, feats=feats, feats_lens=flens, feats_nolen=feats_full,
                         melmat=fe.logmel.melmat.clone())
    torch.save(res, os.path.join(OUT, "frontend.pt"))
    print("frontend:", {k: tuple(v["feats"].shape) for k, v in res.items()})


def pwg_fixture():
    R._activate()
    from espnet2.gan_tts.parallel_wavegan import ParallelWaveGANGenerator

    torch.manual_seed(4)
    gen = ParallelWaveGANGenerator(layers=6, stacks=2, upsample_params={"upsample_scales": [4, 5, 3, 5]})
    gen.remove_weight_norm()
    gen.eval()
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for p in gen.parameters():  # weight-norm init leaves a few tensors tiny; make all of them matter
            p.copy_(0.15 * torch.randn(p.shape, generator=g))
    c = torch.randn(2, 80, 9, generator=g)
    z = torch.randn(2, 1, 9 * 300, generator=g)
    with torch.no_grad():
        y = gen(c, z)
        y1 = gen.inference(c[0].t().contiguous(), z[0].t().contiguous())
    torch.save(dict(layers=6, stacks=2, scales=[4, 5, 3, 5], state_dict={k: v.clone() for k, v in gen.state_dict().items()},
                    c=c, z=z, wav=y, wav_inference=y1), os.path.join(OUT, "pwg.pt"))
    print("pwg:", tuple(y.shape), float(y.abs().mean()))


def model_d384_fixture():
    """Paper WIDTH (D=384, H=2, FF=1536, dw 7/31, postnet 5x256) with 1+1 blocks, B=2, Ts=1024, Tt=128, ragged.
    Weights come from `fixtures.fill_params(seed)` (regenerated by the tests), outputs from the reference."""
    from oracle.fixtures import fill_params, grad_probe

    conf = R.model_conf("paper")
    conf["encoder_conf"].update(num_blocks=1)
    conf["decoder_conf"].update(num_blocks=1)
    vocab, wseed = 73, 20241
    ref = R.build_reference_model(conf, vocab=vocab, dropout_zero=True)
    fill_params(ref, wseed)
    batch, aux = R.synthetic_batch(2, 1024, 128, vocab=vocab, seed=5, ragged=True)
    # make the second utterance clearly shorter (padded keys + padded frames at the paper shape)
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    ref.train()
    ref.zero_grad(set_to_none=True)
    loss, stats, weight = ref(**batch)
    loss.backward()
    grads = {n: p.grad.clone() for n, p in ref.named_parameters()}
    bn_after = {k: v.clone() for k, v in ref.state_dict().items() if "running" in k}
    ref.load_state_dict(sd0)
    ref.eval()
    with torch.no_grad():
        before_e, after_e, _, _ = ref._forward(
            dict(speech_pad=batch["speech"], text_pad=batch["text"], masked_position=batch["masked_position"],
                 speech_mask=batch["speech_mask"], text_mask=batch["text_mask"],
                 speech_segment_pos=batch["speech_segment_pos"], text_segment_pos=batch["text_segment_pos"]),
            batch["speech_segment_pos"])
        loss_e, _, _ = ref(**batch)
    small = {k: v for k, v in batch.items() if k != "speech"}   # speech is regenerated: randn under manual_seed(5)
    torch.save(dict(conf=conf, vocab=vocab, weight_seed=wseed, batch_seed=5, batch_small=small,
                    speech_lengths=batch["speech_lengths"], speech_probe=batch["speech"][:, ::97, ::7].clone(),
                    loss_train=loss.detach(), loss_eval=loss_e,
                    grad_norm={n: float(g.norm()) for n, g in grads.items()},
                    grad_probe={n: grad_probe(g) for n, g in grads.items()},
                    bn_after_probe={k: v.reshape(-1)[:16].clone() for k, v in bn_after.items()},
                    before_eval=before_e[:, ::8].clone(), after_eval=after_e[:, ::8].clone()),
               os.path.join(OUT, "model_d384.pt"))
    print("model_d384: loss_train", float(loss), "loss_eval", float(loss_e), "params",
          sum(p.numel() for p in ref.parameters()))


def pwg30_fixture():
    """The published generator shape: 30 layers / 3 stacks (dilations 1..512), 200 frames -> 60 000 samples."""
    R._activate()
    from espnet2.gan_tts.parallel_wavegan import ParallelWaveGANGenerator
    from oracle.fixtures import fill_params

    torch.manual_seed(4)
    gen = ParallelWaveGANGenerator(layers=30, stacks=3, upsample_params={"upsample_scales": [4, 5, 3, 5]})
    gen.remove_weight_norm()
    gen.eval()
    fill_params(gen, 77, scale=0.8)
    g = torch.Generator().manual_seed(10)
    c = torch.randn(1, 80, 200, generator=g)
    z = torch.randn(1, 1, 200 * 300, generator=g)
    with torch.no_grad():
        y = gen(c, z)
    torch.save(dict(layers=30, stacks=3, scales=[4, 5, 3, 5], weight_seed=77, weight_scale=0.8,
                    param_shapes={k: tuple(v.shape) for k, v in gen.state_dict().items()},
                    c=c, z_seed=10, z_probe=z[0, 0, ::601].clone(), wav=y.clone()), os.path.join(OUT, "pwg30.pt"))
    print("pwg30:", tuple(y.shape), float(y.abs().mean()), float(y.abs().max()))


def collate_fixture():
    """`mlm_collate_fn` end to end (espnet2/train/collate_fn.py:158-287) on raw utterances: text+alignment batch,
    the span_boundary (inference) batch and the speech-only batch."""
    R._activate()
    from espnet2.train.collate_fn import mlm_collate_fn
    from espnet2.tts.feats_extract.log_mel_fbank import LogMelFbank

    kw = dict(fs=24000, n_fft=2048, win_length=1200, hop_length=300, fmin=80, fmax=7600, n_mels=80)
    fe = LogMelFbank(**kw)
    rng = np.random.RandomState(3)
    data = []
    for i, (n, L) in enumerate(((9100, 11), (6400, 7), (7777, 2), (5000, 1))):
        wav = (0.1 * rng.randn(n)).astype(np.float32)
        cuts = np.sort(rng.rand(L + 1)) * (n / 24000.0)
        data.append((f"utt{i}", dict(speech=wav, text=rng.randint(2, 70, size=L).astype(np.int64),
                                     align_start=cuts[:-1].astype(np.float32), align_end=cuts[1:].astype(np.float32))))
    out = {"kw": kw, "data": data}
    np.random.seed(21)
    with torch.no_grad():
        out["train"] = mlm_collate_fn(data, float_pad_value=0.0, int_pad_value=0, mlm_prob=0.8, mean_phn_span=8,
                                      feats_extract=fe, sega_emb=True)
    sb_data = [(u, dict(d, span_boundary=np.array(sb, dtype=np.int64)))
               for (u, d), sb in zip(data, ([3, 9], [0, 15], [10, 20], [5, 5]))]
    np.random.seed(22)
    with torch.no_grad():
        out["span_boundary"] = mlm_collate_fn(sb_data, float_pad_value=0.0, int_pad_value=0, mlm_prob=0.8,
                                              mean_phn_span=8, feats_extract=fe, sega_emb=True)
    so_data = [(u, dict(speech=d["speech"])) for u, d in data]
    np.random.seed(23)
    with torch.no_grad():
        out["speech_only"] = mlm_collate_fn(so_data, float_pad_value=0.0, int_pad_value=0, mlm_prob=0.8,
                                            mean_phn_span=8, feats_extract=fe, sega_emb=True)
    out["seeds"] = dict(train=21, span_boundary=22, speech_only=23)
    torch.save(out, os.path.join(OUT, "collate.pt"))
    print("collate:", {k: {kk: tuple(vv.shape) for kk, vv in out[k][1].items()} for k in ("train", "span_boundary", "speech_only")})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = set(sys.argv[1:]) or {"model", "kat", "frontend", "pwg", "d384", "pwg30", "collate"}
    for name, fn in (("model", model_fixture), ("kat", kat_fixture), ("frontend", frontend_fixture), ("pwg", pwg_fixture),
                     ("d384", model_d384_fixture), ("pwg30", pwg30_fixture), ("collate", collate_fixture)):
        if name in which:
            fn()
