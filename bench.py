#!/usr/bin/env python
"""Benchmark of the A3T masked-mel training hot path (BASELINE.json metric: mel-frames/s of training).

  python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, one process per GPU)
  python bench.py --impl reference --steps K --warmup W    # the reference's own modules (oracle/_ref snapshot) on host cores
  python bench.py --config cfg4 | cfg5                     # BASELINE configs[3] per-GPU shape / configs[4] (infill + vocoder)

A "step" = forward + backward + gradient all-reduce + clip/Adam/Noam of `ESPnetMLMEncAsDecoderModel`
on one synthetic batch per GPU.  Workload = BASELINE configs[1] ("cfg2"): the VCTK paper Conformer
(conf/fsp2_conformer.yaml), bf16 GEMM operands, B=16 utterances x Ts=1024 mel frames + Tt=128 phones
per GPU; frames counted = B*Ts.  Weak scaling for N>1 (per-GPU batch fixed, one NCCL all-reduce).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def paper_conf():
    """Values of egs2/vctk/sedit/conf/fsp2_conformer.yaml:27-75 (encoder_conf, decoder_conf, model_conf)."""
    common = dict(attention_dim=384, attention_heads=2, linear_units=1536, num_blocks=4, dropout_rate=0.2,
                  positional_dropout_rate=0.2, attention_dropout_rate=0.2, macaron_style=True, use_cnn_module=True,
                  selfattention_layer_type="rel_selfattn", activation_type="swish", pos_enc_layer_type="rel_pos",
                  positionwise_layer_type="conv1d", positionwise_conv_kernel_size=3)
    enc = dict(common, input_layer="sega_mlm", pre_speech_layer=0, cnn_module_kernel=7, normalize_before=True)
    dec = dict(common, cnn_module_kernel=31)
    mc = dict(lsm_weight=0.1, length_normalized_loss=False, masking_schema="phn_span", mean_phn_span=8, mlm_prob=0.8,
              dynamic_mlm_prob=False, postnet_layers=5, postnet_filts=5, postnet_chans=256)
    return enc, dec, mc


def synthetic_batch_host(B, Ts, Tt, vocab=73, seed=0, mlm_prob=0.8, mean_phn_span=8):
    """SURVEY 8d synthetic inputs on the HOST (numpy/torch CPU): random phoneme ids, randn mel, even
    phone/frame alignment, T5 span mask drawn from np.random like the reference's collate."""
    from a3t_b200.collate import draw_phone_masks

    g = torch.Generator().manual_seed(seed)
    np.random.seed(seed)
    speech = torch.randn(B, Ts, 80, generator=g)
    text = torch.randint(2, vocab - 1, (B, Tt), generator=g)
    edges = torch.floor(torch.linspace(0, Ts, Tt + 1)).int()
    a_s = edges[:-1].unsqueeze(0).repeat(B, 1).contiguous()
    a_e = edges[1:].unsqueeze(0).repeat(B, 1).contiguous()
    lens = torch.full((B,), Tt, dtype=torch.int64)
    pm = torch.from_numpy(draw_phone_masks(lens.tolist(), mlm_prob, mean_phn_span, Tt))
    return dict(speech=speech, text=text, align_start=a_s, align_end=a_e, align_lengths=lens, phone_mask=pm,
                speech_mask=torch.ones(B, 1, Ts, dtype=torch.bool), text_mask=torch.ones(B, 1, Tt, dtype=torch.bool))


def expand_on_host(h):
    """masked_position / segment ids with the oracle's integer routines (CPU reference arm only)."""
    from oracle import a3t_oracle as O

    B, Ts = h["speech"].shape[:2]
    Tt = h["text"].shape[1]
    mp = O.expand_phone_mask(h["phone_mask"].numpy(), h["align_start"], h["align_end"], h["align_lengths"],
                             h["speech_mask"].reshape(B, Ts))
    sseg, tseg = O.segment_pos(h["align_start"], h["align_end"], h["align_lengths"], Ts, Tt)
    return dict(speech=h["speech"], text=h["text"], masked_position=mp, speech_mask=h["speech_mask"],
                text_mask=h["text_mask"], speech_segment_pos=sseg, text_segment_pos=tseg)


def device_batch(h, device):
    """Host batch -> device batch; the span expansion and segment ids run as CUDA kernels."""
    from a3t_b200 import _lib

    d = {k: v.to(device, non_blocking=True) for k, v in h.items()}
    B, Ts = d["speech"].shape[:2]
    Tt = d["text"].shape[1]
    st = torch.cuda.current_stream(device).cuda_stream
    mp = torch.empty(B, Ts, dtype=torch.uint8, device=device)
    valid = d["speech_mask"].reshape(B, Ts).view(torch.uint8)
    _lib.call("a3t_expand_phone_mask", d["phone_mask"].data_ptr(), d["align_start"].data_ptr(),
              d["align_end"].data_ptr(), d["align_lengths"].data_ptr(), valid.data_ptr(), mp.data_ptr(), B, Ts, Tt, st)
    sseg = torch.empty(B, Ts, dtype=torch.int64, device=device)
    tseg = torch.empty(B, Tt, dtype=torch.int64, device=device)
    _lib.call("a3t_segment_pos", d["align_start"].data_ptr(), d["align_end"].data_ptr(), d["align_lengths"].data_ptr(),
              sseg.data_ptr(), tseg.data_ptr(), B, Ts, Tt, st)
    return dict(speech=d["speech"], text=d["text"], masked_position=mp.view(torch.bool), speech_mask=d["speech_mask"],
                text_mask=d["text_mask"], speech_segment_pos=sseg, text_segment_pos=tseg)


def synthetic_batch(B, Ts, Tt, device="cuda", seed=0):
    return device_batch(synthetic_batch_host(B, Ts, Tt, seed=seed), torch.device(device))


# ---- algorithmic FLOPs -----------------------------------------------------------
def fwd_flops_per_sample(Ts, Tt, D=384, FF=1536, k=3, enc_dw=7, dec_dw=31, blocks=(4, 4), postnet=(5, 256, 5), mel=80):
    S = Ts + Tt
    per_block = lambda kdw: (2 * (2 * S * (k * D) * FF + 2 * S * (k * FF) * D) + 4 * 2 * S * D * D + 3 * 2 * S * S * D
                             + 2 * S * D * 2 * D + 2 * S * D * kdw + 2 * S * D * D)
    f = blocks[0] * per_block(enc_dw) + blocks[1] * per_block(dec_dw)
    f += 2 * Ts * mel * D + 2 * Ts * D * mel
    n, ch, kf = postnet
    f += 2 * Ts * kf * (mel * ch + (n - 2) * ch * ch + ch * mel)
    return f


# ---- clocks -----------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# kernels launched by one C-ABI call (for gpu_launches)
KERNELS_PER_CALL = {"a3t_layernorm_bwd": 1, "a3t_colsum": 1, "a3t_mask_input_bwd": 2, "a3t_bn_stats": 2,
                    "a3t_bn_act_bwd": 3, "a3t_glu_dwconv_bwd": 2, "a3t_masked_l1_fwd": 2, "a3t_grad_sqnorm": 2,
                    "a3t_adam_step": 2, "a3t_stft_logmel": 2}


def _count_launches(fn):
    """Run fn() once counting kernel launches and GEMM calls issued through the C-ABI."""
    from a3t_b200 import _lib

    counts = {"kernels": 0, "gemm": 0}
    orig = _lib.call

    def counting(name, *a):
        rc = orig(name, *a)
        if name not in _lib._PLAIN_INT:
            counts["kernels"] += KERNELS_PER_CALL.get(name, 1)
            if name == "a3t_gemm":
                counts["gemm"] += 1
        return rc

    import a3t_b200.backend as bk
    import a3t_b200.trainer as tr

    _lib.call = bk.call = counting
    tr._lib.call = counting
    try:
        out = fn()
    finally:
        _lib.call = bk.call = orig
        tr._lib.call = orig
    return out, counts


def _time_gemms(fn):
    """One eager step with a CUDA-event pair around every a3t_gemm call (on the launching stream):
    returns (total GEMM ms, total tensor-core-eligible GEMM FLOPs, n calls)."""
    from a3t_b200 import _lib
    import a3t_b200.backend as bk

    orig = _lib.call
    evs = []

    def timing(name, *a):
        if name in ("a3t_relpos_attn_fwd", "a3t_relpos_attn_bwd"):
            # fused attention: QK^T + PV (fwd), QK^T + dO V^T + dS K (bwd) per (b, h); operands in, ctx / lse (fwd) or the
            # three (B,H,S,S) operands of the remaining contractions (bwd) out
            fwd = name.endswith("fwd")
            Bn, H, S, D = (a[6:10] if fwd else a[12:16])
            dk = D // H
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = orig(name, *a)
            e1.record()
            fl = (4.0 if fwd else 6.0) * Bn * H * S * S * dk
            by = 2.0 * Bn * H * S * S * (1 if fwd else 4) + 2.0 * Bn * S * D * (4 if fwd else 7)
            evs.append((e0, e1, fl, by, "fused rel-pos attention (scores on chip)"))
            return rc
        if name != "a3t_gemm":
            return orig(name, *a)
        d = a[0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = orig(name, *a)
        e1.record()
        nb = d.batch1 * d.batch2
        cs = 2 if d.dtype_c == _lib.A3T_BF16 else 4
        if d.mode == _lib.GEMM_CONV:      # activations once (taps re-read them from L2), packed weight, output
            by = 2 * d.M * d.cin + 2 * d.N * d.K + cs * d.M * d.N
        elif d.mode == _lib.GEMM_WGRAD:   # dy (K x M), x (K x cin), fp32 dW
            by = 2 * d.K * d.M + 2 * d.K * d.cin + 4 * d.M * d.N
        else:
            by = nb * (2 * d.M * d.K + 2 * d.N * d.K + cs * d.M * d.N)
        if nb > 1:
            cls = "attention (batched BD_raw = (q+v) p^T, dV, dK, d(q+v), dp)"
        elif d.taps == 3:
            cls = "ffn conv k3 (fwd, dgrad, wgrad)"
        elif d.taps == 1:
            cls = "projections / pointwise / head (k1)"
        else:
            cls = "postnet conv k5"
        evs.append((e0, e1, 2.0 * d.M * d.N * d.K * nb, float(by), cls))
        return rc

    _lib.call = bk.call = timing
    try:
        # keep the GPU behind the host for the whole instrumented step: with an idle GPU an event pair would also
        # span the host's launch latency (tens of us per call) and under-report every short kernel
        torch.cuda._sleep(int(1.9e9 * 0.06))
        fn()
        torch.cuda.synchronize()
    finally:
        _lib.call = bk.call = orig
    ms = sum(e0.elapsed_time(e1) for e0, e1, _, _, _ in evs)
    by_class = {}
    for e0, e1, f, _, cls in evs:
        a = by_class.setdefault(cls, [0, 0.0, 0.0])
        a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += f
    return ms, sum(f for _, _, f, _, _ in evs), len(evs), sum(b for _, _, _, b, _ in evs), by_class


def _clean_thread_env():
    """torchrun exports OMP_NUM_THREADS=1 to its children: a CPU baseline timed under it is noise."""
    env = dict(os.environ)
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        env.pop(k, None)
    return env


def run_reference(args):
    """CPU arm: the REFERENCE's own modules (oracle/_ref: a verbatim snapshot of the files the path imports, made by
    oracle/build_ref.py) -- MLMTask.build_model with conf/fsp2_conformer.yaml, the reference collate helpers for the
    batch, `loss = model(**batch)[0]; loss.backward()` in train mode (dropout on), fp32, all host threads -- on a
    bounded sample of the cfg2 workload (B = --cpu-batch of the 16 utterances; the model and frame counts are the
    full ones).  Falls back to the oracle port of the same graph when the snapshot is absent."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    omp = os.environ.get("OMP_NUM_THREADS")
    if omp is not None and int(omp) < cores and os.environ.get("A3T_REF_CHILD") != "1":
        env = _clean_thread_env()
        env["A3T_REF_CHILD"] = "1"
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
            env.pop(k, None)
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env, capture_output=True, text=True)
        sys.stdout.write(r.stdout)
        sys.stderr.write(r.stderr[-2000:])
        return
    torch.set_num_threads(cores)
    Bs = args.cpu_batch
    from oracle import build_ref

    if build_ref.available():
        os.environ["A3T_REFERENCE_ROOT"] = build_ref.OUT
        from oracle import ref_harness as R

        R.REFERENCE_ROOT = build_ref.OUT
        conf = R.model_conf("paper")
        m = R.build_reference_model(conf, seed=0)
        with torch.no_grad():  # as the GPU arm: BatchNorm gains 1 (xavier init zeroes them) so every layer does real work
            for n, p in m.named_parameters():
                if p.dim() == 1 and n.endswith("weight"):
                    p.fill_(1.0)
        batch, _ = R.synthetic_batch(Bs, args.frames, args.phones, seed=0)
        kind = "reference"

        def step():
            loss, _ = R.reference_step(m, batch, train=True)
            return float(loss)
    else:
        from a3t_b200 import graph
        from a3t_b200.model import build_model
        from oracle.oracle_backend import OracleBackend

        enc, dec, mc = paper_conf()
        torch.manual_seed(0)
        m = build_model(enc, dec, mc)
        hb = expand_on_host(synthetic_batch_host(Bs, args.frames, args.phones, seed=0))
        ops = OracleBackend(seed=1, autograd=True)
        P = {n: p for n, p in m.named_parameters()}
        P.update({n: b for n, b in m.named_buffers()})
        kind = "port"

        def step():
            for p in m.parameters():
                p.grad = None
            loss, _, _, _ = graph.forward(ops, P, graph.WeightCache(), m.cfg, hb, True, True)
            loss.backward()
            return float(loss)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = Bs * args.frames * args.steps / dt
    sample = (f"B={Bs} of {args.batch} utterances x Ts={args.frames}/Tt={args.phones} (full model, full sequence lengths; "
              f"frames/s is batch-size independent on a CPU), {args.steps} step(s), fwd+bwd, fp32, dropout on")
    print(json.dumps({
        "impl": "reference", "metric": f"mel-frames/sec training (VCTK A3T Conformer {args.config})", "value": fps,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: VCTK paper Conformer (4+4 blocks, D=384, H=2, FF=1536 k3, dw 7/31, postnet 5x256), "
                               f"Ts={args.frames}, Tt={args.phones}, train step fwd+bwd, dropout on; CPU sample of B={Bs} of the "
                               f"{args.batch} utterances per step (same model, same sequence lengths)", "batch": Bs},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline_leg(args):
    """Bounded CPU sample on rank 0 (reported beside the GPU number); own process, clean thread environment."""
    env = _clean_thread_env()
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-batch", str(args.cpu_batch), "--frames", str(args.frames), "--phones", str(args.phones)],
                       capture_output=True, text=True, timeout=900, env=env)
    for line in r.stdout.strip().splitlines()[::-1]:
        try:
            return json.loads(line)["cpu_baseline"]
        except Exception:
            continue
    return {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: " + r.stderr[-300:]}


def parity_object(dev):
    """Measured deviation of the benchmarked bf16 / tcgen05 path from the fp32 REFERENCE on the paper-width fixture
    (tests/golden/model_d384.pt: D=384, 1+1 blocks, B=2, Ts=1024, Tt=128; produced by running the reference).
    A checker, not the product path."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import _d384
        from test_gpu_parity_r2 import bf16_parity_numbers

        fx, b = _d384.load(os.path.join(ROOT, "tests", "golden"))
        r = bf16_parity_numbers(fx, b)
        r["fixture"] = "tests/golden/model_d384.pt (reference-generated), IMPL_TC, dropout 0"
        r["fp32_gate"] = "fp32 kernels meet |dloss| <= 1e-4 |loss| on the same fixture (tests/test_gpu_parity_r2.py)"
        return r
    except Exception as e:
        return {"error": repr(e)[:200]}


def run_cfg5(args):
    """BASELINE configs[4]: speech-editing inference -- masked-mel infill of 32 utterances (Ts=1024, Tt=128, span [384,640))
    in ONE batched call of the model + ParallelWaveGAN on 32 x 1024 frames -> 32 x 307 200 samples."""
    from a3t_b200 import _lib
    from a3t_b200.model import build_model
    from a3t_b200.vocoder import ParallelWaveGANGenerator

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    enc, dec, mc = paper_conf()
    torch.manual_seed(0)
    model = build_model(enc, dec, mc, act_dtype=torch.bfloat16)
    model.gemm_impl = _lib.IMPL_TC
    model = model.to(dev).eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1 and n.endswith("weight"):
                p.fill_(1.0)
    B, Ts, Tt, hop, fs = 32, 1024, 128, 300, 24000
    host = synthetic_batch_host(B, Ts, Tt, seed=0)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    gen = ParallelWaveGANGenerator(upsample_params={"upsample_scales": [4, 5, 3, 5]}).to(dev).eval()
    z = torch.randn(B, 1, Ts * hop, device=dev)
    span = [[384, 640]] * B

    def infill():
        b = device_batch(pinned, dev)
        b["masked_position"] = torch.zeros(B, Ts, dtype=torch.bool, device=dev)
        b["masked_position"][:, 384:640] = True
        return model.inference_batch(**b, span_boundary=span)

    def one():
        mel = infill()                                   # (B, Ts, 80): original frames outside the span, generated inside
        return gen.generate(mel.transpose(1, 2).contiguous(), z)

    W = max(args.warmup, 3)
    for _ in range(W):
        wav = one()
    torch.cuda.synchronize()
    clocks = ClockSampler(0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_inf = t_all = 0.0
    for _ in range(args.steps):
        ev[0].record()
        mel = infill()
        ev[1].record()
        wav = gen.generate(mel.transpose(1, 2).contiguous(), z)
        ev[2].record()
        torch.cuda.synchronize()
        t_inf += ev[0].elapsed_time(ev[1])
        t_all += ev[0].elapsed_time(ev[2])
    clk = clocks.stop()
    ms = t_all / args.steps
    audio_s = B * Ts * hop / fs
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    print(json.dumps({
        "metric": "speech-editing inference: masked-mel infill + ParallelWaveGAN, real-time factor", "value": ms / 1e3 / audio_s,
        "unit": "RTF (s compute / s audio)", "n_gpus": 1, "steps": args.steps, "warmup": W, "ms_per_step": ms,
        "higher_is_better": False, "scaling": "replicas only", "vs_baseline": None, "dtype": "bf16 (model) / f32 (vocoder)",
        "data": "synthetic",
        "config": {"workload": "cfg5: 32 utterances x Ts=1024 / Tt=128, span [384,640), batched infill + PWG 30 layers "
                               "-> 32 x 307200 samples", "infill_ms": t_inf / args.steps, "vocoder_ms": (t_all - t_inf) / args.steps,
                   "utterances_per_s": B / (ms / 1e3), "audio_seconds_per_s": audio_s / (ms / 1e3)},
        "e2e": {"value": ms / 1e3 / audio_s, "unit": "RTF", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 0},
        "clocks": clk}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="a3t_b200", choices=["a3t_b200", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--phones", type=int, default=128)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--cpu-batch", type=int, default=4)
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg4", "cfg5"],
                    help="cfg2: B=16, Ts=1024, Tt=128 per GPU (headline); cfg4: B=8 per GPU (64 over 8 GPUs), Ts=1500, Tt=192; "
                         "cfg5: batched infill + vocoder inference (own metric)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="seconds of untimed back-to-back steps before the timed region")
    ap.add_argument("--bucket-mb", type=float, default=0.0,
                    help="N>1: 0 = one all-reduce after the backward sweep (default); > 0 = minimum size of the gradient ranges "
                         "handed to NCCL while the sweep is still running (profiles/r02_scaling_overlap.md)")
    ap.add_argument("--diag-no-exchange", action="store_true",
                    help="diagnostic, N>1: run the ranks side by side WITHOUT the gradient all-reduce (not a valid bench line)")
    ap.add_argument("--diag-tiny-exchange", action="store_true",
                    help="diagnostic, N>1: no gradient exchange but a 16-byte all-reduce per step (cost of lock step alone)")
    ap.add_argument("--exchange-ctas", type=int, default=0, help="N>1: NCCL max_ctas of the gradient-exchange communicator (0 = NCCL default)")
    ap.add_argument("--exchange-last-full", action="store_true",
                    help="N>1 with --bucket-mb: the last gradient range (nothing left to overlap) uses the default communicator")
    ap.add_argument("--kernel-trace", default=None,
                    help="after the timed region: CUPTI kernel timeline (name, start, duration, stream, grid) of 2 steps on rank 0 -> CSV")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the frontend / vocoder side measurements")
    args = ap.parse_args()
    if args.config == "cfg4":
        args.batch, args.frames, args.phones = 8, 1500, 192
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "cfg5":
        return run_cfg5(args)

    import torch.distributed as dist
    from a3t_b200 import _lib
    from a3t_b200.model import build_model
    from a3t_b200.trainer import DataParallelTrainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    enc, dec, mc = paper_conf()
    torch.manual_seed(0)
    act = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    model = build_model(enc, dec, mc, act_dtype=act)
    model.gemm_impl = _lib.IMPL_TC   # a GEMM that does not qualify for the tcgen05 kernel is an error, never a silent fallback
    model = model.to(dev).train()
    with torch.no_grad():  # xavier init zeroes BatchNorm gamma: give the conv module / postnet real work
        for n, p in model.named_parameters():
            if p.dim() == 1 and n.endswith("weight"):
                p.fill_(1.0)
    trainer = DataParallelTrainer(model, bucket_bytes=int(args.bucket_mb * (1 << 20)), exchange_max_ctas=args.exchange_ctas,
                                  last_range_full_speed=args.exchange_last_full)
    if args.diag_no_exchange or args.diag_tiny_exchange:
        trainer.world = 1
    if args.diag_tiny_exchange:  # gradients stay local; one 16-byte all-reduce per step keeps the ranks in lock step
        _step = trainer.step

        def _step_sync(b):
            r = _step(b)
            dist.all_reduce(trainer.stats)
            return r

        trainer.step = _step_sync
    B, Ts, Tt = args.batch, args.frames, args.phones
    host = synthetic_batch_host(B, Ts, Tt, seed=rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    static = device_batch(pinned, dev)
    torch.cuda.synchronize()

    # ---- warm-up (eager), then capture the step in a CUDA graph -------------------------------
    _, counts = _count_launches(lambda: trainer.step(static))
    for _ in range(W - 1):
        trainer.step(static)
    torch.cuda.synchronize()
    graph_obj, used_graph, graph_err = None, False, None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                trainer.step(static)
                graph_obj = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph_obj, stream=side):
                    stats_out = trainer.step(static)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            graph_obj.replay()
            torch.cuda.synchronize()
            used_graph = True
        except Exception as e:  # report, then measure eagerly
            graph_err = repr(e)[:200]
            graph_obj = None
            torch.cuda.synchronize()

    def one_step():
        if used_graph:
            graph_obj.replay()
        else:
            trainer.step(static)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- at least 2 s of back-to-back steps first: the timed region then runs at the clocks / power state of a long
    # job (the "sustained" tensor peak is the right denominator), not in the first-second burst
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < args.sustain_s:
        for _ in range(10):
            one_step()
        torch.cuda.synchronize()
    # ---- timed region: K steps, device-timed, max over ranks -----------------------------------
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if clocks else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    frames = world * B * Ts * args.steps
    value = frames / (ms / 1e3)

    if args.kernel_trace:  # diagnostic: do the collectives overlap the backward kernels?
        barrier()
        if rank == 0:
            from torch.profiler import ProfilerActivity, profile

            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(2):
                    one_step()
                torch.cuda.synchronize()
            tmp = args.kernel_trace + ".json"
            prof.export_chrome_trace(tmp)
            ev = [e for e in json.load(open(tmp))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
            ev.sort(key=lambda e: e["ts"])
            t0 = ev[0]["ts"] if ev else 0
            with open(args.kernel_trace, "w") as f:
                f.write("start_us,dur_us,stream,grid,block,name\n")
                for e in ev:
                    a = e.get("args", {})
                    f.write(f"{e['ts'] - t0:.3f},{e['dur']:.3f},{a.get('stream')},\"{a.get('grid')}\",\"{a.get('block')}\",\"{e['name'][:90]}\"\n")
            os.remove(tmp)
        else:
            for _ in range(2):
                one_step()
            torch.cuda.synchronize()
        barrier()

    # ---- e2e: host (pinned) buffers in, loss out, through the public trainer API ---------------
    barrier()
    e0.record()
    for _ in range(args.steps):
        fresh = device_batch(pinned, dev)
        for k, v in fresh.items():
            static[k].copy_(v)
        one_step()
        stats_host = trainer.stats.to("cpu", non_blocking=False)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = frames / (float(t) / 1e3)
    loss_now = float(stats_host[0] / stats_host[2])

    # ---- roofline of the dominant kernel (GEMM) from one instrumented eager step ---------------
    gemm_ms, gemm_flops, n_gemm, gemm_bytes, gemm_classes = _time_gemms(lambda: trainer.step(static))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic, traffic_src = None, None
    try:  # DRAM bytes per GEMM launch from the committed ncu capture of the same step (profiles/)
        tn = [f for f in ("r02_gemm_traffic.json", "r01_gemm_traffic.json") if os.path.exists(os.path.join(ROOT, "profiles", f))][0]
        tj = json.load(open(os.path.join(ROOT, "profiles", tn)))
        traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained", 1590.0 * 0.88)
    peak_burst = peaks.get("bf16_tflops", 1590.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (timed after >= 2 s of steps)" if peaks else "fallback"
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    alg_flops_step = 3.0 * B * fwd_flops_per_sample(Ts, Tt)

    aux = None
    if rank == 0 and not args.no_aux:
        # the two bandwidth-framed satellites of the path (north_star): STFT->log-mel frontend and PWG generator
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_aux
            aux = bench_aux.measure(pwg_batch=8, peaks=peaks)
        except Exception as e:  # reported, never hidden
            aux = {"error": repr(e)[:200]}
    if rank == 0:
        # the CPU baseline is a property of the box, not of N: measured at N = 1 only (with other ranks alive their
        # NCCL / barrier threads share the cores and the number is noise)
        cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline_leg(args)
        out = {
            "metric": f"mel-frames/sec training (VCTK A3T Conformer {args.config})", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"{args.config}: VCTK paper Conformer (4+4 blocks, D=384, H=2, FF=1536 k3, dw 7/31, postnet 5x256), "
                                   f"B={B}/GPU, Ts={Ts}, Tt={Tt}, train step fwd+bwd+allreduce+clip/Adam/Noam, dropout on",
                       "global_batch": B * world, "parallelism": f"dp{world}", "cuda_graph": used_graph,
                       "graph_error": graph_err, **({"diag": "NO gradient exchange (diagnostic run)"} if (args.diag_no_exchange or args.diag_tiny_exchange) else {}),
                       "grad_exchange_ranges": len(trainer.exchange_ranges) if world > 1 else 0, "l2": "activations per step (GBs) exceed the 126 MB L2; no explicit flush",
                       "loss": loss_now, "alg_tflop_per_step_per_gpu": alg_flops_step / 1e12,
                       "step_tensor_frac_of_peak": alg_flops_step / (ms / args.steps / 1e3) / 1e12 / peak},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None, "frac_of_burst_peak": achieved / peak_burst, "traffic": traffic, "traffic_unit": "B/launch (DRAM read+write, mean over the step's GEMM launches)",
                         "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": gemm_bytes / max(n_gemm, 1),
                         "algorithmic_flop_per_launch": gemm_flops / max(n_gemm, 1),
                         "by_class": {k: {"launches": v[0], "ms": v[1], "tflops": v[2] / (v[1] / 1e3) / 1e12 if v[1] > 0 else None,
                                          "frac": (v[2] / (v[1] / 1e3) / 1e12 / peak) if v[1] > 0 and peak else None}
                                      for k, v in gemm_classes.items()},
                         "kernel": "tc::gemm_tc_kernel (a3t_gemm) + fa::attn_{fwd,bwd}_kernel: all dense contractions of the step",
                         "launches": n_gemm,
                         "how": "CUDA-event pair around every a3t_gemm call of one eager step (GPU kept busy ahead of the host so the pairs hold kernel time only); FLOPs = 2*M*N*K*batch per call",
                         "peak_source": peak_src},
            "cpu_baseline": cpu, "parity": None if args.no_parity else parity_object(dev),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16},
            "gpu_launches": counts["kernels"] * args.steps, "clocks": clk, "aux": aux,
        }
        print(json.dumps(out))
    if world > 1:
        # NCCL kernels captured in a CUDA graph keep the communicator busy at teardown: destroy_process_group()
        # was observed to hang after the result line had been printed.  Release the graph, agree that every
        # rank is done, flush, and leave without running the NCCL / graph destructors.
        graph_obj = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
