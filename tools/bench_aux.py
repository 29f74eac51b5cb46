#!/usr/bin/env python
"""Device-timed frontend (STFT->log-mel) and ParallelWaveGAN generator numbers for bench.py / DESIGN.md.
  frontend: B=16 utterances of T=1024 frames (fs 24 k, n_fft 2048, win 1200, hop 300), fp32
  vocoder : B utterances x T=1024 frames -> T*300 samples, 30 residual blocks (parallel_wavegan.v1 shape)
Prints one JSON object."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def measure(pwg_batch=8, peaks=None):
    from a3t_b200.frontend import LogMelFbank
    from a3t_b200.vocoder import ParallelWaveGANGenerator
    peaks = peaks or {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    dev = torch.device("cuda")
    out = {}
    # ---- frontend
    B, T, hop = 16, 1024, 300
    fe = LogMelFbank(fs=24000, n_fft=2048, win_length=1200, hop_length=hop, fmin=80, fmax=7600, n_mels=80).to(dev)
    wav = torch.randn(B, (T - 1) * hop, device=dev) * 0.1
    ms = timeit(lambda: fe(wav))
    frames = B * T
    byts = frames * (hop * 4 + 80 * 4)
    out["frontend"] = {"workload": f"B={B} x T={T} frames, fs 24 kHz n_fft 2048 win 1200 hop 300, 80 mels, fp32",
                       "ms": ms, "frames_per_s": frames / ms * 1e3, "alg_bytes_per_frame": hop * 4 + 80 * 4,
                       "achieved_gbs": byts / ms / 1e6, "hbm_peak_gbs": hbm, "frac_of_hbm": byts / ms / 1e6 / hbm,
                       "bound": "issue / latency (register FFT: ~70 kFLOP per 1.5 kB frame), not HBM"}
    # issue roofline from the counter that names the bound: ncu smsp__inst_executed.sum = 56.9 M warp instructions for
    # these 16 384 frames (profiles/r02_frontend_pwg_ncu.md) at 4 issue slots per clock per SM
    props = torch.cuda.get_device_properties(dev)
    wi = 56896976 / 16384.0
    issue_peak = props.multi_processor_count * 4 * 1.965e9 / wi
    out["frontend"].update({"warp_instructions_per_frame": wi, "issue_peak_frames_per_s": issue_peak,
                            "frac_of_issue": frames / ms * 1e3 / issue_peak})
    # ---- vocoder
    gen = ParallelWaveGANGenerator(upsample_params={"upsample_scales": [4, 5, 3, 5]}).to(dev).eval()
    c = torch.randn(pwg_batch, 80, T, device=dev)
    z = torch.randn(pwg_batch, 1, T * hop, device=dev)
    ms = timeit(lambda: gen.generate(c, z), n=3, warm=1)
    audio_s = pwg_batch * T * hop / 24000.0
    samples = pwg_batch * T * hop
    # per residual block and sample: x hi/lo planes in 256 B + out 256 B, aux planes 320 B, fp32 skip read + write 512 B
    bps = 30 * 1344
    out["pwg"] = {"workload": f"{pwg_batch} utterances x {T} frames -> {T*hop} samples each, 30 residual blocks, "
                              "split-fp16 tensor-core blocks (3 passes, fp32 accumulation), fp32 in / out",
                  "ms": ms, "rtf": ms / 1e3 / audio_s, "audio_seconds_per_s": audio_s / (ms / 1e3),
                  "alg_bytes_per_sample": bps, "achieved_gbs": samples * bps / ms / 1e6,
                  "hbm_peak_gbs": hbm, "frac_of_hbm": samples * bps / ms / 1e6 / hbm,
                  "gflop_per_audio_s": 62.0, "achieved_tflops_algorithmic": 62.0 * audio_s / ms}
    return out


if __name__ == "__main__":
    pk = {}
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    print(json.dumps(measure(int(sys.argv[1]) if len(sys.argv) > 1 else 8, pk)))
