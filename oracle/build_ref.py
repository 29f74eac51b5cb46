"""Recipe for `oracle/_ref/`: a verbatim snapshot of the REFERENCE modules this path imports.  TEST / BASELINE
INFRASTRUCTURE ONLY -- nothing under a3t_b200/ imports it.

The reference is pure Python, so there is nothing to compile: "building" it = copying, from where they lie under
/root/reference, exactly the files that `import espnet2.tasks.mlm`, the in-tree ParallelWaveGAN generator and the
recipe YAML pull in (the import closure, ~150 files), byte for byte, into `oracle/_ref/` together with a MANIFEST
of their sha256.  `oracle/_ref/` is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so
it travels to the GPU box like a built `.so`; there `bench.py --impl reference` / `cpu_baseline` run the reference's
OWN modules on the host cores (`cpu_baseline.kind = "reference"`).  Without the snapshot they fall back to the
oracle port (`kind = "port"`).

    python -m oracle.build_ref          # run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
EXTRA = ["egs2/vctk/sedit/conf/fsp2_conformer.yaml", "espnet/version.txt"]


def build(reference_root: str = "/root/reference") -> str:
    if not os.path.isdir(os.path.join(reference_root, "espnet2")):
        raise RuntimeError(f"no reference tree at {reference_root}")
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import ref_harness as R

    os.environ["A3T_REFERENCE_ROOT"] = reference_root
    R.REFERENCE_ROOT = reference_root
    R._activate()
    import espnet2.tts.feats_extract.log_mel_fbank  # noqa: F401  (frontend baseline)
    # some reference modules are imported lazily inside functions (activation lookup, ...): exercise the path once
    # (tiny model step, frontend, vocoder) so that the closure below is complete
    import torch

    conf = R.model_conf("paper")
    for k in ("encoder_conf", "decoder_conf"):
        conf[k].update(num_blocks=1, attention_dim=32, attention_heads=2, linear_units=64)
    conf["model_conf"].update(postnet_chans=32)
    model = R.build_reference_model(conf, vocab=20)
    batch, _ = R.synthetic_batch(2, 40, 8, vocab=20, seed=0)
    R.reference_step(model, batch, train=True)
    from espnet2.gan_tts.parallel_wavegan import ParallelWaveGANGenerator
    from espnet2.tts.feats_extract.log_mel_fbank import LogMelFbank

    LogMelFbank(fs=24000, n_fft=2048, win_length=1200, hop_length=300, fmin=80, fmax=7600, n_mels=80)(torch.zeros(1, 3000))
    gen = ParallelWaveGANGenerator(layers=3, stacks=1, upsample_params={"upsample_scales": [4, 5, 3, 5]})
    gen.remove_weight_norm()
    with torch.no_grad():
        gen(torch.zeros(1, 80, 4), torch.zeros(1, 1, 1200))

    root = os.path.realpath(reference_root)
    files = set()
    for m in list(sys.modules.values()):
        f = getattr(m, "__file__", None)
        if f and os.path.realpath(f).startswith(root + os.sep):
            files.add(os.path.relpath(os.path.realpath(f), root))
    for rel in list(files):                       # package markers of every directory on the way
        d = os.path.dirname(rel)
        while d:
            init = os.path.join(d, "__init__.py")
            if os.path.exists(os.path.join(root, init)):
                files.add(init)
            d = os.path.dirname(d)
    files.update(EXTRA)
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    manifest = {}
    for rel in sorted(files):
        src, dst = os.path.join(root, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    json.dump({"source": reference_root, "files": manifest}, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=0)
    return OUT


def available() -> bool:
    return os.path.exists(os.path.join(OUT, "MANIFEST.json"))


if __name__ == "__main__":
    out = build()
    n = len(json.load(open(os.path.join(out, "MANIFEST.json")))["files"])
    print(f"snapshot of {n} reference files -> {out}")
