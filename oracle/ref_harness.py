"""Import the REFERENCE (richardbaihe/a3t under /root/reference) as pure PyTorch.  TEST INFRASTRUCTURE ONLY.

Works only in the build container (the GPU box has no /root/reference): used by
`oracle/make_golden.py` to generate the fixtures under `tests/golden/` and by the CPU tests that
compare against the live reference when it is present.  Absent third-party imports are satisfied
by the stub modules in `oracle/shims/`.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("A3T_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "espnet2"))


_activated = False


def _activate():
    """Put the reference and the stub packages on sys.path and import the reference's task module once.

    Under pytest the REAL `typeguard` (4.x, loaded by its pytest plugin) is already in sys.modules; ESPnet's
    `str = None` defaults do not survive it, so the permissive shim is swapped in for the duration
    of the reference's imports (its modules bind `check_argument_types` at import time) and the real one restored."""
    global _activated
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    if _activated:
        return
    import importlib.util

    real = sys.modules.get("typeguard")
    swapped = real is not None and not getattr(real, "_A3T_SHIM", False)
    if swapped:
        spec = importlib.util.spec_from_file_location("typeguard", os.path.join(_SHIMS, "typeguard", "__init__.py"))
        shim = importlib.util.module_from_spec(spec)
        sys.modules["typeguard"] = shim
        spec.loader.exec_module(shim)
    try:
        import espnet2.tasks.mlm  # noqa: F401  (pulls in every reference module the path uses)
        import espnet2.gan_tts.parallel_wavegan  # noqa: F401
    finally:
        if swapped:
            sys.modules["typeguard"] = real
    _activated = True


PAPER_YAML = "egs2/vctk/sedit/conf/fsp2_conformer.yaml"


def model_conf(name: str) -> dict:
    """cfg1 = SURVEY 8d plumbing config; paper = conf/fsp2_conformer.yaml as shipped."""
    _activate()
    import yaml

    y = yaml.safe_load(open(os.path.join(REFERENCE_ROOT, PAPER_YAML)))
    conf = dict(encoder_conf=dict(y["encoder_conf"]), decoder_conf=dict(y["decoder_conf"]), model_conf=dict(y["model_conf"]))
    if name == "cfg1":
        for k in ("encoder_conf", "decoder_conf"):
            conf[k].update(num_blocks=2, attention_dim=128, attention_heads=2, linear_units=512)
    elif name != "paper":
        raise ValueError(name)
    return conf


def build_reference_model(conf: dict, vocab: int = 73, dropout_zero: bool = False, seed: int = 0):
    """MLMTask.build_model (espnet2/tasks/mlm.py:329) -> ESPnetMLMEncAsDecoderModel."""
    _activate()
    from espnet2.tasks.mlm import MLMTask

    enc, dec, mc = dict(conf["encoder_conf"]), dict(conf["decoder_conf"]), dict(conf["model_conf"])
    if dropout_zero:
        for c in (enc, dec):
            c.update(dropout_rate=0.0, positional_dropout_rate=0.0, attention_dropout_rate=0.0)
    torch.manual_seed(seed)
    args = argparse.Namespace(
        token_list=["<blank>", "<unk>"] + [f"p{i}" for i in range(vocab - 3)] + ["<sos/eos>"], odim=80, input_size=80,
        feats_extract="fbank", feats_extract_conf={}, normalize=None, normalize_conf={}, use_scaled_pos_enc=False,
        encoder="conformer", encoder_conf=enc, decoder="conformer", decoder_conf=dec, model_conf=mc,
        init="xavier_uniform")
    model = MLMTask.build_model(args)
    if dropout_zero and model.postnet is not None:
        for m in model.postnet.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    return model


def randomize_degenerate_params(model, seed: int = 1):
    """`initialize(model,'xavier_uniform')` zeroes every 1-D parameter, BatchNorm gamma included
    (SURVEY Appendix B), which makes the conv module and the postnet output exactly 0.  Draw
    non-trivial values for all 1-D parameters and BatchNorm running statistics so that parity
    tests exercise every kernel."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() != 1:
                continue
            if n.endswith("weight"):  # every 1-D weight is a LayerNorm / BatchNorm gamma
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
        for n, b in model.named_buffers():
            if n.endswith("running_mean"):
                b.copy_(0.1 * torch.randn(b.shape, generator=g))
            elif n.endswith("running_var"):
                b.copy_(1.0 + 0.2 * torch.rand(b.shape, generator=g))


def synthetic_batch(B: int, Ts: int, Tt: int, vocab: int = 73, seed: int = 0, mlm_prob: float = 0.8,
                    mean_phn_span: int = 8, ragged: bool = False):
    """SURVEY 8d synthetic inputs, produced with the REFERENCE's own collate helpers
    (espnet2/train/collate_fn.py phones_masking / get_segment_pos)."""
    _activate()
    from espnet2.train.collate_fn import get_segment_pos, phones_masking
    from espnet.nets.pytorch_backend.nets_utils import make_non_pad_mask

    torch.manual_seed(seed)
    np.random.seed(seed)
    speech = torch.randn(B, Ts, 80)
    text = torch.randint(2, vocab - 1, (B, Tt))
    if ragged:
        slens = torch.tensor([Ts - (i * 7) % max(Ts // 3, 1) for i in range(B)])
        tlens = torch.tensor([Tt - (i * 3) % max(Tt // 3, 1) for i in range(B)])
        slens[0], tlens[0] = Ts, Tt
    else:
        slens = torch.full((B,), Ts)
        tlens = torch.full((B,), Tt)
    align_start = torch.zeros(B, Tt, dtype=torch.int32)
    align_end = torch.zeros(B, Tt, dtype=torch.int32)
    for b in range(B):
        L, n = int(tlens[b]), int(slens[b])
        edges = torch.floor(torch.linspace(0, n, L + 1)).int()
        align_start[b, :L] = edges[:-1]
        align_end[b, :L] = edges[1:]
        speech[b, n:] = 0.0
        text[b, L:] = 0
    speech_mask = make_non_pad_mask(slens.tolist(), speech[:, :, 0], length_dim=1).unsqueeze(-2)
    text_mask = make_non_pad_mask(tlens.tolist(), text, length_dim=1).unsqueeze(-2)
    masked_position, _ = phones_masking(speech, speech_mask, align_start, align_end, tlens, mlm_prob, mean_phn_span)
    sseg, tseg = get_segment_pos(speech, text, align_start, align_end, tlens, True)
    batch = dict(speech=speech, text=text, masked_position=masked_position, speech_mask=speech_mask,
                 text_mask=text_mask, speech_segment_pos=sseg, text_segment_pos=tseg,
                 speech_lengths=slens, text_lengths=tlens)
    aux = dict(align_start=align_start, align_end=align_end, align_lengths=tlens)
    return batch, aux


def reference_step(model, batch, train: bool = True):
    """loss (and grads when train) from the reference model on `batch`."""
    model.train(train)
    model.zero_grad(set_to_none=True)
    if train:
        loss, stats, weight = model(**batch)
        loss.backward()
        grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        return loss.detach(), grads
    with torch.no_grad():
        loss, stats, weight = model(**batch)
    return loss.detach(), {}
