"""Registration of the a3t_b200 classes into the reference's plug-in tables (the binding INTEGRATION.md
describes).  The reference has no FFI: its drop-in boundary for this path is the `ClassChoices` registry of
`espnet2/tasks/mlm.py` (:58-95) plus the model classes `MLMTask.build_model` instantiates (:416-435) and
`build_model_from_file` type-checks (:476-479).

    import a3t_b200.espnet_plugin as plug; plug.register()      # e.g. at the top of espnet2/bin/mlm_train.py

after which `MLMTask.build_model(args)` with `conf/fsp2_conformer.yaml` returns
`a3t_b200.model.ESPnetMLMEncAsDecoderModel` (an `AbsESPnetModel`: `abs_task.py:1097-1100` accepts it), and
`MLMTask.build_model_from_file` loads published checkpoints into it (same 363 state_dict keys; the
`encoder.embed` -> `encoder.speech_embed` rename of mlm.py:490-493 applies unchanged).
"""
from __future__ import annotations

import importlib
from typing import Optional


def espnet_base(modname: str, clsname: str) -> Optional[type]:
    """The reference's abstract base class when `espnet2` is importable (so isinstance checks of the
    reference's runtime hold), else None."""
    try:
        return getattr(importlib.import_module(modname), clsname)
    except Exception:
        return None


def register(collate: bool = False, vocoder: bool = True):
    """Patch `espnet2.tasks.mlm` in place.  collate=True also swaps `MLMCollateFn` for the device-side functor
    (only valid when the DataLoader runs in the training process: CUDA must not be touched from forked workers,
    SURVEY 8b)."""
    mlm = importlib.import_module("espnet2.tasks.mlm")
    from . import collate as collate_mod
    from . import frontend, model

    abs_model = espnet_base("espnet2.train.abs_espnet_model", "AbsESPnetModel")
    abs_feats = espnet_base("espnet2.tts.feats_extract.abs_feats_extract", "AbsFeatsExtract")
    # If a3t_b200 was imported before espnet2 became importable its classes derive from nn.Module only:
    # make them virtual subclasses of the (ABC) bases so the reference's isinstance checks still hold.
    if abs_model is not None and not issubclass(model.ESPnetMLMModel, abs_model):
        abs_model.register(model.ESPnetMLMModel)
    if abs_feats is not None and not issubclass(frontend.LogMelFbank, abs_feats):
        abs_feats.register(frontend.LogMelFbank)
    mlm.encoder_choices.classes["conformer"] = model.MLMEncoder              # mlm.py:77-84
    mlm.decoder_choices.classes["conformer"] = model.MLMDecoder              # mlm.py:86-95
    mlm.feats_extractor_choices.classes["fbank"] = frontend.LogMelFbank      # mlm.py:58-67
    mlm.ESPnetMLMModel = model.ESPnetMLMModel                                # type check at mlm.py:476
    mlm.ESPnetMLMEncAsDecoderModel = model.ESPnetMLMEncAsDecoderModel        # instantiated at mlm.py:427
    if hasattr(model, "ESPnetMLMTTSModel"):
        mlm.ESPnetMLMTTSModel = model.ESPnetMLMTTSModel                      # instantiated at mlm.py:417
    if collate:
        mlm.MLMCollateFn = collate_mod.MLMCollateFn                          # mlm.py:290
    if vocoder:
        try:
            tts = importlib.import_module("espnet2.tasks.tts")
            from . import vocoder as voc

            tts.ParallelWaveGANPretrainedVocoder = voc.ParallelWaveGANPretrainedVocoder   # tasks/tts.py:366-401
        except Exception:
            pass
    return mlm.MLMTask
