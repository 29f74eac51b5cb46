"""a3t_b200 — Blackwell-native A3T (alignment-aware masked-mel pretraining) hot path.

Public surface (mirrors the reference's plug-in API for this path):
  model.MLMEncoder / MLMDecoder / ESPnetMLMEncAsDecoderModel / build_model
  frontend.LogMelFbank
  collate.MLMCollateFn / phones_masking / get_segment_pos
  vocoder.ParallelWaveGANGenerator / ParallelWaveGANPretrainedVocoder
  trainer.DataParallelTrainer (flat-buffer data parallel step with one NCCL all-reduce)
Everything computes through the C-ABI library `lib/liba3t_b200.so` (include/a3t_b200.h).
"""
__version__ = "0.1.0"
