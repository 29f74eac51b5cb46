class ComplexTensor:  # stub: import-only in espnet2/layers/{stft,log_mel}.py
    pass
