import sys, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from a3t_b200.frontend import LogMelFbank
fe = LogMelFbank(fs=24000, n_fft=2048, win_length=1200, hop_length=300, fmin=80, fmax=7600, n_mels=80).cuda()
wav = torch.randn(16, 1023 * 300, device="cuda") * 0.1
for _ in range(3):
    m, l = fe(wav)
torch.cuda.synchronize()
print(m.shape, float(m.mean()))
