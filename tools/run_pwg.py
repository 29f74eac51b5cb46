import sys, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from a3t_b200.vocoder import ParallelWaveGANGenerator
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
gen = ParallelWaveGANGenerator(upsample_params={"upsample_scales": [4, 5, 3, 5]}).cuda().eval()
c = torch.randn(B, 80, 1024, device="cuda"); z = torch.randn(B, 1, 1024 * 300, device="cuda")
y = gen.generate(c, z)
torch.cuda.synchronize()
print(y.shape)
