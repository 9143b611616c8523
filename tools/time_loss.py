import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import denoising_diffusion_deep_fake_b200 as d3
dev = "cuda:0"
x = torch.randn(256, 3, 64, 64, device=dev).clamp(-1, 1)
p = (x + 0.3 * torch.randn_like(x)).requires_grad_(True)
crit = d3.MseStructuralSimilarityLoss(-1.0, 1.0)
for _ in range(3):
    crit(p, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    l = crit(p, x)
e1.record(); torch.cuda.synchronize()
print(f"criterion forward (value+grad kernel): {e0.elapsed_time(e1)/50*1000:.1f} us")
