"""Is the training step host-bound?  Host enqueue time per step (no sync inside the loop) against the device time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from denoising_diffusion_deep_fake_b200.train import DenoiserModule
dev = torch.device("cuda:0")
mod = DenoiserModule(encoder_name="resnet34", learning_rate=0.02, noise_exponential_sampling_lambda=5,
                     cosine_scheduler_max_epoch=100, precision="bf16").to(dev).train()
mod.configure_optimizers(fused=True)
x = torch.randn(256, 3, 64, 64, device=dev).clamp(-1, 1)
for _ in range(5):
    mod.training_step(x)
torch.cuda.synchronize()
N = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(N):
    mod.training_step(x)
t1 = time.perf_counter()
e1.record()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / N:.3f} ms/step; device {e0.elapsed_time(e1) / N:.3f} ms/step; host wait at the end {1e3 * (t2 - t1):.2f} ms")
# split the host time: forward+loss, backward, rest
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    mod.training_step(x)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
