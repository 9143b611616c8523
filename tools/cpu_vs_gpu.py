"""Is the training step launch-bound?  Host enqueue time per step vs device time per step."""
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
N = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(N):
    mod.training_step(x)
t1 = time.perf_counter(); e1.record()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/N:.3f} ms/step, device {e0.elapsed_time(e1)/N:.3f} ms/step, wall incl. drain {1e3*(t2-t0)/N:.3f} ms/step")
plan = next(p for plans in mod.model._plans.values() for p in plans if p.training)
s = torch.cuda.current_stream().cuda_stream
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    plan.fwd_ops.run(s)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"forward op list ({len(plan.fwd_ops)} ops): host {1e3*(t1-t0)/10:.3f} ms, wall {1e3*(t2-t0)/10:.3f} ms")
