"""Time the slab weight-gradient ops of the training plan (decoder tail) under the current library / knobs."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from denoising_diffusion_deep_fake_b200 import _lib
from denoising_diffusion_deep_fake_b200.train import DenoiserModule
dev = torch.device("cuda:0")
mod = DenoiserModule(encoder_name="resnet34", learning_rate=0.02, noise_exponential_sampling_lambda=5,
                     cosine_scheduler_max_epoch=100, precision="bf16").to(dev).train()
mod.configure_optimizers(fused=True)
x = torch.randn(256, 3, 64, 64, device=dev).clamp(-1, 1)
mod.training_step(x)
torch.cuda.synchronize()
plan = next(p for plans in mod.model._plans.values() for p in plans if p.training)
s = torch.cuda.current_stream().cuda_stream
tot = 0.0
for op in plan.bwd_segments[0]:
    if op.kind != _lib.OP_WGRAD:
        continue
    p = _lib.op_params(op)
    M, K = p.B * p.Ho * p.Wo, p.kh * p.kw * (p.c0 + p.c1)
    if M < 200000:
        continue
    copies = []
    for _ in range(10):
        c = _lib.Op(); ctypes.memmove(ctypes.byref(c), ctypes.byref(op), ctypes.sizeof(_lib.Op)); c.lane = 0
        copies.append(c)
    os.environ["D3FK_FORK_WGRAD"] = "0"
    ol = _lib.OpList(copies)
    ol.run(s); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ol.run(s); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10 * 1e3
    tot += t
    print(f"wgrad M={M} C={p.c0} N={p.Cout} K={K}: {t:.1f} us")
print(f"total {tot:.1f} us")
