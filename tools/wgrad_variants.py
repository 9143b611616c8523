"""Time the weight-gradient ops of one training plan under the current D3FK_* environment (steady state)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from denoising_diffusion_deep_fake_b200 import _lib
from denoising_diffusion_deep_fake_b200.train import DenoiserModule
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda:0")
mod = DenoiserModule(encoder_name="resnet34", learning_rate=0.02, noise_exponential_sampling_lambda=5,
                     cosine_scheduler_max_epoch=100, precision="bf16").to(dev).train()
mod.configure_optimizers(fused=True)
x = torch.randn(256, 3, 64, 64, device=dev).clamp(-1, 1)
for _ in range(2):
    mod.training_step(x)
torch.cuda.synchronize()
plan = next(p for plans in mod.model._plans.values() for p in plans if p.training)
s = torch.cuda.current_stream().cuda_stream
import ctypes
kind = _lib.OP_WGRAD if len(sys.argv) < 2 or sys.argv[1] == "wgrad" else _lib.OP_CONV
tot = 0.0
seen = {}
for seg in plan.bwd_segments:
    for op in seg:
        if op.kind != kind and not (kind == _lib.OP_CONV and op.kind == _lib.OP_CONV_BN):
            continue
        p = _lib.op_params(op)
        if op.kind == _lib.OP_CONV_BN:
            p = p.conv
        key = (p.B * p.Ho * p.Wo, p.Cout, p.kh * p.kw * (p.c0 + p.c1), p.stride, p.up0)
        copies = []
        for _ in range(20):
            c = _lib.Op(); ctypes.memmove(ctypes.byref(c), ctypes.byref(op), ctypes.sizeof(_lib.Op)); copies.append(c)
        ol = _lib.OpList(copies)
        ol.run(s); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ol.run(s); e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 20
        tot += t
        seen.setdefault(key, []).append(t)
print(f"total {tot*1000:.1f} us  env=" + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("D3FK_")))
for k, v in sorted(seen.items(), key=lambda kv: -sum(kv[1])):
    print(f"  M={k[0]:8d} N={k[1]:4d} K={k[2]:5d} s{k[3]} up{k[4]}  x{len(v):2d}  {1000*sum(v)/len(v):7.1f} us")
