"""Achieved HBM bandwidth of the bandwidth-bound kernels (algorithmic bytes / CUDA-event time) against the measured copy peak
in MEASURED_PEAKS.json.  Each kernel runs on buffers larger than the 126 MB L2 (or is rotated over several buffers)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import denoising_diffusion_deep_fake_b200 as d3
from denoising_diffusion_deep_fake_b200 import _lib
from denoising_diffusion_deep_fake_b200.functional import (q_sample, posterior_step_, adam_step_, frames_to_tensor,
                                                            tensor_to_frames, affine_q_sample, random_affine_inverse_maps)
dev = torch.device("cuda:0")
_lib.init(0)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
rows = []

def timed(name, fn, nbytes, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    gbs = nbytes / us / 1e3
    rows.append((name, nbytes / 1e6, us, gbs, gbs / peak))
    print(f"{name:58s} {nbytes / 1e6:9.1f} MB {us:9.1f} us {gbs:8.0f} GB/s  {100 * gbs / peak:5.1f} % of {peak:.0f}")

# q_sample / posterior at the sampling shape x a batch large enough to leave L2 (B=1024 @128x128: 201 MB per tensor)
B, H = 1024, 128
x = torch.randn(B, 3, H, H, device=dev)
noise = torch.randn_like(x)
y = torch.rand(B, 1, 1, 1, device=dev)
n = x.numel()
timed("q_sample, Philox noise in-kernel (read x, write out)", lambda: q_sample(x, 5.0, seed=1), 8 * n)
timed("q_sample, noise given (read x, noise, write out)", lambda: q_sample(x, 5.0, noise=noise, y=y), 12 * n)
maps = random_affine_inverse_maps(B, H, H).to(dev)
timed("affine_q_sample, Philox (read x, write augmented + noisy)", lambda: affine_q_sample(x, maps, 5.0, seed=1), 12 * n)
h = torch.randn_like(x)
timed("posterior_step eta=0 (read x, x0_hat, write x)", lambda: posterior_step_(x, h, 0.5, 0.4, eta=0.0), 12 * n)
timed("posterior_step eta=1, Philox z", lambda: posterior_step_(x, h, 0.5, 0.4, eta=1.0, seed=3), 12 * n)
del noise, h
# Adam over the model's 24.4 M parameters (28 B / parameter) and over 4x that (leaves L2)
for scale in (1, 4):
    npar = 24436659 // 4 * 4 * scale
    p, g, m, v = (torch.randn(npar, device=dev) * 0.01 for _ in range(4))
    v.abs_()
    timed(f"adam, {npar / 1e6:.1f} M parameters (28 B each)", lambda: adam_step_(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, 3), 28 * npar)
    del p, g, m, v
# video frames: 64 frames of 448x448 (the reference's working size) and 256 of them
for N in (64, 256):
    fr = torch.randint(0, 256, (N, 448, 448, 3), dtype=torch.uint8, device=dev)
    t = frames_to_tensor(fr, [0.5] * 3, [0.5] * 3)
    npx = N * 448 * 448
    timed(f"frames_to_tensor, {N} x 448x448 (3 B in, 12 B out / px)", lambda: frames_to_tensor(fr, [0.5] * 3, [0.5] * 3), 15 * npx)
    timed(f"tensor_to_frames, {N} x 448x448 (12 B in, 3 B out / px)", lambda: tensor_to_frames(t, [0.5] * 3, [0.5] * 3), 15 * npx)
    del fr, t
# BatchNorm kernels on the biggest layers of the training plan (B=256 @64x64): through the per-op profiler
from denoising_diffusion_deep_fake_b200.train import DenoiserModule
mod = DenoiserModule(encoder_name="resnet34", learning_rate=0.02, noise_exponential_sampling_lambda=5,
                     cosine_scheduler_max_epoch=100, precision="bf16").to(dev).train()
mod.configure_optimizers(fused=True)
xb = torch.randn(256, 3, 64, 64, device=dev).clamp(-1, 1)
mod.training_step(xb)
torch.cuda.synchronize()
plan = next(p for plans in mod.model._plans.values() for p in plans if p.training)
import ctypes
s = torch.cuda.current_stream().cuda_stream
seen = set()
for op in [op for seg in plan.bwd_segments for op in seg] + list(plan.fwd_ops):
    if op.kind not in (_lib.OP_BN_APPLY, _lib.OP_BN_BWD_REDUCE, _lib.OP_BN_BWD_APPLY, _lib.OP_BN_BWD):
        continue
    p = _lib.op_params(op)
    key = (op.kind, p.count, p.C)
    if key in seen or p.count * p.C < 8 << 20:
        continue
    seen.add(key)
    c = _lib.Op(); ctypes.memmove(ctypes.byref(c), ctypes.byref(op), ctypes.sizeof(_lib.Op))
    ol = _lib.OpList([c])
    el = p.count * p.C
    if op.kind == _lib.OP_BN_APPLY:
        name, nbytes = "bn_apply (read raw, write act; bf16)", 4 * el
    elif op.kind == _lib.OP_BN_BWD_REDUCE:
        name, nbytes = "bn_bwd_reduce (read g, x, act; bf16)", 6 * el
    elif op.kind == _lib.OP_BN_BWD:
        name, nbytes = "bn_bwd reduce+apply (2 launches: 6 B + 8 B / element)", 14 * el
    else:
        name, nbytes = "bn_bwd_apply (read g, x, act, write dx; bf16)", 8 * el
    timed(f"{name} count={p.count} C={p.C}", lambda: ol.run(s), nbytes, reps=10)
json.dump([dict(kernel=r[0], mbytes=r[1], us=r[2], gbs=r[3], frac_of_hbm_peak=r[4]) for r in rows],
          open("gpurun_out/elementwise_bw.json", "w"), indent=1)
