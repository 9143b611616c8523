"""Where does an OP_CONV_BN's time go?  For every conv+BN op of the training forward: the op as the plan runs it, its
convolution alone (raw output + batch statistics) and its BatchNorm apply pass alone, each as N back-to-back launches
between two events.  D3FK_VERBOSE=1 prints the launch geometry of every kernel once.
    python tools/split_convbn.py [--batch 256 --size 64 --repeat 20]"""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from denoising_diffusion_deep_fake_b200 import _lib
from denoising_diffusion_deep_fake_b200.train import DenoiserModule

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--repeat", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
mod = DenoiserModule(encoder_name="resnet34", learning_rate=0.02, noise_exponential_sampling_lambda=5,
                     cosine_scheduler_max_epoch=100, precision="bf16").to(dev).train()
mod.configure_optimizers(fused=True)
x = torch.randn(a.batch, 3, a.size, a.size, device=dev).clamp(-1, 1)
for _ in range(3):
    mod.training_step(x)
torch.cuda.synchronize()
plan = next(p for plans in mod.model._plans.values() for p in plans if p.training)
s = torch.cuda.current_stream().cuda_stream


def timed(op, reps):
    copies = []
    for _ in range(reps):
        c = _lib.Op()
        ctypes.memmove(ctypes.byref(c), ctypes.byref(op), ctypes.sizeof(_lib.Op))
        c.lane = 0
        copies.append(c)
    ol = _lib.OpList(copies)
    ol.run(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ol.run(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print(f"{'layer':44s} {'conv+bn':>9s} {'conv':>9s} {'bn_apply':>9s}   us")
tot = [0.0, 0.0, 0.0]
for op in plan.fwd_ops:
    if op.kind != _lib.OP_CONV_BN:
        continue
    p = _lib.op_params(op)
    cv, bn = p.conv, p.bn
    M, K = cv.B * cv.Ho * cv.Wo, cv.kh * cv.kw * (cv.c0 + cv.c1)
    conv_only = _lib.Op()
    conv_only.kind = _lib.OP_CONV
    ctypes.memmove(ctypes.byref(conv_only.u.conv), ctypes.byref(cv), ctypes.sizeof(_lib.ConvParams))
    bn_only = _lib.Op()
    bn_only.kind = _lib.OP_BN_APPLY
    ctypes.memmove(ctypes.byref(bn_only.u.bn), ctypes.byref(bn), ctypes.sizeof(_lib.BnParams))
    t = (timed(op, a.repeat), timed(conv_only, a.repeat), timed(bn_only, a.repeat))
    for i in range(3):
        tot[i] += t[i]
    print(f"M={M:8d} N={cv.Cout:4d} K={K:5d} k{cv.kh}s{cv.stride} up{cv.up0} lane{op.lane}        {t[0]:9.1f} {t[1]:9.1f} {t[2]:9.1f}", flush=True)
print(f"{'total':44s} {tot[0]:9.1f} {tot[1]:9.1f} {tot[2]:9.1f}")
print("device error flag:", _lib.load().d3fk_device_error_flag())
