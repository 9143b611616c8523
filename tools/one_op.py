"""Run ONE op of the training plan a few times (for ncu captures): python tools/one_op.py wgrad 4096 256 2304"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from denoising_diffusion_deep_fake_b200 import _lib
from denoising_diffusion_deep_fake_b200.train import DenoiserModule
kind_name, M, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dev = torch.device("cuda:0")
mod = DenoiserModule(encoder_name="resnet34", learning_rate=0.02, noise_exponential_sampling_lambda=5,
                     cosine_scheduler_max_epoch=100, precision="bf16").to(dev).train()
mod.configure_optimizers(fused=True)
x = torch.randn(256, 3, 64, 64, device=dev).clamp(-1, 1)
mod.training_step(x)
torch.cuda.synchronize()
plan = next(p for plans in mod.model._plans.values() for p in plans if p.training)
s = torch.cuda.current_stream().cuda_stream
kind = _lib.OP_WGRAD if kind_name == "wgrad" else _lib.OP_CONV
ops = list(plan.fwd_ops) + [op for seg in plan.bwd_segments for op in seg]
if kind_name in ("bn_bwd", "bn_apply"):          # python tools/one_op.py bn_bwd <count> <C> 0 [launch index to capture]
    want = (_lib.OP_BN_BWD, _lib.OP_BN_BWD_REDUCE) if kind_name == "bn_bwd" else (_lib.OP_BN_APPLY,)
    for op in ops:
        p = _lib.op_params(op)
        if op.kind in want and (p.count, p.C) == (M, N):
            c = _lib.Op(); ctypes.memmove(ctypes.byref(c), ctypes.byref(op), ctypes.sizeof(_lib.Op))
            ol = _lib.OpList([c])
            for _ in range(3):
                ol.run(s)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            ol.run(s)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            print("ran", kind_name, M, N)
            break
    else:
        print("op not found")
    sys.exit(0)
for op in ops:
    if op.kind != kind and not (kind == _lib.OP_CONV and op.kind == _lib.OP_CONV_BN):
        continue
    p = _lib.op_params(op)
    if op.kind == _lib.OP_CONV_BN:
        p = p.conv
    if (p.B * p.Ho * p.Wo, p.Cout, p.kh * p.kw * (p.c0 + p.c1)) != (M, N, K):
        continue
    if kind == _lib.OP_CONV and p.mode != mode:
        continue
    c = _lib.Op(); ctypes.memmove(ctypes.byref(c), ctypes.byref(op), ctypes.sizeof(_lib.Op))
    ol = _lib.OpList([c] * 1)
    for _ in range(3):
        ol.run(s)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()       # ncu --profile-from-start off: capture from here
    for _ in range(2):
        ol.run(s)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("ran", kind_name, M, N, K)
    break
else:
    print("op not found")
