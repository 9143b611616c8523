"""Phase timeline of a tiny tensor-core conv launched back to back (needs the -DD3FK_TIMELINE build:
D3FK_LIB=tools/libd3fk_tl.so python tools/timeline.py M N K [mode])."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from denoising_diffusion_deep_fake_b200 import _lib
from denoising_diffusion_deep_fake_b200.train import DenoiserModule
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda:0")
mod = DenoiserModule(encoder_name="resnet34", learning_rate=0.02, noise_exponential_sampling_lambda=5,
                     cosine_scheduler_max_epoch=100, precision="bf16").to(dev).train()
mod.configure_optimizers(fused=True)
x = torch.randn(256, 3, 64, 64, device=dev).clamp(-1, 1)
mod.training_step(x)
torch.cuda.synchronize()
plan = next(p for plans in mod.model._plans.values() for p in plans if p.training)
s = torch.cuda.current_stream().cuda_stream
ops = list(plan.fwd_ops) + [op for seg in plan.bwd_segments for op in seg]
lib = _lib.load()
for op in ops:
    if op.kind not in (_lib.OP_CONV, _lib.OP_CONV_BN):
        continue
    p = _lib.op_params(op)
    if op.kind == _lib.OP_CONV_BN:
        p = p.conv
    if (p.B * p.Ho * p.Wo, p.Cout, p.kh * p.kw * (p.c0 + p.c1), p.mode) != (M, N, K, mode):
        continue
    c = _lib.Op(); ctypes.memmove(ctypes.byref(c), ctypes.byref(op), ctypes.sizeof(_lib.Op))
    if c.kind == _lib.OP_CONV_BN:
        c.kind = _lib.OP_CONV
    if os.environ.get('NOSTATS') == '1':
        c.u.conv.stats = None      # timing experiment: epilogue without the batch-statistics reduction
    reps = 12
    ol = _lib.OpList([c] * reps)
    ol.run(s); torch.cuda.synchronize()
    n0 = ctypes.c_uint(); buf = (ctypes.c_ulonglong * (512 * 16))()
    lib.d3fk_debug_timeline(buf, 512 * 16, ctypes.byref(n0))
    ol.run(s); torch.cuda.synchronize()
    n1 = ctypes.c_uint()
    lib.d3fk_debug_timeline(buf, 512 * 16, ctypes.byref(n1))
    names = ["entry", "prologue", "pdl_wait", "first_data", "mma_issued", "acc_full", "epi_done", "exit", "chunk0", "chunks", "tiles_done", "csync1", "scattered", "csync2", "reduced", "flushed"]
    prev_exit = None
    print("launch: " + " ".join(f"{n:>10s}" for n in names) + "   (ns since this launch's entry; gap = entry - previous exit)")
    for i in range(n0.value, n1.value):
        row = [buf[(i % 512) * 16 + k] for k in range(16)]
        gap = (row[0] - prev_exit) if prev_exit else 0
        print(f"{i:6d}: " + " ".join(f"{(v - row[0]) if v else -1:10d}" for v in row) + f"   gap {gap}")
        prev_exit = row[7]
    break
else:
    print("op not found")
