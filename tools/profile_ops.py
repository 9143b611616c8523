"""Per-op device-time breakdown of one training step / one sampling step (CUDA events around every op,
through d3fk_run_profile).  Run on the GPU box:  python tools/profile_ops.py [--batch 256 --size 64]"""
import argparse, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import denoising_diffusion_deep_fake_b200 as d3
from denoising_diffusion_deep_fake_b200 import _lib
from denoising_diffusion_deep_fake_b200.train import DenoiserModule

KIND = {v: k for k, v in vars(_lib).items() if k.startswith("OP_")}

def describe(op):
    p = _lib.op_params(op)
    if op.kind == _lib.OP_CONV_BN:
        p = p.conv
        M = p.B * p.Ho * p.Wo; K = p.kh * p.kw * (p.c0 + p.c1)
        return f"+BN mode{p.mode} M={M} N={p.Cout} K={K} k{p.kh}s{p.stride} up{p.up0}", 2.0 * M * p.Cout * K
    if op.kind == _lib.OP_CONV:
        M = p.B * p.Ho * p.Wo; K = p.kh * p.kw * (p.c0 + p.c1)
        return f"mode{p.mode} M={M} N={p.Cout} K={K} k{p.kh}s{p.stride} up{p.up0}", 2.0 * M * p.Cout * K
    if op.kind == _lib.OP_WGRAD:
        M = p.B * p.Ho * p.Wo; K = p.kh * p.kw * (p.c0 + p.c1)
        return f"M={M} N={p.Cout} K={K} k{p.kh}s{p.stride} up{p.up0}", 2.0 * M * p.Cout * K
    if op.kind == _lib.OP_WGRAD_GROUP:
        b = p.base
        M = b.B * b.Ho * b.Wo; K = b.kh * b.kw * b.c0
        return f"x{p.count} M={M} N={b.Cout} K={K} k{b.kh}s{b.stride}", 2.0 * M * b.Cout * K * p.count
    if op.kind in (_lib.OP_BN_APPLY, _lib.OP_BN_BWD_REDUCE, _lib.OP_BN_BWD_APPLY):
        return f"count={p.count} C={p.C}", 0.0
    return "", 0.0

def repeat_ms(oplist, stream, reps):
    """Steady-state time of every op: `reps` back-to-back launches of the same op between two events
    (warm L2, launch latency pipelined) — the per-op event profile has a ~8 us floor per op."""
    import ctypes
    out = []
    for op in oplist:
        copies = []
        for _ in range(reps):
            c = _lib.Op()
            ctypes.memmove(ctypes.byref(c), ctypes.byref(op), ctypes.sizeof(_lib.Op))
            copies.append(c)
        ol = _lib.OpList(copies)
        ol.run(stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ol.run(stream)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / reps)
    return out


def report(name, oplist, stream, top=25):
    if REPEAT:
        ms = repeat_ms(oplist, stream, REPEAT)
    else:
        oplist.profile(stream)
        ms = oplist.profile(stream)
    rows = []
    by_kind = collections.defaultdict(float)
    for op, t in zip(oplist, ms):
        d, fl = describe(op)
        rows.append((t, KIND[op.kind], d, fl))
        by_kind[KIND[op.kind]] += t
    total = sum(ms)
    print(f"==== {name}: {len(ms)} ops, {total:.3f} ms")
    for k, t in sorted(by_kind.items(), key=lambda kv: -kv[1]):
        print(f"   {k:22s} {t:8.3f} ms  {100*t/total:5.1f}%")
    rows.sort(key=lambda r: -r[0])
    for t, k, d, fl in rows[:top]:
        tf = f"{fl / t / 1e9:8.1f} TF/s" if fl else ""
        print(f"   {t:8.4f} ms {k:18s} {d} {tf}")
    return total

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--top", type=int, default=25)
ap.add_argument("--repeat", type=int, default=0, help="time each op as N back-to-back launches (steady state)")
a = ap.parse_args()
REPEAT = a.repeat
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
report("pack", plan.pack_ops, s, top=5)
report("train forward", plan.fwd_ops, s, a.top)
for i, seg in enumerate(plan.bwd_segments):
    report(f"backward segment {i}", seg, s, a.top if (i == 0 or a.top > 100) else 12)
# whole-step pieces outside the plan
import time
def timed(fn, n=5):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
pred = mod.model(x).detach().requires_grad_(True)
def loss_fb():
    l = mod.training_criterion(pred, x); l.backward()
print(f"loss fwd+bwd (torch ops): {timed(loss_fb):.3f} ms")
print(f"q_sample: {timed(lambda: d3.q_sample(x, 5.0, seed=1)):.3f} ms")
print(f"fused adam: {timed(lambda: mod.optimizer.step()):.3f} ms")
print(f"full training_step: {timed(lambda: mod.training_step(x)):.3f} ms")
mod.model.eval()
with torch.no_grad():
    xe = torch.randn(64, 3, 128, 128, device=dev)
    mod.model(xe)
plan_e = next(p for plans in mod.model._plans.values() for p in plans if not p.training)
report("eval forward B=64 128x128", plan_e.fwd_ops, s, a.top)
