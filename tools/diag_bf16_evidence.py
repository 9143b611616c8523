"""Evidence for the bf16 tolerance: on the SAME GPU, against the float64 oracle,
   d3fk bf16  |  torch eager/cuDNN bf16 autocast (channels_last)  |  d3fk fp32  |  torch eager/cuDNN fp32 (TF32 off)
for x0_hat and for the gradient arena of one training step, on random-init weights and on weights after a short run of the
reference training step, at B = 8 and B = 256 (64x64).  Prints a table; run on the GPU box:
    python tools/diag_bf16_evidence.py > gpurun_out/bf16_evidence.txt"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import oracle
import denoising_diffusion_deep_fake_b200 as d3

DEV = "cuda:0"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-300)).item()


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-300)).item()


def faces(B, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = 0.5 * torch.randn(B, 3, 64, 64, generator=g, device=DEV)
    return (torch.nn.functional.avg_pool2d(x, 5, 1, 2) * 2.5).clamp(-1, 1).contiguous()


def torch_run(ref, sd, dtype, amp, noisy, x0):
    m = copy.deepcopy(ref)
    m.load_state_dict(sd)
    m = m.to(DEV).to(dtype).train()
    if amp:
        m = m.to(memory_format=torch.channels_last)
        noisy = noisy.contiguous(memory_format=torch.channels_last)
    crit = oracle.MseStructuralSimilarityLoss(-1.0, 1.0)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        pred = m(noisy.to(dtype))
    loss = crit(pred.to(dtype if not amp else torch.float32), x0.to(dtype if not amp else torch.float32))
    loss.backward()
    names = [n for n, _ in m.named_parameters()]
    return pred.detach(), {n: p.grad.detach() for n, p in m.named_parameters()}, names


def d3fk_run(sd, precision, noisy, x0):
    m = d3.Unet(precision=precision)
    m.load_state_dict(sd)
    m.to(DEV).train()
    crit = d3.MseStructuralSimilarityLoss(-1.0, 1.0)
    pred = m(noisy)
    crit(pred, x0).backward()
    torch.cuda.synchronize()
    return pred.detach(), {n: p.grad.detach() for n, p in m.named_parameters()}


def main():
    torch.manual_seed(0)
    ref = oracle.Unet()
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    sd1 = {k: v.cpu() for k, v in oracle.short_training_run(ref, sd0, steps=150, device=DEV).items()}
    print(f"{'weights':12s} {'B':>4s} {'implementation':34s} {'x0_hat rel':>11s} {'grad arena rel':>15s} {'grad cos':>11s} {'median/tensor':>14s}")
    for wname, sd in (("random-init", sd0), ("trained-150", sd1)):
        for B in (8, 256):
            x0 = faces(B, 7)
            g = torch.Generator(device=DEV).manual_seed(11)
            noise = torch.randn(x0.shape, generator=g, device=DEV)
            y = torch.rand((B, 1, 1, 1), generator=g, device=DEV)
            noisy = d3.q_sample(x0, 5.0, noise=noise, y=y)
            p64, g64, names = torch_run(ref, sd, torch.float64, False, noisy, x0)
            flat64 = torch.cat([g64[n].flatten() for n in names])
            rows = []
            p, gr, _ = torch_run(ref, sd, torch.float32, True, noisy, x0)
            rows.append(("torch cuDNN bf16 autocast ch_last", p, gr))
            p, gr = d3fk_run(sd, "bf16", noisy, x0)
            rows.append(("d3fk bf16 (tcgen05)", p, gr))
            p, gr, _ = torch_run(ref, sd, torch.float32, False, noisy, x0)
            rows.append(("torch cuDNN fp32 (TF32 off)", p, gr))
            p, gr = d3fk_run(sd, "fp32", noisy, x0)
            rows.append(("d3fk fp32 (parity mode)", p, gr))
            for name, p, gr in rows:
                flat = torch.cat([gr[n].flatten().double() for n in names])
                per = sorted(rel(gr[n], g64[n]) for n in names)
                print(f"{wname:12s} {B:4d} {name:34s} {rel(p, p64):11.3e} {rel(flat, flat64):15.3e} {cos(flat, flat64):11.8f} "
                      f"{per[len(per) // 2]:14.3e}", flush=True)
    print("device error flag:", d3._lib.load().d3fk_device_error_flag())


if __name__ == "__main__":
    main()
