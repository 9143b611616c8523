"""Diagnostics (GPU box): print parity numbers instead of asserting."""
import sys, os, statistics, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle
import denoising_diffusion_deep_fake_b200 as d3
from denoising_diffusion_deep_fake_b200.sampler import Sampler

def rel(a, b): return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
def psnr(a, b):
    mse = ((a - b) ** 2).mean().item(); peak = (b.max() - b.min()).item()
    return 10 * torch.log10(torch.tensor(peak ** 2 / max(mse, 1e-30))).item()

torch.manual_seed(0)
ref = oracle.Unet()
sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
for B in (8, 32):
    x = torch.randn(B, 3, 64, 64).clamp(-1.5, 1.5); dy = torch.randn(B, 3, 64, 64)
    ref.load_state_dict(sd0); ref.train(); ref.zero_grad()
    y_ref = ref(x); y_ref.backward(dy)
    gref = {n: p.grad.clone() for n, p in ref.named_parameters()}
    ref.load_state_dict(sd0)
    with torch.autocast('cpu', dtype=torch.bfloat16):
        y_ac = ref(x)
    for prec in ('fp32', 'bf16'):
        m = d3.Unet(precision=prec); m.load_state_dict(sd0); m.cuda().train()
        y = m(x.cuda()); y.backward(dy.cuda()); torch.cuda.synchronize()
        errs = {n: rel(p.grad.cpu(), gref[n]) for n, p in m.named_parameters()}
        worst = max(errs, key=errs.get)
        print(f'B={B} {prec}: fwd {rel(y.detach().cpu(), y_ref.detach()):.3e} (cpu autocast bf16 oracle {rel(y_ac.float().detach(), y_ref.detach()):.3e}) '
              f'grad median {statistics.median(errs.values()):.3e} max {errs[worst]:.3e} ({worst}) head {errs["segmentation_head.0.weight"]:.3e}', flush=True)
# eval + sampler
ref.load_state_dict(sd0); ref.train()
with torch.no_grad():
    for _ in range(3): ref(torch.randn(8, 3, 64, 64))
sd1 = {k: v.clone() for k, v in ref.state_dict().items()}
ref.eval()
for prec in ('fp32', 'bf16'):
    m = d3.Unet(precision=prec); m.load_state_dict(sd1); m.cuda().eval()
    for B in (1, 8):
        x = torch.randn(B, 3, 64, 64)
        with torch.no_grad():
            print(f'eval {prec} B={B}: {rel(m(x.cuda()).cpu(), ref(x)):.3e}', flush=True)
    B, n = 4, 20
    g = torch.Generator().manual_seed(5)
    x_start = torch.randn(B, 3, 64, 64, generator=g)
    out_ref, traj = oracle.sample_loop(ref, x_start, n, eta=0.0, return_trajectory=True)
    # free running, eager, with per-step comparison
    smp = Sampler(m, B, 64, 64, n, eta=0.0, use_graph=False)
    smp.refresh_weights()
    xs = x_start.cuda().clone()
    stream = torch.cuda.current_stream().cuda_stream
    grid = smp.grid
    line = []
    for i in range(n):
        smp.plan.run_forward(xs, smp.x0_hat, stream)
        d3.posterior_step_(xs, smp.x0_hat, grid[i], grid[i+1], eta=0.0)
        line.append(f'{psnr(xs.cpu(), traj[i]):.1f}')
    print(f'sampler {prec} free-running psnr/step:', ' '.join(line), flush=True)
    # teacher forced
    line = []
    prev = x_start
    for i in range(n):
        xs = prev.cuda().clone()
        smp.plan.run_forward(xs, smp.x0_hat, stream)
        d3.posterior_step_(xs, smp.x0_hat, grid[i], grid[i+1], eta=0.0)
        line.append(f'{psnr(xs.cpu(), traj[i]):.1f}')
        prev = traj[i]
    print(f'sampler {prec} teacher-forced psnr/step:', ' '.join(line), flush=True)
    out = Sampler(m, B, 64, 64, n, eta=0.0, use_graph=True).run(x_start.cuda()).cpu()
    print(f'sampler {prec} graph final psnr {psnr(out, out_ref):.1f}', flush=True)
