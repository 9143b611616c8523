"""q_sample / posterior_step parity (-m gpu) against the oracle formulas
(d3f/train_denoiser/lit_module.py:128-153).  fp32, tolerance 1e-6."""
import pytest
import torch

import oracle
import denoising_diffusion_deep_fake_b200 as d3
from gpu_harness import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("lam", [3.0, 5.0, 8.0])
@pytest.mark.parametrize("shape", [(8, 3, 64, 64), (1, 3, 32, 32), (5, 3, 128, 96)])
def test_q_sample_given_noise(lam, shape):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(shape, generator=g).clamp(-1, 1)
    noise = torch.randn(shape, generator=g)
    y = torch.rand(shape[0], 1, 1, 1, generator=g)
    r_ref = oracle.sample_noise_ratio(y, lam)
    ref = oracle.blend_noise(x, noise, r_ref)
    out, n_used, r = d3.q_sample(x.to(DEV), lam, noise=noise.to(DEV), y=y.to(DEV), return_aux=True)
    assert rel_err(out.cpu(), ref) < 1e-6
    assert rel_err(r.cpu(), r_ref) < 1e-6
    assert (r_ref > 0).all() and (r_ref <= 1).all()


def test_q_sample_philox_statistics():
    x = torch.zeros(64, 3, 64, 64, device=DEV)
    out, noise, r = d3.q_sample(x, 5.0, seed=123, offset=7, return_aux=True)
    # with x = 0: out = sqrt(r) * eps
    assert abs(noise.mean().item()) < 5e-3 and abs(noise.std().item() - 1.0) < 5e-3
    assert rel_err(out.cpu(), (torch.sqrt(r) * noise).cpu()) < 1e-6
    assert (r > 0).all() and (r <= 1).all()
    out2 = d3.q_sample(x, 5.0, seed=123, offset=7)
    assert torch.equal(out, out2)                     # deterministic given (seed, offset)
    out3 = d3.q_sample(x, 5.0, seed=124, offset=7)
    assert not torch.equal(out, out3)
    # fixed ratio variant (balance_training_images/lit_module.py:109-120)
    o4, n4, r4 = d3.q_sample(x + 1.0, 5.0, seed=1, fixed_r=0.7, return_aux=True)
    assert rel_err(o4.cpu(), ((0.3 ** 0.5) * 1.0 + (0.7 ** 0.5) * n4).cpu()) < 1e-6


@pytest.mark.parametrize("eta", [0.0, 1.0])
def test_posterior_step(eta):
    g = torch.Generator().manual_seed(1)
    shape = (4, 3, 64, 64)
    grid = oracle.noise_ratio_grid(10).tolist()
    for i in (0, 4, 9):
        x = torch.randn(shape, generator=g)
        x0 = torch.randn(shape, generator=g)
        z = torch.randn(shape, generator=g)
        ref = oracle.posterior_step(x, x0, grid[i], grid[i + 1], z, eta)
        xd = x.to(DEV).clone()
        d3.posterior_step_(xd, x0.to(DEV), grid[i], grid[i + 1], z=z.to(DEV), eta=eta)
        assert rel_err(xd.cpu(), ref) < 1e-6, i


@pytest.mark.parametrize("shape", [(8, 3, 64, 64), (2, 3, 32, 96), (3, 3, 128, 128)])
def test_fused_mse_ssim_loss(shape):
    """Fused loss kernel (value + dL/dprediction) vs the oracle's MseStructuralSimilarityLoss under autograd
    (d3f/loss_functions/structural_similarity_loss.py:14-26 + piqa SSIM).  fp32: value 1e-5, gradient 1e-4."""
    g = torch.Generator().manual_seed(3)
    target = torch.nn.functional.avg_pool2d(torch.randn(shape, generator=g), 5, 1, 2).clamp(-1, 1) * 1.5
    pred = (target + 0.4 * torch.randn(shape, generator=g)).requires_grad_(True)     # some values outside [-1, 1]
    ref = oracle.MseStructuralSimilarityLoss(-1.0, 1.0)(pred, target)
    ref.backward()
    crit = d3.MseStructuralSimilarityLoss(-1.0, 1.0)
    pd = pred.detach().to(DEV).requires_grad_(True)
    loss = crit(pd, target.to(DEV))
    (loss * 3.0).backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    assert rel_err(pd.grad.cpu() / 3.0, pred.grad) < 1e-4
    # torch-composed fallback path of the same module agrees too (no-grad / non-fused inputs)
    with torch.no_grad():
        l2 = d3.loss.ssim(pd.clamp(-1, 1) * 0.5 + 0.5, target.to(DEV).clamp(-1, 1) * 0.5 + 0.5)
        assert abs(l2.item() - oracle.ssim(pred.detach().clamp(-1, 1) * 0.5 + 0.5, target.clamp(-1, 1) * 0.5 + 0.5).item()) < 1e-5


@pytest.mark.parametrize("n,betas,with_ema", [(1000003, (0.9, 0.999), False), (4096, (0.5, 0.999), True), (37, (0.9, 0.999), True)])
def test_fused_adam_matches_torch_optim(n, betas, with_ema):
    """d3fk_adam vs torch.optim.Adam (d3f/train_denoiser/lit_module.py:95, train_deep_fake/lit_module.py:116-120) and the
    ema_pytorch lerp, three steps, odd sizes exercise the scalar tail."""
    from denoising_diffusion_deep_fake_b200.functional import adam_step_
    g = torch.Generator().manual_seed(11)
    npad = (n + 3) // 4 * 4 + 4
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=0.02, betas=betas)
    flat = torch.zeros(npad, device=DEV)
    p = flat[:n]
    p.copy_(p0)
    m, v = torch.zeros(npad, device=DEV)[:n], torch.zeros(npad, device=DEV)[:n]
    ema = torch.zeros(npad, device=DEV)[:n] if with_ema else None
    ema_ref = torch.zeros(n)
    for step in range(1, 4):
        grad = torch.randn(n, generator=g)
        ref.grad = grad.clone()
        opt.step()
        adam_step_(p, grad.to(DEV), m, v, 0.02, betas[0], betas[1], 1e-8, step, ema=ema, ema_decay=0.9)
        ema_ref.lerp_(ref.detach(), 1 - 0.9)
    assert rel_err(p.cpu(), ref.detach()) < 1e-6
    if with_ema:
        assert rel_err(ema.cpu(), ema_ref) < 1e-6
