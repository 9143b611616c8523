"""Noising ("q_sample") restated from d3f/train_denoiser/lit_module.py:128-153
(duplicate at d3f/train_deep_fake/lit_module.py:208-233).  TEST INFRASTRUCTURE ONLY."""
import math
import torch


def sample_noise_ratio(y, lam):
    """lit_module.py:141-153 given the uniform draw ``y`` of shape [B,1,1,1]:
    c = 1/e^lam ; x = 1/lam * log(1 / (y*(1-c) + c))."""
    c = 1 / math.exp(lam)
    return 1 / lam * torch.log(1 / (y * (1 - c) + c))


def blend_noise(batch, noise, r):
    """lit_module.py:137: sqrt(1-r)*batch + sqrt(r)*noise, r broadcast as [B,1,1,1]."""
    return torch.sqrt(1 - r) * batch + torch.sqrt(r) * noise


def blend_random_amount_of_noise_with_each_sample(batch, lam, generator=None):
    """Same RNG call order as the reference: randn_like (line 131) then rand (line 143)."""
    noise = torch.randn(batch.shape, dtype=batch.dtype, device=batch.device, generator=generator)
    y = torch.rand((batch.shape[0], 1, 1, 1), device=batch.device, generator=generator)
    r = sample_noise_ratio(y, lam)
    return blend_noise(batch, noise, r), noise, r
