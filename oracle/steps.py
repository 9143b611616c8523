"""The reference's training-step FLOWS restated as plain functions (which tensor goes where, in which RNG order).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned to the reference's own code: tests/golden/make_ref_golden.py
lifts the three methods out of /root/reference with `ast`, runs them on a stub module with stand-in networks, and
tests/test_ref_pin.py asserts these functions reproduce their losses bit for bit.

  denoiser_training_step               d3f/train_denoiser/lit_module.py:107-126
  training_denoise_step_for_one_model  d3f/train_deep_fake/lit_module.py:168-181
  training_swap_step_for_one_model     d3f/train_deep_fake/lit_module.py:183-206
"""
import torch
import torch.nn.functional as F

from .noising import blend_random_amount_of_noise_with_each_sample


def denoiser_training_step(model, criterion, image, lam, augmentation=None, generator=None):
    """image -> [augmentation] -> noising -> model -> criterion(prediction, image).  Returns (loss, aux)."""
    if augmentation is not None:
        image = augmentation(image)                                                      # :113
    image_noisy, noise, r = blend_random_amount_of_noise_with_each_sample(image, lam, generator)   # :115
    image_prediction = model(image_noisy)                                                # :117
    loss = criterion(image_prediction, image)                                            # :119
    return loss, dict(image=image, image_noisy=image_noisy, noise=noise, r=r, image_prediction=image_prediction)


def training_denoise_step_for_one_model(real, real_model, criterion, lam, generator=None):
    with torch.no_grad():
        noisy_real, noise, r = blend_random_amount_of_noise_with_each_sample(real, lam, generator)   # :171
    real_prediction = real_model(noisy_real)                                             # :173
    loss = criterion(real_prediction, real)                                              # :175
    return loss, dict(noisy_real=noisy_real, noise=noise, r=r, real_prediction=real_prediction)


def training_swap_step_for_one_model(real, real_model, fake_model, criterion, lam, generator=None):
    """fake_model is the EMA wrapper of the OTHER identity's network (ema_pytorch.EMA; oracle.EMA)."""
    fake_model.update()                                                                  # :185
    with torch.no_grad():
        fake = fake_model(real)                                                          # :189  one pass, no noise
        swap_diff = F.mse_loss(real, fake)                                               # :191  logged only
        noisy_fake, noise, r = blend_random_amount_of_noise_with_each_sample(fake, lam, generator)   # :193
    real_prediction = real_model(noisy_fake)                                             # :195
    loss = criterion(real_prediction, real)                                              # :197
    return loss, dict(fake=fake, swap_diff=swap_diff, noisy_fake=noisy_fake, noise=noise, r=r,
                      real_prediction=real_prediction)


def short_training_run(ref, sd0, steps, batch=16, device="cpu", lam=5.0, lr=0.02, size=64, seed0=100, dtype=torch.float32,
                       calibrate=0):
    """State dict after `steps` reference training steps on synthetic faces (Adam lr 0.02, lambda 5:
    d3f/train_denoiser/denoiser_config.yml:3,8), starting from state `sd0` of oracle network `ref` (not modified).
    The parity tests use it to leave the chaotic random-init regime: BatchNorm over freshly initialised weights amplifies
    1e-6-class forward differences into 1e-3-class gradient differences between ANY two implementations
    (tests/test_ref_pin.py::test_oracle_fp32_gradients_against_fp64).
    dtype float64 makes a GPU run reproducible (atomics-order noise of 1e-16 does not grow to anything visible in 150 steps;
    in float32 two runs of the same script end in visibly different states).  calibrate: that many extra train-mode
    forward passes with frozen weights, so that the BatchNorm running statistics the eval-mode forward folds in belong to
    the final weights (at lr 0.02 they lag several steps behind).  Returned tensors are float32 (integers unchanged)."""
    import copy
    from .loss import MseStructuralSimilarityLoss
    m = copy.deepcopy(ref)
    m.load_state_dict(sd0)
    m = m.to(device).to(dtype).train()
    crit = MseStructuralSimilarityLoss(-1.0, 1.0)
    opt = torch.optim.Adam(m.parameters(), lr=lr)
    gen = torch.Generator(device=device).manual_seed(1)

    def batch_of(i):
        g = torch.Generator(device=device).manual_seed(seed0 + i)
        x = F.avg_pool2d(0.5 * torch.randn(batch, 3, size, size, generator=g, device=device), 5, 1, 2).mul(2.5).clamp(-1, 1)
        return x.to(dtype)

    for i in range(steps):
        loss, _ = denoiser_training_step(m, crit, batch_of(i), lam, generator=gen)
        opt.zero_grad()
        loss.backward()
        opt.step()
    with torch.no_grad():
        for i in range(calibrate):
            noisy, _, _ = blend_random_amount_of_noise_with_each_sample(batch_of(steps + i), lam, gen)
            m(noisy)
    return {k: (v.detach().float() if v.is_floating_point() else v.detach().clone()) for k, v in m.state_dict().items()}
