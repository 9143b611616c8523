"""Generates tests/golden/d3f_golden.npz from the oracle (run here, in the build container):
    python tests/golden/make_golden.py
The reference itself cannot be imported (smp / piqa / lightning are not installed — SURVEY §8c), so these are
vectors of the ORACLE, whose encoder is torchvision's own ResNet; they pin the oracle against silent drift and
give the GPU suite fixed known answers.  Weights are the oracle's torch.manual_seed(0) construction."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402


def main():
    torch.manual_seed(0)
    model = oracle.Unet()
    g = torch.Generator().manual_seed(1234)
    x = torch.nn.functional.avg_pool2d(0.5 * torch.randn(2, 3, 32, 32, generator=g), 5, 1, 2).mul(2.5).clamp(-1, 1)
    noise = torch.randn(2, 3, 32, 32, generator=g)
    y = torch.rand(2, 1, 1, 1, generator=g)
    lam = 5.0
    r = oracle.sample_noise_ratio(y, lam)
    noisy = oracle.blend_noise(x, noise, r)
    model.train()
    pred = model(noisy)
    loss = oracle.MseStructuralSimilarityLoss(-1.0, 1.0)(pred, x)
    loss.backward()
    head_w_grad = model.segmentation_head[0].weight.grad.clone()
    stem_g_grad = model.encoder.bn1.weight.grad.clone()
    rm = model.encoder.bn1.running_mean.clone()
    model.eval()
    with torch.no_grad():
        pred_eval = model(noisy)
        samp = oracle.sample_loop(model, noise, 4, eta=0.0)
    out = dict(x=x, noise=noise, y=y, r=r, noisy=noisy, pred_train=pred.detach(), loss=loss.detach(),
               head_w_grad=head_w_grad, stem_gamma_grad=stem_g_grad, bn1_running_mean=rm, pred_eval=pred_eval,
               sample4=samp, ssim_xn=oracle.ssim(x.clamp(0, 1), noisy.clamp(0, 1)))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "d3f_golden.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
    print("wrote", path, {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
