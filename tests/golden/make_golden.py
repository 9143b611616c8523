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


def main_rows_f3_f4():
    """Second fixture: frame conversion bytes and the affine warp.  Kept in its own file so that adding it
    did not rewrite the first one."""
    rng = np.random.default_rng(7)
    frames = rng.integers(0, 256, size=(2, 8, 12, 3), dtype=np.uint8)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    tens = oracle.cv2_to_tensor_normalised(frames, mean, std)
    g = torch.Generator().manual_seed(11)
    net_out = tens + 0.3 * torch.randn(tens.shape, generator=g)          # stands in for a network output
    back = oracle.tensor_cv2_to_denormalised(net_out, mean, std)
    x = torch.randn(3, 3, 16, 24, generator=g).clamp(-1, 1)
    angle, tx, ty, sc = oracle.sample_affine_params(3, 16, 24, generator=g)
    _, minv = oracle.affine_matrices(angle, tx, ty, sc, 16, 24)
    warped = oracle.warp_affine_bilinear(x, minv[:, :2])
    out = dict(frames=frames, mean=np.array(mean, np.float32), std=np.array(std, np.float32), tensor=tens.numpy(),
               net_out=net_out.numpy(), frames_back=back, aff_x=x.numpy(), aff_angle=angle.numpy(), aff_tx=tx.numpy(),
               aff_ty=ty.numpy(), aff_scale=sc.numpy(), aff_minv=minv[:, :2].numpy(), aff_warped=warped.numpy())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "d3f_golden_f3_f4.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: tuple(np.shape(v)) for k, v in out.items()})


if __name__ == "__main__":
    if "--rows-f3-f4" in sys.argv:
        main_rows_f3_f4()
    else:
        main()
        main_rows_f3_f4()
