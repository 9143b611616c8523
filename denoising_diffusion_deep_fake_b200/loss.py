"""MSE + (1 - SSIM) training criterion of the reference
(d3f/loss_functions/structural_similarity_loss.py:5-26 with piqa.SSIM() defaults).
First cut stays composed of torch ops on the GPU (SURVEY §8f row f1 schedules the fused kernel)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _gaussian_1d(size, sigma, dtype, device):
    x = torch.arange(size, dtype=dtype, device=device) - (size - 1) / 2
    g = torch.exp(-x ** 2 / (2 * sigma ** 2))
    return g / g.sum()


def ssim(x, y, window_size=11, sigma=1.5, value_range=1.0, k1=0.01, k2=0.03):
    c = x.shape[1]
    g = _gaussian_1d(window_size, sigma, x.dtype, x.device)
    kh = g.view(1, 1, -1, 1).expand(c, 1, -1, 1).contiguous()
    kw = g.view(1, 1, 1, -1).expand(c, 1, 1, -1).contiguous()
    n = x.shape[0]
    stack = torch.cat([x, y, x * x, y * y, x * y], dim=0)
    f = F.conv2d(F.conv2d(stack, kh, groups=c), kw, groups=c)
    mu_x, mu_y, e_xx, e_yy, e_xy = f[:n], f[n:2 * n], f[2 * n:3 * n], f[3 * n:4 * n], f[4 * n:]
    c1, c2 = (k1 * value_range) ** 2, (k2 * value_range) ** 2
    mu_xx, mu_yy, mu_xy = mu_x * mu_x, mu_y * mu_y, mu_x * mu_y
    cs = (2 * (e_xy - mu_xy) + c2) / ((e_xx - mu_xx) + (e_yy - mu_yy) + c2)
    ss = (2 * mu_xy + c1) / (mu_xx + mu_yy + c1) * cs
    return ss.mean(dim=(1, 2, 3)).mean()


class MseStructuralSimilarityLoss(nn.Module):
    def __init__(self, input_min_value, input_max_value):
        super().__init__()
        self.input_min_value = input_min_value
        self.input_max_value = input_max_value

    def normalise_between_zero_and_one(self, x):
        x = (x - self.input_min_value) / (self.input_max_value - self.input_min_value)
        return x.clip(0.0, 1.0)

    def forward(self, prediction, target):
        mse_loss = F.mse_loss(prediction, target)
        p = self.normalise_between_zero_and_one(prediction)
        t = self.normalise_between_zero_and_one(target)
        return (mse_loss + (1.0 - ssim(p, t))) / 2.0
