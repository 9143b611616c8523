"""MseStructuralSimilarityLoss restated from d3f/loss_functions/structural_similarity_loss.py:5-26
plus piqa.SSIM() defaults (piqa is not installed).  TEST INFRASTRUCTURE ONLY."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def gaussian_kernel_1d(size=11, sigma=1.5, dtype=torch.float32, device=None):
    x = torch.arange(size, dtype=dtype, device=device) - (size - 1) / 2
    g = torch.exp(-x ** 2 / (2 * sigma ** 2))
    return g / g.sum()


def ssim(x, y, window_size=11, sigma=1.5, value_range=1.0, k1=0.01, k2=0.03):
    """piqa.SSIM defaults: separable depthwise valid Gaussian filter, per-image mean of ss over (C,H',W'),
    then mean over the batch."""
    c = x.shape[1]
    g = gaussian_kernel_1d(window_size, sigma, x.dtype, x.device)
    kh = g.view(1, 1, -1, 1).repeat(c, 1, 1, 1)
    kw = g.view(1, 1, 1, -1).repeat(c, 1, 1, 1)

    def filt(t):
        return F.conv2d(F.conv2d(t, kh, groups=c), kw, groups=c)

    c1 = (k1 * value_range) ** 2
    c2 = (k2 * value_range) ** 2
    mu_x, mu_y = filt(x), filt(y)
    mu_xx, mu_yy, mu_xy = mu_x ** 2, mu_y ** 2, mu_x * mu_y
    s_xx = filt(x ** 2) - mu_xx
    s_yy = filt(y ** 2) - mu_yy
    s_xy = filt(x * y) - mu_xy
    cs = (2 * s_xy + c2) / (s_xx + s_yy + c2)
    ss = (2 * mu_xy + c1) / (mu_xx + mu_yy + c1) * cs
    return ss.mean(dim=(1, 2, 3)).mean()


class MseStructuralSimilarityLoss(nn.Module):
    def __init__(self, input_min_value, input_max_value):
        super().__init__()
        self.input_min_value = input_min_value
        self.input_max_value = input_max_value

    def forward(self, prediction, target):
        mse_loss = F.mse_loss(prediction, target)                      # :15
        prediction = self.normalise_between_zero_and_one(prediction)    # :17
        target = self.normalise_between_zero_and_one(target)            # :18
        ssim_loss = 1.0 - ssim(prediction, target)                      # :19
        return (mse_loss + ssim_loss) / 2.0                             # :21

    def normalise_between_zero_and_one(self, x):
        x = (x - self.input_min_value) / (self.input_max_value - self.input_min_value)
        return x.clip(0.0, 1.0)
