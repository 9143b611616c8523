"""MSE + (1 - SSIM) training criterion of the reference
(d3f/loss_functions/structural_similarity_loss.py:5-26 with piqa.SSIM() defaults).
On CUDA the forward value AND dL/dprediction come from one fused libd3fk kernel;
`ssim()` below is the same computation composed of torch ops (used for logging / non-fp32 inputs)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


def _window(size=11, sigma=1.5):
    g = [math.exp(-((i - (size - 1) / 2) ** 2) / (2 * sigma ** 2)) for i in range(size)]
    s = sum(g)
    return [v / s for v in g]


class _FusedMseSsim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prediction, target, lo, hi):
        _lib.init(prediction.device.index)
        pred = prediction.contiguous()
        tgt = target.contiguous()
        B, C, H, W = pred.shape
        need_grad = prediction.requires_grad
        grad = torch.empty_like(pred) if need_grad else None
        acc = torch.zeros(2, dtype=torch.float64, device=pred.device)
        win = (_lib.f32 * 12)(*(_window() + [0.0]))
        op = _lib.make_op(_lib.OP_LOSS, B=B, C=C, H=H, W=W, pred=pred.data_ptr(), target=tgt.data_ptr(),
                          grad=None if grad is None else grad.data_ptr(), acc=acc.data_ptr(), lo=float(lo), hi=float(hi),
                          grad_scale=1.0, win=win)
        _lib.run_single(op, torch.cuda.current_stream(pred.device).cuda_stream)
        n_tot = pred.numel()
        n_map = B * C * (H - 10) * (W - 10)
        loss = ((acc[0] / n_tot + 1.0 - acc[1] / n_map) * 0.5).to(torch.float32)
        ctx.save_for_backward(grad) if need_grad else None
        ctx.has_grad = need_grad
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if not ctx.has_grad:
            return None, None, None, None
        (grad,) = ctx.saved_tensors
        return grad * grad_out, None, None, None


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
        if (prediction.is_cuda and prediction.dtype == torch.float32 and target.dtype == torch.float32
                and prediction.dim() == 4 and prediction.shape[-1] % 32 == 0 and prediction.shape[-2] % 32 == 0
                and not target.requires_grad):
            return _FusedMseSsim.apply(prediction, target, self.input_min_value, self.input_max_value)
        mse_loss = F.mse_loss(prediction, target)
        p = self.normalise_between_zero_and_one(prediction)
        t = self.normalise_between_zero_and_one(target)
        return (mse_loss + (1.0 - ssim(p, t))) / 2.0
