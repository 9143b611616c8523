"""MSE + (1 - SSIM) training criterion of the reference
(d3f/loss_functions/structural_similarity_loss.py:5-26 with piqa.SSIM() defaults).
The forward value AND dL/dprediction come from ONE fused libd3fk kernel launch; there is no
CPU or torch fallback for the criterion.  `ssim()` is a CUDA logging helper composed of torch ops."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


def _window(size=11, sigma=1.5):
    g = [math.exp(-((i - (size - 1) / 2) ** 2) / (2 * sigma ** 2)) for i in range(size)]
    s = sum(g)
    return [v / s for v in g]


_WORKSPACE = {}   # device -> 3 zeroed doubles (sum of squared errors, sum of the SSIM map, block ticket); self-resetting


def _workspace(device):
    ws = _WORKSPACE.get(device)
    if ws is None:
        ws = torch.zeros(4, dtype=torch.float64, device=device)
        _WORKSPACE[device] = ws
    return ws


def _launch(prediction, target, lo, hi, need_grad):
    """ONE launch: the scalar loss and (optionally) dL/dprediction."""
    _lib.init(prediction.device.index)
    pred = prediction.contiguous()
    tgt = target.contiguous()
    B, C, H, W = pred.shape
    grad = torch.empty_like(pred) if need_grad else None
    loss = torch.empty((), dtype=torch.float32, device=pred.device)
    win = (_lib.f32 * 12)(*(_window() + [0.0]))
    op = _lib.make_op(_lib.OP_LOSS, B=B, C=C, H=H, W=W, pred=pred.data_ptr(), target=tgt.data_ptr(),
                      grad=None if grad is None else grad.data_ptr(), acc=_workspace(pred.device).data_ptr(),
                      lo=float(lo), hi=float(hi), grad_scale=1.0, win=win, loss_out=loss.data_ptr())
    _lib.run_single(op, torch.cuda.current_stream(pred.device).cuda_stream)
    return loss, grad


class _FusedMseSsim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prediction, target, lo, hi):
        need_grad = prediction.requires_grad
        loss, grad = _launch(prediction, target, lo, hi, need_grad)
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
    """piqa.SSIM value composed of torch CUDA ops — a logging / diagnostic helper, never on the training hot path
    (the criterion below is the fused kernel)."""
    if not x.is_cuda:
        raise _lib.D3fkError("d3fk.ssim is a CUDA logging helper; there is no CPU path")
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

    def value_and_grad(self, prediction, target):
        """(loss, dL/dprediction) without an autograd node: a trainer that owns the step calls
        `prediction.backward(grad)` itself and saves the `grad * grad_output` pass autograd would add (train.DenoiserModule)."""
        self._check(prediction, target)
        return _launch(prediction.detach(), target, self.input_min_value, self.input_max_value, True)

    def forward(self, prediction, target):
        self._check(prediction, target)
        return _FusedMseSsim.apply(prediction, target, self.input_min_value, self.input_max_value)

    def _check(self, prediction, target):
        if not (prediction.is_cuda and target.is_cuda):
            raise _lib.D3fkError("d3fk.MseStructuralSimilarityLoss runs only on a B200 (sm_100a) CUDA device; "
                                 "there is no CPU path")
        if (prediction.dtype != torch.float32 or target.dtype != torch.float32 or prediction.dim() != 4
                or prediction.shape != target.shape or prediction.shape[-1] % 32 or prediction.shape[-2] % 32):
            raise ValueError("expected float32 [B,C,H,W] prediction and target of equal shape with H and W multiples of 32 "
                             f"(the U-Net's own constraint), got {tuple(prediction.shape)} {prediction.dtype}")
        if target.requires_grad:
            raise NotImplementedError("the criterion does not back-propagate into the target (the reference never does)")
