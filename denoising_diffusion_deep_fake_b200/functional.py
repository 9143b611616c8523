"""Fused elementwise entry points of the hot path: noising (q_sample) and the posterior update.

q_sample replaces `LitModule.blend_random_amount_of_noise_with_each_sample` +
`sample_random_number_from_exponential_distribution` (d3f/train_denoiser/lit_module.py:128-153,
duplicated at d3f/train_deep_fake/lit_module.py:208-233) — ~10 eager launches -> 1 kernel.
posterior_step is the sampler update of SURVEY §8a row S (no reference counterpart)."""
import math

import torch

from . import _lib
from ._lib import make_op


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda_f32(t, name):
    if not t.is_cuda:
        raise _lib.D3fkError(f"{name}: d3fk kernels need a CUDA (sm_100a) tensor; there is no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32")
    _lib.init(t.device.index)


def q_sample(batch, lam, noise=None, y=None, seed=0, offset=0, fixed_r=None, return_aux=False, out=None):
    """noisy = sqrt(1-r)*batch + sqrt(r)*noise with per-sample r = 1/lam*log(1/(y(1-c)+c)), c = e^-lam.
    noise / y: optional tensors (parity with the reference's torch.randn_like / torch.rand draws);
    otherwise drawn in-kernel from Philox4x32-10(seed, offset)."""
    _require_cuda_f32(batch, "batch")
    batch = batch.contiguous()
    B = batch.shape[0]
    if out is None:
        out = torch.empty_like(batch)
    elif out.shape != batch.shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != batch.device:
        raise ValueError("q_sample: `out` must be a contiguous float32 tensor of the batch's shape on its device")
    r_out = torch.empty(B, dtype=torch.float32, device=batch.device) if return_aux else None
    n_out = torch.empty_like(batch) if (return_aux and noise is None) else None
    if noise is not None:
        noise = noise.contiguous()
    if y is not None:
        y = y.reshape(B).contiguous()
    op = make_op(_lib.OP_QSAMPLE, B=B, chw=batch[0].numel(), lam=float(lam),
                 fixed_r=-1.0 if fixed_r is None else float(fixed_r), x=batch.data_ptr(),
                 noise=None if noise is None else noise.data_ptr(), y=None if y is None else y.data_ptr(),
                 out=out.data_ptr(), r_out=None if r_out is None else r_out.data_ptr(),
                 noise_out=None if n_out is None else n_out.data_ptr(), seed=int(seed), offset=int(offset))
    _lib.run_single(op, _stream(batch))
    if return_aux:
        return out, (noise if noise is not None else n_out), r_out.view(B, 1, 1, 1)
    return out


def noise_ratio_grid(n_steps, r_start=1.0):
    """r_N = r_start > ... > r_0 = 0, linear in r (alpha_bar = 1 - r)."""
    return [r_start * (1.0 - i / n_steps) for i in range(n_steps + 1)]


def posterior_coeffs(r_i, r_prev, eta):
    """Coefficients (k_xi, k_x0, sigma) of x_prev = k_xi*x_i + k_x0*x0_hat + sigma*z for
    x_prev = sqrt(1-r_prev) x0_hat + sqrt(r_prev - sigma^2) eps_hat + sigma z,
    eps_hat = (x_i - sqrt(1-r_i) x0_hat)/sqrt(r_i), sigma = eta*sqrt(r_prev/r_i)*sqrt(1-(1-r_i)/(1-r_prev))."""
    if r_prev <= 0.0:
        return 0.0, 1.0, 0.0
    var_ratio = 1.0 if r_i >= 1.0 else 1.0 - (1.0 - r_i) / (1.0 - r_prev)
    sigma = eta * math.sqrt(r_prev / r_i) * math.sqrt(max(var_ratio, 0.0))
    c_eps = math.sqrt(max(r_prev - sigma * sigma, 0.0))
    k_xi = c_eps / math.sqrt(r_i)
    k_x0 = math.sqrt(1.0 - r_prev) - c_eps * math.sqrt(1.0 - r_i) / math.sqrt(r_i)
    return k_xi, k_x0, sigma


def posterior_step_(x_i, x0_hat, r_i, r_prev, z=None, eta=0.0, seed=0, offset=0):
    """In-place posterior update of x_i (one fused kernel)."""
    _require_cuda_f32(x_i, "x_i")
    if not x_i.is_contiguous() or not x0_hat.is_contiguous():
        raise ValueError("posterior_step_ needs contiguous tensors")
    k_xi, k_x0, sigma = posterior_coeffs(float(r_i), float(r_prev), eta)
    op = make_op(_lib.OP_POSTERIOR, n=x_i.numel(), x=x_i.data_ptr(), x0_hat=x0_hat.data_ptr(),
                 z=None if z is None else z.contiguous().data_ptr(), k_xi=k_xi, k_x0=k_x0, sigma=sigma,
                 seed=int(seed), offset=int(offset))
    _lib.run_single(op, _stream(x_i))
    return x_i


def adam_scalars(lr, beta1, beta2, step, ema_decay=0.0):
    """The per-step scalars of d3fk_adam: (lr, bias1, bias2, ema_decay)."""
    return float(lr), 1.0 - beta1 ** step, 1.0 - beta2 ** step, float(ema_decay)


def adam_step_(p, g, m, v, lr, beta1, beta2, eps, step, ema=None, ema_decay=0.0, grad_scale=1.0, dyn=None):
    """Fused torch.optim.Adam update (+optional EMA lerp) over flat fp32 arenas.  dyn: device float tensor holding
    adam_scalars(...) — the launch then reads its per-step scalars from there (CUDA-graph replay)."""
    _require_cuda_f32(p, "p")
    op = make_op(_lib.OP_ADAM, n=p.numel(), p=p.data_ptr(), g=g.data_ptr(), m=m.data_ptr(), v=v.data_ptr(),
                 ema=None if ema is None else ema.data_ptr(), lr=float(lr), beta1=float(beta1), beta2=float(beta2),
                 eps=float(eps), bias1=1.0 - beta1 ** step, bias2=1.0 - beta2 ** step, ema_decay=float(ema_decay),
                 grad_scale=float(grad_scale), dyn=None if dyn is None else dyn.data_ptr())
    _lib.run_single(op, _stream(p))


def _frames_op(kind, frames, tensor, mean, std):
    N, H, W, _ = frames.shape
    op = make_op(kind, N=N, H=H, W=W, frames=frames.data_ptr(), tensor=tensor.data_ptr(),
                 mean=[float(v) for v in mean], std=[float(v) for v in std])
    _lib.run_single(op, _stream(tensor))


def frames_to_tensor(frames_bgr, mean, std):
    """LitModule.cv2_to_tensor_normalised (d3f/train_deep_fake/lit_module.py:272-283), batched and on the device:
    uint8 BGR frames [N,H,W,3] (or one [H,W,3]) -> normalised fp32 RGB [N,3,H,W] = (v - mean*255) / (std*255)."""
    if frames_bgr.dim() == 3:
        frames_bgr = frames_bgr.unsqueeze(0)
    if not frames_bgr.is_cuda:
        raise _lib.D3fkError("frames must be on a B200 (sm_100a) CUDA device; there is no CPU path")
    if frames_bgr.dtype != torch.uint8 or frames_bgr.dim() != 4 or frames_bgr.shape[-1] != 3:
        raise RuntimeError(f"expected uint8 BGR frames [N,H,W,3], got {frames_bgr.dtype} {tuple(frames_bgr.shape)}")
    frames_bgr = frames_bgr.contiguous()
    N, H, W, _ = frames_bgr.shape
    _lib.init(frames_bgr.device.index if frames_bgr.device.index is not None else torch.cuda.current_device())
    out = torch.empty((N, 3, H, W), dtype=torch.float32, device=frames_bgr.device)
    _frames_op(_lib.OP_FRAMES_TO_TENSOR, frames_bgr, out, mean, std)
    return out


def tensor_to_frames(tensor, mean, std):
    """LitModule.tensor_cv2_to_denormalised (d3f/train_deep_fake/lit_module.py:285-300), batched and on the device:
    fp32 RGB [N,3,H,W] -> uint8 BGR frames [N,H,W,3] = clamp(int(t*std*255 + mean*255), 0, 255).  The input is left
    untouched (the reference scales it in place)."""
    _require_cuda_f32(tensor, "tensor")
    if tensor.dim() != 4 or tensor.shape[1] != 3:
        raise RuntimeError(f"expected fp32 RGB [N,3,H,W], got {tuple(tensor.shape)}")
    tensor = tensor.contiguous()
    N, _, H, W = tensor.shape
    out = torch.empty((N, H, W, 3), dtype=torch.uint8, device=tensor.device)
    _frames_op(_lib.OP_TENSOR_TO_FRAMES, out, tensor, mean, std)
    return out


def random_affine_inverse_maps(B, H, W, degrees=15.0, translate=(0.2, 0.2), scale=(0.8, 1.2), generator=None, device=None,
                               p=1.0):
    """Per-sample inverse affine maps [B,6] (output pixel -> source pixel) with the parameter distribution of the reference's
    `K.RandomAffine(degrees=15, translate=[0.2, 0.2], scale=[0.8, 1.2], shear=0, p=1.0)`
    (d3f/train_denoiser/lit_module.py:55-65): angle ~ U(-degrees, degrees), translation ~ U(-t*W, t*W) x U(-t*H, t*H), one
    isotropic scale ~ U(lo, hi); rotation + scale about the image centre ((W-1)/2, (H-1)/2), then the translation.  Host
    arithmetic on B x 4 numbers (float64, closed-form inverse); the warp itself is d3fk_affine_q_sample.
    p < 1: each sample is augmented with probability p and gets the identity map otherwise (albumentations'
    ShiftScaleRotate(p=0.7) of d3f/train_deep_fake/lit_module.py:99-111; one extra uniform draw per sample)."""
    u = torch.rand(B, 4, generator=generator, dtype=torch.float64)
    keep = torch.rand(B, generator=generator, dtype=torch.float64) < p if p < 1.0 else None
    th = (2 * u[:, 0] - 1) * degrees * math.pi / 180.0
    tx = (2 * u[:, 1] - 1) * translate[0] * W
    ty = (2 * u[:, 2] - 1) * translate[1] * H
    sc = scale[0] + u[:, 3] * (scale[1] - scale[0])
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    # forward: p_out = s R (p_src - c) + c + t, R = [[cos, sin], [-sin, cos]]  =>  p_src = R^T (p_out - c - t) / s + c
    ca, sa = torch.cos(th) / sc, torch.sin(th) / sc
    m = torch.stack([ca, -sa, cx - ca * (cx + tx) + sa * (cy + ty),
                     sa, ca, cy - sa * (cx + tx) - ca * (cy + ty)], dim=1)
    if keep is not None:
        ident = torch.tensor([1.0, 0.0, 0.0, 0.0, 1.0, 0.0], dtype=torch.float64).expand(B, 6)
        m = torch.where(keep.view(B, 1), m, ident)
    return m.to(torch.float32).to(device) if device is not None else m.to(torch.float32)


def affine_q_sample(batch, inverse_maps, lam, noise=None, y=None, seed=0, offset=0, fixed_r=None, return_aux=False,
                    out=None):
    """The reference's `image = augment(image); image_noisy = blend_noise(image)` (d3f/train_denoiser/lit_module.py:113-115) in
    ONE kernel: per-sample affine warp (bilinear, zero padding; `inverse_maps` [B,6] or [B,2,3], output pixel -> source pixel)
    and the noising of the warped image.  Returns (image_augmented, image_noisy) — the loss target and the network input —
    and with return_aux also the per-sample noise ratio.  Noise is that of `q_sample(image_augmented, ...)` with the same
    seed / offset, bit for bit."""
    _require_cuda_f32(batch, "batch")
    if batch.dim() != 4 or batch.shape[-1] % 4:
        raise RuntimeError(f"expected [B,C,H,W] with W a multiple of 4, got {tuple(batch.shape)}")
    batch = batch.contiguous()
    B, C, H, W = batch.shape
    m = inverse_maps.reshape(B, 6).to(device=batch.device, dtype=torch.float32).contiguous()
    aug, noisy = out if out is not None else (torch.empty_like(batch), torch.empty_like(batch))
    for t in (aug, noisy):
        if t.shape != batch.shape or t.dtype != torch.float32 or not t.is_contiguous() or t.device != batch.device:
            raise ValueError("affine_q_sample: `out` must be two contiguous float32 tensors of the batch's shape on its device")
    r_out = torch.empty(B, dtype=torch.float32, device=batch.device) if return_aux else None
    if noise is not None:
        noise = noise.contiguous()
    if y is not None:
        y = y.reshape(B).contiguous()
    op = make_op(_lib.OP_AFFINE_QSAMPLE, B=B, C=C, H=H, W=W, lam=float(lam), fixed_r=-1.0 if fixed_r is None else float(fixed_r),
                 x=batch.data_ptr(), minv=m.data_ptr(), noise=None if noise is None else noise.data_ptr(),
                 y=None if y is None else y.data_ptr(), out_aug=aug.data_ptr(), out_noisy=noisy.data_ptr(),
                 r_out=None if r_out is None else r_out.data_ptr(), seed=int(seed), offset=int(offset))
    _lib.run_single(op, _stream(batch))
    if return_aux:
        return aug, noisy, r_out.view(B, 1, 1, 1)
    return aug, noisy
