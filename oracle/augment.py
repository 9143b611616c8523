"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the GPU augmentation of the denoiser's training step,
`K.RandomAffine(degrees=15, translate=[0.2, 0.2], scale=[0.8, 1.2], shear=0, p=1.0)`
(d3f/train_denoiser/lit_module.py:55-65, applied at :113).  kornia is an un-vendored, unpinned dependency that is not
installed here: PARITY UNPINNED against kornia itself.  What is restated, from kornia's published
behaviour: per-sample parameters angle ~ U(-degrees, degrees), translation ~ U(-t*W, t*W) x U(-t*H, t*H), one isotropic
scale ~ U(lo, hi); the map is rotation+scale about the image centre ((W-1)/2, (H-1)/2) in the OpenCV convention
(`get_rotation_matrix2d`) followed by the translation; the image is resampled bilinearly with zero padding at the same
size.  The oracle is pinned to torch's own `grid_sample(bilinear, zeros, align_corners=True)` on the same inverse maps
(tests/test_oracle.py)."""
import math

import torch


def sample_affine_params(B, H, W, degrees=15.0, translate=(0.2, 0.2), scale=(0.8, 1.2), generator=None):
    """[B] tensors: angle (degrees), tx, ty (pixels), scale."""
    u = torch.rand(B, 4, generator=generator, dtype=torch.float64)
    angle = (2 * u[:, 0] - 1) * degrees
    tx = (2 * u[:, 1] - 1) * translate[0] * W
    ty = (2 * u[:, 2] - 1) * translate[1] * H
    sc = scale[0] + u[:, 3] * (scale[1] - scale[0])
    return angle, tx, ty, sc


def affine_matrices(angle, tx, ty, sc, H, W):
    """Forward maps (source pixel -> output pixel) [B,3,3] float64 and their inverses."""
    B = angle.shape[0]
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    th = angle.double() * math.pi / 180.0
    alpha, beta = sc.double() * torch.cos(th), sc.double() * torch.sin(th)
    M = torch.zeros(B, 3, 3, dtype=torch.float64)
    M[:, 0, 0], M[:, 0, 1], M[:, 0, 2] = alpha, beta, (1 - alpha) * cx - beta * cy + tx.double()
    M[:, 1, 0], M[:, 1, 1], M[:, 1, 2] = -beta, alpha, beta * cx + (1 - alpha) * cy + ty.double()
    M[:, 2, 2] = 1.0
    return M, torch.linalg.inv(M)


def warp_affine_bilinear(x, minv):
    """x [B,C,H,W]; minv [B,3,3] or [B,2,3] (output pixel -> source pixel).  Bilinear, zero padding, same size."""
    B, C, H, W = x.shape
    m = minv.to(torch.float32)
    oy, ox = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    out = torch.zeros_like(x)
    for b in range(B):
        sx = m[b, 0, 0] * ox + (m[b, 0, 1] * oy + m[b, 0, 2])
        sy = m[b, 1, 0] * ox + (m[b, 1, 1] * oy + m[b, 1, 2])
        x0, y0 = torch.floor(sx), torch.floor(sy)
        wx1, wy1 = sx - x0, sy - y0
        wx0, wy0 = 1 - wx1, 1 - wy1

        def tap(yy, xx):
            ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
            v = x[b][:, yy.clamp(0, H - 1).long(), xx.clamp(0, W - 1).long()]
            return torch.where(ok, v, torch.zeros_like(v))

        out[b] = wy0 * (wx0 * tap(y0, x0) + wx1 * tap(y0, x0 + 1)) + wy1 * (wx0 * tap(y0 + 1, x0) + wx1 * tap(y0 + 1, x0 + 1))
    return out
