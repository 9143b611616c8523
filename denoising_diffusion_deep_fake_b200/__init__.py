"""denoising_diffusion_deep_fake_b200 ("d3fk"): B200-native (sm_100a) hot path of the d3f denoiser —
the resnet34 U-Net forward/backward, q_sample, posterior update and the iterative sampler — behind the
reference's own nn.Module / CLI surface.  Python + torch are plumbing; the arithmetic is libd3fk.so."""
from . import _lib  # noqa: F401
from ._lib import D3fkError  # noqa: F401
from .unet import Unet  # noqa: F401
from .functional import q_sample, posterior_step_, posterior_coeffs, noise_ratio_grid, adam_step_  # noqa: F401
from .loss import MseStructuralSimilarityLoss, ssim  # noqa: F401

__all__ = ["Unet", "q_sample", "posterior_step_", "posterior_coeffs", "noise_ratio_grid", "adam_step_",
           "MseStructuralSimilarityLoss", "ssim", "D3fkError"]
