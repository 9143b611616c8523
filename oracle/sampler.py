"""Iterative reverse-diffusion sampler over the x0-predicting U-Net.  The reference has no
iterative loop (face swap is one pass: d3f/train_deep_fake/lit_module.py:187-189, :259-270); the
N-step sampler is defined by BASELINE.json and SURVEY §8a row S.  This file is its plain-PyTorch
definition; the N=1 / r_start=0 degenerate case is exactly ``fake = model(real)``.
TEST INFRASTRUCTURE ONLY."""
import math
import torch


def noise_ratio_grid(n_steps, r_start=1.0, dtype=torch.float64):
    """r_N = r_start > ... > r_0 = 0, linear in r (alpha_bar = 1 - r)."""
    return torch.linspace(r_start, 0.0, n_steps + 1, dtype=dtype)


def posterior_coeffs(r_i, r_prev, eta):
    """x_prev = c_x0 * x0_hat + c_eps * eps_hat + sigma * z, eps_hat = (x_i - sqrt(1-r_i) x0_hat)/sqrt(r_i).
    Returned as coefficients on (x_i, x0_hat, z)."""
    if r_prev <= 0.0:
        return 0.0, 1.0, 0.0                      # last step returns x0_hat
    if r_i >= 1.0:
        var_ratio = 1.0
    else:
        var_ratio = 1.0 - (1.0 - r_i) / (1.0 - r_prev)
    sigma = eta * math.sqrt(r_prev / r_i) * math.sqrt(max(var_ratio, 0.0))
    c_eps = math.sqrt(max(r_prev - sigma * sigma, 0.0))
    k_xi = c_eps / math.sqrt(r_i)
    k_x0 = math.sqrt(1.0 - r_prev) - c_eps * math.sqrt(1.0 - r_i) / math.sqrt(r_i)
    return k_xi, k_x0, sigma


def posterior_step(x_i, x0_hat, r_i, r_prev, z=None, eta=0.0):
    k_xi, k_x0, sigma = posterior_coeffs(float(r_i), float(r_prev), eta)
    out = k_xi * x_i + k_x0 * x0_hat
    if sigma != 0.0:
        out = out + sigma * z
    return out


@torch.no_grad()
def sample_loop(model, x_start, n_steps, r_start=1.0, eta=0.0, noises=None, return_trajectory=False):
    """x_start is x at ratio r_start (pure noise when r_start = 1).  ``noises[i]`` is z for the step
    from grid index i (only used when eta > 0)."""
    grid = noise_ratio_grid(n_steps, r_start).tolist()
    x = x_start
    traj = []
    for i in range(n_steps):
        x0_hat = model(x)
        z = None if noises is None else noises[i]
        x = posterior_step(x, x0_hat, grid[i], grid[i + 1], z, eta)
        if return_trajectory:
            traj.append(x.clone())
    return (x, traj) if return_trajectory else x
