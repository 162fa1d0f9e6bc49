"""Reverse-diffusion sampler (oracle; test infrastructure only).

Restates ``DiffusionModel.sample`` (networks/dm3d.py:477-508 == conditional_dm3d.py:517-548)
and ``DiffusionModel.generate`` (dm3d.py:510-532; conditional_dm3d.py:550-575): float32
tables, float32 arithmetic in the written order, the posterior MEAN is clipped to [-1,1],
sigma = exp(0.5*ln(max(var,1e-20))), noise = 0 at i == 0.  Noise is INJECTED (the
reference's TF generator is not reproducible) or drawn from oracle.philox.
"""
from __future__ import annotations

import numpy as np
import torch

from .schedule import Betas
from . import philox


def sample(b: Betas, x_t: torch.Tensor, pred_noise: torch.Tensor, t: int):
    """-> (posterior_mean, variance).  Same t for the whole batch, as in generate()."""
    f = lambda a: torch.tensor(a[t], dtype=torch.float32)  # noqa: E731
    beta, sqa, ab, ab_prev = f(b.beta), f(b.sqrt_alpha), f(b.alpha_bar), f(b.alpha_bar_prev)
    sqab, sqab_prev, sq1ab = f(b.sqrt_alpha_bar), f(b.sqrt_alpha_bar_prev), f(b.sqrt_one_minus_alpha_bar)
    x_0 = (x_t - sq1ab * pred_noise) / sqab
    mean = (beta * sqab_prev / (1 - ab)) * x_0 + ((1 - ab_prev) * sqa / (1 - ab)) * x_t
    var = (1 - ab_prev) * beta / (1 - ab)
    return mean, var


def ddpm_step(b: Betas, x_t, pred_noise, t: int, noise):
    mean, var = sample(b, x_t, pred_noise, t)
    mean = torch.clamp(mean, -1.0, 1.0)
    sigma = torch.exp(0.5 * torch.log(torch.clamp(var, min=1e-20)))
    if t == 0 or noise is None:
        return mean
    return mean + sigma * noise


def ddim_step(b: Betas, x_t, pred_noise, t: int, t_prev: int):
    """EXTENSION (not in the reference): deterministic DDIM (eta=0) step from t
    to t_prev (< t, or -1 for the final step) with the reference's clip moved to x0."""
    f = lambda a, i: torch.tensor(a[i], dtype=torch.float32)  # noqa: E731
    sqab, sq1ab = f(b.sqrt_alpha_bar, t), f(b.sqrt_one_minus_alpha_bar, t)
    x_0 = torch.clamp((x_t - sq1ab * pred_noise) / sqab, -1.0, 1.0)
    if t_prev < 0:
        return x_0
    return f(b.sqrt_alpha_bar, t_prev) * x_0 + f(b.sqrt_one_minus_alpha_bar, t_prev) * pred_noise


def generate(network, b: Betas, x_T: torch.Tensor, last_step: int = 0, noises=None,
             seed=None, sample_ids=None, record=None):
    """network(x, t_int) -> eps_hat.  ``noises[i]`` is the noise used at step i (i>0), or
    Philox noise when ``seed`` is given."""
    x = x_T.clone()
    B = x.shape[0]
    n_elem = x[0].numel()
    if sample_ids is None:
        sample_ids = np.arange(B)
    for i in range(b.timesteps - 1, last_step - 1, -1):
        if i > 0:
            if noises is not None:
                noise = noises[i]
            else:
                noise = torch.from_numpy(philox.normal(seed, i, sample_ids, n_elem)).reshape(x.shape)
        else:
            noise = None
        eps = network(x, i)
        if record is not None:
            record(i, x, eps)
        x = ddpm_step(b, x, eps, i, noise)
    return x
