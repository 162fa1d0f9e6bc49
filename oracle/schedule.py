"""Noise schedule tables (oracle; test infrastructure only).

Restates ``Betas`` (networks/dm3d.py:194-214 == networks/conditional_dm3d.py:215-235):
float64 numpy tables, each cast to float32 constants.
"""
from __future__ import annotations

import numpy as np


class Betas:
    def __init__(self, timesteps: int):
        beta = np.linspace(0.0001, 0.02, timesteps)
        alpha = 1 - beta
        sqrt_alpha = np.sqrt(alpha)
        alpha_bar = np.cumprod(alpha, 0)
        alpha_bar_prev = np.append(1.0, alpha_bar[:-1])
        sqrt_alpha_bar = np.sqrt(alpha_bar)
        sqrt_alpha_bar_prev = np.sqrt(alpha_bar_prev)
        sqrt_one_minus_alpha_bar = np.sqrt(1 - alpha_bar)
        f32 = lambda a: np.asarray(a, dtype=np.float32)  # noqa: E731
        self.timesteps = timesteps
        self.beta = f32(beta)
        self.alpha = f32(alpha)
        self.sqrt_alpha = f32(sqrt_alpha)
        self.alpha_bar = f32(alpha_bar)
        self.alpha_bar_prev = f32(alpha_bar_prev)
        self.sqrt_alpha_bar = f32(sqrt_alpha_bar)
        self.sqrt_alpha_bar_prev = f32(sqrt_alpha_bar_prev)
        self.sqrt_one_minus_alpha_bar = f32(sqrt_one_minus_alpha_bar)
