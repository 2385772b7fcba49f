"""Host-side scheduler tables for the fused CFG + scheduler kernel (c2d_cfg_sched_step).

SD-1.5 schedule (diffusers ``scaled_linear`` betas 0.00085..0.012 over 1000 steps, ``steps_offset=1``,
``set_alpha_to_one=False``, epsilon prediction).  Every update is expressed as
    x <- ca * x + cb * eps ;   next UNet input = x * in_scale
so one kernel (and one captured CUDA graph) serves DDIM (eta = 0) and Euler-discrete alike.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List

import numpy as np


def alphas_cumprod() -> np.ndarray:
    betas = np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=np.float64) ** 2
    return np.cumprod(1.0 - betas)


@dataclass
class Plan:
    timesteps: List[float]        # UNet timestep per step
    coef: np.ndarray              # float32 [steps, 3] = (ca, cb, in_scale_for_next_input)
    init_scale: float             # x0 = noise * init_scale
    first_in_scale: float         # first UNet input = x0 * first_in_scale


def leading_timesteps(n: int) -> List[int]:
    ratio = 1000 // n
    return [(n - 1 - i) * ratio + 1 for i in range(n)]


def ddim_plan(n: int) -> Plan:
    ac = alphas_cumprod()
    ratio = 1000 // n
    ts = leading_timesteps(n)
    coef = np.zeros((n, 3), dtype=np.float64)
    for i, t in enumerate(ts):
        tp = t - ratio
        a_t = ac[t]
        a_p = ac[tp] if tp >= 0 else ac[0]
        # x0 = (x - sqrt(1-a_t) eps) / sqrt(a_t);  x' = sqrt(a_p) x0 + sqrt(1-a_p) eps
        coef[i, 0] = math.sqrt(a_p / a_t)
        coef[i, 1] = math.sqrt(1.0 - a_p) - math.sqrt(a_p) * math.sqrt(1.0 - a_t) / math.sqrt(a_t)
        coef[i, 2] = 1.0
    return Plan([float(t) for t in ts], coef.astype(np.float32), 1.0, 1.0)


def euler_plan(n: int) -> Plan:
    ac = alphas_cumprod()
    sig_all = np.sqrt((1.0 - ac) / ac)
    ts = np.asarray(leading_timesteps(n), dtype=np.float64)
    sig = np.concatenate([np.interp(ts, np.arange(1000), sig_all), [0.0]])
    coef = np.zeros((n, 3), dtype=np.float64)
    for i in range(n):
        coef[i, 0] = 1.0
        coef[i, 1] = sig[i + 1] - sig[i]
        coef[i, 2] = 1.0 / math.sqrt(sig[i + 1] ** 2 + 1.0)
    return Plan([float(t) for t in ts], coef.astype(np.float32), float(sig[0]), 1.0 / math.sqrt(sig[0] ** 2 + 1.0))


def make_plan(name: str, steps: int) -> Plan:
    if name == "ddim":
        return ddim_plan(steps)
    if name == "euler":
        return euler_plan(steps)
    raise ValueError(f"unknown scheduler {name!r} (ddim | euler)")
