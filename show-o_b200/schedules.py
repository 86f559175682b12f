"""Host-side mask-ratio schedules for the MaskGIT sampler.

Public names and numerical behaviour follow the reference's schedule helpers (models/sampling.py:39-78:
cosine / linear / powN / sigmoid, selected by `get_mask_chedule`), because `Showo.t2i_generate` accepts the schedule
as a Python callable that the host evaluates once per denoise step on a 0-d fp32 tensor.  The per-token work
(gumbel noise, confidence ranking) lives in csrc/sampler.cu.
"""
from __future__ import annotations

import math

import torch

_HALF_PI = math.pi * 0.5
_FLOOR = 1e-6


def cosine_schedule(t):
    # NB: evaluated as (t * pi) * 0.5 in fp32 like the reference; cos(pi/2) = -4.37e-8 at the last step on purpose
    return torch.cos(t * math.pi * 0.5)


def linear_schedule(t):
    return (1 - t).clamp(min=_FLOOR, max=1.0)


class _PowSchedule:
    def __init__(self, method: str):
        self.exponent = float(method.replace("pow", ""))

    def __call__(self, t):
        return (1.0 - t ** self.exponent).clamp(min=_FLOOR, max=1.0)


class _SigmoidSchedule:
    def __init__(self, start=-3, end=3, tau=1.0, clip_min=_FLOOR):
        self.start, self.end, self.tau, self.clip_min = start, end, tau, clip_min

    def __call__(self, t):
        lo = torch.sigmoid(torch.tensor(self.start / self.tau))
        hi = torch.sigmoid(torch.tensor(self.end / self.tau))
        mid = torch.sigmoid((t * (self.end - self.start) + self.start) / self.tau)
        return torch.clip((hi - mid) / (hi - lo), self.clip_min, 1.0)


def get_mask_chedule(method, **schedule_kwargs):
    """(sic) -- the reference spells it `get_mask_chedule`; kept so `from models import get_mask_chedule` works."""
    table = {"cosine": lambda: cosine_schedule, "linear": lambda: linear_schedule,
             "sigmoid": lambda: _SigmoidSchedule(**schedule_kwargs)}
    if method in table:
        return table[method]()
    if "pow" in method:
        return _PowSchedule(method)
    raise ValueError("Unknown schedule method: {}".format(method))


def step_schedule(noise_schedule, timesteps: int, num_vq_tokens: int, temperature: float):
    """Per-step host scalars of Showo.t2i_generate (modeling_showo.py:157-173): for step s,
    floor(N * noise_schedule((s+1)/T)) and the temperature after compounding by (1 - ratio)."""
    floors, temps = [], []
    for step in range(timesteps):
        ratio = 1.0 * (step + 1) / timesteps
        floors.append(int((num_vq_tokens * noise_schedule(torch.tensor(ratio))).floor().item()))
        temperature = temperature * (1.0 - ratio)
        temps.append(float(temperature))
    return floors, temps
