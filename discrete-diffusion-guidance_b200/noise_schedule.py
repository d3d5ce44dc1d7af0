"""Noise schedules read by the sampler (reference noise_schedule.py).

Only `total_noise` is on the sampling path (`diffusion.py:1311-1316`); it is a
host-side scalar table computed once per `sample()` with the same fp32 torch
ops the reference uses, then uploaded for the kernels.
"""
from __future__ import annotations

import torch


class Noise:
  def total_noise(self, t):
    raise NotImplementedError

  def __call__(self, t):
    return self.total_noise(t), None


class LogLinearNoise(Noise):
  """sigma(t) = -log1p(-(1 - eps) t)  => move chance (1 - eps) t
  (noise_schedule.py:134-153)."""

  def __init__(self, eps=1e-3):
    self.eps = eps

  def total_noise(self, t):
    return -torch.log1p(-(1 - self.eps) * t)


class CosineNoise(Noise):
  # noise_schedule.py:65-77
  def __init__(self, eps=1e-3):
    self.eps = eps

  def total_noise(self, t):
    return -torch.log(self.eps + (1 - self.eps) * torch.cos(t * torch.pi / 2))


class CosineSqrNoise(Noise):
  # noise_schedule.py:80-94
  def __init__(self, eps=1e-3):
    self.eps = eps

  def total_noise(self, t):
    return -torch.log(
      self.eps + (1 - self.eps) * torch.cos(t * torch.pi / 2) ** 2)


class LinearNoise(Noise):
  # noise_schedule.py:97-112
  def __init__(self, sigma_min=0, sigma_max=10):
    self.sigma_min = torch.tensor(float(sigma_min))
    self.sigma_max = torch.tensor(float(sigma_max))

  def total_noise(self, t):
    return self.sigma_min + t * (self.sigma_max - self.sigma_min)


class GeometricNoise(Noise):
  # noise_schedule.py:122-131
  def __init__(self, sigma_min=1e-3, sigma_max=1):
    self.sigmas = 1.0 * torch.tensor([sigma_min, sigma_max])

  def total_noise(self, t):
    return self.sigmas[0] ** (1 - t) * self.sigmas[1] ** t


def get_noise(config):
  """noise_schedule.get_noise (noise_schedule.py:13-30)."""
  kind = config.noise.type
  if kind == 'geometric':
    return GeometricNoise(config.noise.sigma_min, config.noise.sigma_max)
  if kind == 'loglinear':
    return LogLinearNoise()
  if kind == 'cosine':
    return CosineNoise()
  if kind == 'cosinesqr':
    return CosineSqrNoise()
  if kind == 'linear':
    return LinearNoise(config.noise.sigma_min, config.noise.sigma_max)
  if kind == 'ar':
    return None
  raise NotImplementedError(f'{kind} noise schedule is not implemented.')
