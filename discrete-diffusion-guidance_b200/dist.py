"""Batch-sharded sampling across the GPUs of one box.

The reference samples on a single device (`diffusion.py:948-950`,
`configs/config.yaml:39-41`).  Sequences are independent, so here each rank
samples its own `sampling.batch_size` slice with a rank-distinct Philox key and
the token ids are gathered once at the end (one `all_gather` of (B_local, L)
int64 over NCCL/NVLink; gloo on CPU for the tests).  No collective touches the
per-step path.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
  """Initialises torch.distributed from RANK/WORLD_SIZE/MASTER_* if launched
  under torchrun; returns (rank, world_size, local_rank)."""
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if world > 1 and not dist.is_initialized():
    if backend is None:
      backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29500')
    dist.init_process_group(backend=backend, rank=rank, world_size=world)
  return rank, world, local_rank


def rank_seed(seed: int, rank: int) -> int:
  """Distinct Philox key per rank (same formula on every rank)."""
  return (int(seed) * 1_000_003 + 7_777 * int(rank)) & 0x7FFFFFFFFFFFFFFF


def shard_bounds(total: int, world: int, rank: int):
  """Contiguous slice [lo, hi) of `total` sequences owned by `rank`."""
  base, rem = divmod(total, world)
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def gather_samples(local_ids: torch.Tensor) -> torch.Tensor:
  """all_gather of (B_local, L) ids -> (world * B_local, L) in rank order."""
  if not dist.is_initialized() or dist.get_world_size() == 1:
    return local_ids
  parts = [torch.empty_like(local_ids) for _ in range(dist.get_world_size())]
  dist.all_gather(parts, local_ids.contiguous())
  return torch.cat(parts, dim=0)


def sample_sharded(model, eps=1e-5):
  """Every rank calls `model.sample()` on its shard; returns the gathered ids
  on every rank."""
  rank = dist.get_rank() if dist.is_initialized() else 0
  model.seed = rank_seed(model.config.get('seed', 1), rank)
  return gather_samples(model.sample(eps=eps))
