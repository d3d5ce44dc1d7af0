"""world_size-2 gloo test of the multi-GPU host logic (batch sharding + the one
all_gather of token ids)."""
import os
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeSampler:
  """Stands in for Diffusion on CPU: ids derived from the per-rank seed."""

  def __init__(self, B, L):
    self.config = {'seed': 3}
    self.seed = None
    self.B, self.L = B, L

  def sample(self, eps=1e-5):
    g = torch.Generator().manual_seed(self.seed)
    return torch.randint(0, 30522, (self.B, self.L), generator=g)


def _worker(rank, world, port, q):
  sys.path.insert(0, ROOT)
  os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                    MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  import ddg_b200
  from ddg_b200 import dist as ddist
  r, w, _ = ddist.init_from_env(backend='gloo')
  assert (r, w) == (rank, world)
  out = ddist.sample_sharded(_FakeSampler(4, 16))
  q.put((rank, out))
  torch.distributed.barrier()
  torch.distributed.destroy_process_group()


def test_sample_sharded_world2():
  sys.path.insert(0, ROOT)
  import ddg_b200  # noqa: F401
  from ddg_b200 import dist as ddist
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = 29600 + os.getpid() % 300
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  results = dict(q.get(timeout=120) for _ in range(2))
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert torch.equal(results[0], results[1])        # every rank holds the gather
  assert tuple(results[0].shape) == (8, 16)
  for r in range(2):                                 # rank order, rank-distinct keys
    g = torch.Generator().manual_seed(ddist.rank_seed(3, r))
    assert torch.equal(results[0][4 * r:4 * r + 4],
                       torch.randint(0, 30522, (4, 16), generator=g))
  assert not torch.equal(results[0][:4], results[0][4:])
