"""world_size-2 gloo test of the multi-GPU host logic (no GPU needed): shards cover the
trajectory list exactly once, in the reference's array_split order, and rank 0 reassembles
per-trajectory results in order."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_items, q):
  sys.path.insert(0, ROOT)
  from nearl_b200 import sharding
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  mine = sharding.shard_indices(n_items, rank, world)
  local = {i: np.full(3, i, dtype=np.float32) for i in mine}
  merged = sharding.gather_in_order(local, n_items)
  q.put((rank, mine, None if merged is None else [int(m[0]) for m in merged]))
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
  from nearl_b200 import sharding
  assert sharding.shard_indices(7, 0, 2) == [0, 1, 2, 3] and sharding.shard_indices(7, 1, 2) == [4, 5, 6]
  assert sharding.shard_indices(3, 3, 8) == [] and sum(len(sharding.shard_indices(10, r, 8)) for r in range(8)) == 10
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29500 + (os.getpid() % 2000)
  procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
  for p in procs:
    p.start()
  got = [q.get(timeout=120) for _ in procs]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  by_rank = {r: (mine, merged) for r, mine, merged in got}
  assert by_rank[0][0] == [0, 1, 2, 3] and by_rank[1][0] == [4, 5, 6]
  assert by_rank[0][1] == list(range(7)) and by_rank[1][1] is None
