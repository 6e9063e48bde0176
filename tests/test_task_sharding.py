"""One trajectory over several GPUs (frame-slice x focal point first, slabs of the voxel grid when
there are fewer of those than GPUs).  CPU: the partition covers every part exactly once, slabs tile the grid, the parts
of a 2-rank gloo world are assembled by rank 0 in the reference's order.  GPU: the assembled result of a sharded run has
the bits of the unsharded run, whatever the world size."""
import os
import sys

import numpy as np
import pytest

import nearl_b200
from nearl_b200 import features, featurizer, io, sharding, utils

nearl_b200.update_config(quiet=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parts_and_slabs_cover_exactly_once():
  for n in (7, 8, 16, 20, 24, 32, 100, 128):
    for n_slabs in (1, 2, 3, 4, 8, 20):
      cuts = [sharding.slab_range(n, s, n_slabs) for s in range(n_slabs)]
      live = [c for c in cuts if c[1] > 0]
      assert live[0][0] == 0 and sum(c for _, c in live) == n
      assert all(b0 + c0 == b1 and b1 % sharding.TILE == 0 for (b0, c0), (b1, _) in zip(live, live[1:]))
      assert all(c == (0, 0) for c in cuts[len(live):])                  # more slabs than tiles: the last ones stay empty
      assert max(c for _, c in live) - min(c for _, c in live[:-1] or live) <= sharding.TILE
  for units, world in ((10, 8), (2, 8), (1, 3), (5, 5), (0, 4), (3, 8), (40, 8), (7, 1)):
    owner = {}
    for rank in range(world):
      for unit, cuts in sharding.task_parts(units, rank, world).items():
        for cut in cuts:
          assert (unit, cut) not in owner
          owner[(unit, cut)] = rank
    n_slabs = max((c[1] for _, c in owner), default=1)
    assert len(owner) == units * n_slabs
    if units:
      loads = [sum(1 for r in owner.values() if r == rank) for rank in range(world)]
      assert max(loads) - min(loads) <= 1 and (units >= world or min(loads) >= 1)
  assert sharding.task_parts(2, 5, 8) == {1: [(1, 4)]} and sharding.task_parts(2, 0, 2, n_slabs=2) == {0: [(0, 2)], 1: [(0, 2)]}


class _Recorder:
  """Stands in for the device: every part a rank 'computes' is a grid filled with a code of (trajectory, task, row)."""

  @staticmethod
  def grid(tid, bid, pid, fidx, xb, xc, dim):
    rows = np.arange(xb, xb + (xc or dim), dtype=np.float32)
    return np.broadcast_to((1000 * tid + 100 * bid + 10 * pid + fidx + rows / 1000)[:, None, None], (len(rows), dim, dim)).copy()


def _worker(rank, world, port, q):
  sys.path.insert(0, ROOT)
  import torch.distributed as dist
  from nearl_b200 import featurizer as fzmod, sharding as sh
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  dim, n_traj, n_win, n_focal, n_feat = 16, 3, 1, 1, 2

  class Fz(fzmod.Featurizer):
    def run(self, trajectory_indices=None, task_shard=None):
      r, w, n_slabs = task_shard
      self.SHARD_RESULTS = {}
      for tid in range(n_traj):
        for unit, cuts in sorted(sh.task_parts(n_win * n_focal, r, w, n_slabs).items()):
          bid, pid = divmod(unit, n_focal)
          for slab, of in cuts:
            xb, xc = sh.slab_range(dim, slab, of)
            for fidx in range(n_feat):
              self.SHARD_RESULTS.setdefault(tid, []).append((bid, pid, fidx, xb, _Recorder.grid(tid, bid, pid, fidx, xb, xc, dim)))
      return {"trajectories": n_traj}

  fz = Fz({"dimensions": dim, "lengths": 8, "time_window": 1})

  class Feat:
    outfile = None
    def __init__(self, key): self.outkey = key
  fz.FEATURESPACE = [Feat("a"), Feat("b")]
  seen = []
  fz.result_sink = lambda outkey, tid, bid, pid, grid: seen.append((outkey, tid, bid, pid, grid.shape, float(grid[0, 0, 0]), float(grid[-1, 0, 0])))
  sh.run_task_sharded(fz, rank=rank, world=world, device=0)
  q.put((rank, seen))
  dist.barrier()
  dist.destroy_process_group()


def test_two_ranks_assemble_slabs_in_order_over_gloo():
  import torch.multiprocessing as mp
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29500 + (os.getpid() % 2000) + 7
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  got = dict(q.get(timeout=180) for _ in procs)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert got[1] == []                                       # only rank 0 hands results on
  want = []
  for tid in range(3):
    for fidx, key in enumerate("ab"):
      want.append((key, tid, 0, 0, (16, 16, 16), float(np.float32(1000 * tid + fidx)), float(np.float32(1000 * tid + fidx + np.float32(15) / 1000))))
  assert got[0] == want                                     # one unit, two ranks: two 8-row slabs put together, reference order


def _make(trajs, dim, window, outfile=None):
  fz = featurizer.Featurizer({"dimensions": dim, "lengths": dim * 0.5, "time_window": window, "cutoff": 3.0, "sigma": 1.2,
                              **({"outfile": outfile} if outfile else {})})
  fz.register_features([
    features.Mass(outkey="mass"),
    features.MarchingObservers(obs="existence", agg="mean", outkey="exist"),
    features.MarchingObservers(obs="distinct_count", agg="max", weight_type="atomic_number", outkey="distinct"),
    features.MarchingObservers(obs="mean_distance", agg="standard_deviation", weight_type="mass", outkey="meandist"),
    features.MarchingObservers(obs="radius_of_gyration", agg="mean", weight_type="mass", outkey="rog"),      # scan kernel
    features.DensityFlow(agg="mean", weight_type="mass", outkey="flow"),
    features.HeavyAtom(outkey="heavy24", dimensions=24, lengths=12.0),                                      # another grid size
  ])
  fz.register_focus([":ARG", ":LYS"], "mask")
  fz.register_trajloader(trajs)
  return fz


@pytest.mark.gpu
@pytest.mark.parametrize("world,n_slabs", [(2, None), (3, None), (8, None), (16, None), (4, 3)])
def test_sharded_run_has_the_bits_of_the_unsharded_run(tmp_path, world, n_slabs):
  """4 (window, focal point) units per trajectory: 2 and 3 ranks split by unit, 8 ranks cut every 40^3 grid into two slabs,
  16 into four (the 24^3 feature has only three tiles: its fourth slab is empty); n_slabs = 3 forces uneven slabs."""
  trajs = [io.synthetic_trajectory(2500 + 100 * i, 6, seed=900 + i, pin=False) for i in range(2)]
  ref = _make(trajs, 40, 3, outfile=str(tmp_path / "one.h5"))
  ref.run()
  fz = _make(trajs, 40, 3, outfile=str(tmp_path / "many.h5"))
  merged, n_parts = {}, []
  for rank in range(world):
    fz.run(task_shard=(rank, world, n_slabs))
    n_parts.append(sum(len(v) for v in fz.SHARD_RESULTS.values()))
    for tid, items in fz.SHARD_RESULTS.items():
      merged.setdefault(tid, []).extend(items)
  assert sum(1 for n in n_parts if n) == min(world, 4 * (n_slabs or max(1, -(-world // 4))))
  fz.deliver_shard_results(merged)
  with utils.open_hdf(str(tmp_path / "one.h5"), "r") as a, utils.open_hdf(str(tmp_path / "many.h5"), "r") as b:
    keys = sorted(k for k in a.keys() if k != "featurizer_parms")
    assert keys == sorted(k for k in b.keys() if k != "featurizer_parms") and len(keys) == 7
    for key in keys:
      assert a[key].shape == b[key].shape and a[key].shape[0] == 2 * 2 * 2
      assert np.array_equal(a[key][:], b[key][:]), key
    assert a["mass"][:].sum() > 0 and a["rog"][:].sum() > 0


@pytest.mark.gpu
def test_sharded_run_skips_what_the_unsharded_run_skips(tmp_path):
  """A trajectory without focal point is skipped by every rank alike; one window (a single unit) over 5 ranks is cut into 5
  slabs of a 40^3 grid = 5 tiles, rank by rank; results kept in memory (keep_results) and sunk arrive in the same order."""
  trajs = [io.synthetic_trajectory(2000 + 100 * i, 3, seed=950 + i, pin=False, ligand_atoms=(0 if i == 1 else 16)) for i in range(3)]

  def make():
    fz = featurizer.Featurizer({"dimensions": 40, "lengths": 20.0, "time_window": 3, "cutoff": 3.0, "sigma": 1.2})
    fz.register_features([features.Mass(outkey="mass"), features.MarchingObservers(obs="direct_count", agg="mean", outkey="count")])
    fz.register_focus([":LIG"], "mask")
    fz.register_trajloader(trajs)
    fz.keep_results = True
    return fz

  ref = make()
  ref.run()
  fz = make()
  order = []
  fz.result_sink = lambda outkey, tid, bid, pid, grid: order.append((outkey, tid, bid, pid))
  merged = {}
  for rank in range(5):
    stats = fz.run(task_shard=(rank, 5))
    assert stats["trajectories"] == 2
    assert all(arr.shape == (8, 40, 40) for items in fz.SHARD_RESULTS.values() for *_, arr in items)
    for tid, items in fz.SHARD_RESULTS.items():
      merged.setdefault(tid, []).extend(items)
  assert sorted(merged) == [0, 2]
  fz.deliver_shard_results(merged)
  assert order == [(k, t, 0, 0) for t in (0, 2) for k in ("mass", "count")]
  for key in ("mass", "count"):
    assert len(fz.RESULTS[key]) == 2 and all(np.array_equal(a, b) for a, b in zip(fz.RESULTS[key], ref.RESULTS[key])), key
  with pytest.raises(ValueError):
    fz.run(task_shard=(5, 5))
