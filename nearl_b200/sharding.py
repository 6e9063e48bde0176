"""Multi-GPU execution of the featuriser: one process per GPU, trajectories sharded by rank.

The path shards embarrassingly (every (trajectory, window, focal point, feature) task is
independent, /root/reference/nearl/featurizer.py:429-441), exactly as the reference scales
out with SLURM array tasks over ``np.array_split`` of the trajectory list
(/root/reference/scripts/featurize_misato.py:46-49,67-69).  There is no collective on the
data path; ``torch.distributed`` (gloo or nccl) is only used to agree on the shard and,
optionally, to bring per-rank results to rank 0 in trajectory order."""
import os

import numpy as np

__all__ = ["shard_indices", "rank_world", "run_sharded", "gather_in_order", "task_parts", "slab_range", "run_task_sharded"]

TILE = 8     # edge of a kernel tile along the slab axis (NRL_TX): slabs start on tile boundaries


def rank_world():
  return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard_indices(n_items, rank, world):
  """Contiguous split, the reference's ``np.array_split(trajlist, task_nr)[task_index]``."""
  return [int(i) for i in np.array_split(np.arange(n_items), world)[rank]]


def gather_in_order(local, n_items, group=None):
  """`local`: {trajectory index: object}.  Returns on rank 0 the list ordered by trajectory
  index (None elsewhere).  Works over gloo (CPU) or nccl process groups."""
  import torch.distributed as dist
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return [local.get(i) for i in range(n_items)]
  world = dist.get_world_size(group)
  bucket = [None] * world if dist.get_rank(group) == 0 else None
  dist.gather_object(local, bucket, dst=0, group=group)
  if dist.get_rank(group) != 0:
    return None
  merged = {}
  for part in bucket:
    merged.update(part)
  return [merged.get(i) for i in range(n_items)]


def run_sharded(featurizer, rank=None, world=None, device=None):
  """Run `featurizer` on this rank's trajectories.  Output files get the reference's
  per-task suffix (``{prefix}{rank}.h5``-style) so that shards never collide."""
  from . import _lib
  if rank is None or world is None:
    rank, world = rank_world()
  if device is None:
    device = int(os.environ.get("LOCAL_RANK", rank))
  if _lib.device_count() > 0:
    _lib.check(_lib.lib().nrl_set_device(device % _lib.device_count()))
  if world > 1:
    for feat in featurizer.FEATURESPACE:
      if feat.outfile and not getattr(feat, "_sharded_name", False):
        stem, ext = os.path.splitext(feat.outfile)
        feat.outfile = f"{stem}{rank}{ext}"
        feat._sharded_name = True
  mine = shard_indices(featurizer.TRAJECTORYNUMBER, rank, world)
  return featurizer.run(trajectory_indices=mine), mine


# ---- one trajectory over several GPUs -----------------------------------------------------------------------------
# few large trajectories (BASELINE configs[4]: ONE 200k-atom system, 128^3 grid, two windows, eight
# GPUs) do not shard by trajectory.  Their independent units are the (window, focal point) pairs, and -- when there are
# fewer of those than GPUs -- slabs of the voxel grid: every voxel of a task is independent (featurizer.py:429-441,
# marching_observers.cu:405-429; the gather form of the voxeliser, whose per-atom normalisation does not depend on the
# slab).  Still no collective on the data path: every rank computes its parts of every trajectory, rank 0 receives them.

def task_parts(n_units, rank, world, n_slabs=None):
  """{unit: [(slab, n_slabs), ...]} -- the parts of a trajectory's `n_units` (window, focal point) units this rank
  computes.  Units are dealt round-robin; with fewer units than ranks every unit is cut into `n_slabs` grid slabs
  (default: enough for every rank to get a part) and the (unit, slab) parts are dealt instead."""
  if n_slabs is None:
    n_slabs = 1 if n_units >= world or n_units == 0 else -(-world // n_units)
  n_slabs = max(1, int(n_slabs))
  parts = {}
  for q in range(n_units * n_slabs):
    if q % world == rank:
      unit, slab = divmod(q, n_slabs)
      parts.setdefault(unit, []).append((slab, n_slabs))
  return parts


def slab_range(n, slab, n_slabs):
  """(x_begin, x_count) of slab `slab` of `n_slabs` along an axis of `n` voxels, cut on tile boundaries (contiguous,
  disjoint, covering; the first slabs get the extra tiles); (0, 0) when the axis has fewer tiles than slabs and this one
  is left without."""
  tiles = -(-n // TILE)
  base, extra = divmod(tiles, min(n_slabs, max(tiles, 1)))
  if slab >= min(n_slabs, tiles):
    return 0, 0
  t0 = slab * base + min(slab, extra)
  t1 = t0 + base + (1 if slab < extra else 0)
  return t0 * TILE, min(n, t1 * TILE) - t0 * TILE


def run_task_sharded(featurizer, rank=None, world=None, device=None, group=None, n_slabs=None, trajectory_indices=None):
  """Featurise every trajectory on ALL ranks, each computing its `task_parts`; rank 0 receives the parts of a trajectory,
  puts the grids together and hands them on -- `result_sink`, `keep_results`, `Feature.dump` -- in the reference's order,
  so the HDF5 file is the one a single GPU writes.  Returns the run's statistics (rank 0: of the assembled results)."""
  import torch.distributed as dist
  from . import _lib
  if rank is None or world is None:
    rank, world = rank_world()
  if device is None:
    device = int(os.environ.get("LOCAL_RANK", rank))
  if _lib.device_count() > 0:
    _lib.check(_lib.lib().nrl_set_device(device % _lib.device_count()))
  stats = featurizer.run(trajectory_indices=trajectory_indices, task_shard=(rank, world, n_slabs))
  local = featurizer.SHARD_RESULTS
  featurizer.SHARD_RESULTS = {}
  if world > 1 and dist.is_initialized():
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local, bucket, dst=0, group=group)
  else:
    bucket = [local]
  if rank != 0:
    return stats
  merged = {}
  for part in bucket:
    for tid, items in part.items():
      merged.setdefault(tid, []).extend(items)
  featurizer.deliver_shard_results(merged)
  return stats
