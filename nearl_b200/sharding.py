"""Multi-GPU execution of the featuriser: one process per GPU, trajectories sharded by rank.

The path shards embarrassingly (every (trajectory, window, focal point, feature) task is
independent, /root/reference/nearl/featurizer.py:429-441), exactly as the reference scales
out with SLURM array tasks over ``np.array_split`` of the trajectory list
(/root/reference/scripts/featurize_misato.py:46-49,67-69).  There is no collective on the
data path; ``torch.distributed`` (gloo or nccl) is only used to agree on the shard and,
optionally, to bring per-rank results to rank 0 in trajectory order."""
import os

import numpy as np

__all__ = ["shard_indices", "rank_world", "run_sharded", "gather_in_order"]


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
