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


def run_task_sharded(featurizer, rank=None, world=None, device=None, group=None, n_slabs=None, trajectory_indices=None,
                     gather_uploads=True):
  """Featurise every trajectory on ALL ranks, each computing its `task_parts`; rank 0 receives the parts of a trajectory,
  puts the grids together and hands them on -- `result_sink`, `keep_results`, `Feature.dump` -- in the reference's order,
  so the HDF5 file is the one a single GPU writes.  Over an NCCL process group the parts never visit the host on the
  way: every rank leaves them in device memory and sends them to rank 0 point to point (NVLink), rank 0 copies each slab
  to its place in a page-locked buffer of whole grids; and the coordinates go up once, split over the ranks, and are
  completed on every device by an all-gather (`gather_uploads`).  Otherwise (gloo, tests) the parts travel as host arrays.
  Returns the run's statistics."""
  import torch.distributed as dist
  from . import _lib
  if rank is None or world is None:
    rank, world = rank_world()
  if device is None:
    device = int(os.environ.get("LOCAL_RANK", rank))
  if _lib.device_count() > 0:
    _lib.check(_lib.lib().nrl_set_device(device % _lib.device_count()))
  if world > 1 and dist.is_initialized() and dist.get_backend(group) == "nccl":
    return _run_task_sharded_nccl(featurizer, rank, world, group, n_slabs, trajectory_indices, gather_uploads)
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


def _lib_module():
  from . import _lib
  return _lib


def _run_task_sharded_nccl(featurizer, rank, world, group, n_slabs, trajectory_indices, gather_uploads=True):
  import torch
  import torch.distributed as dist
  indices = list(range(featurizer.TRAJECTORYNUMBER) if trajectory_indices is None else trajectory_indices)
  depth = max(1, min(int(featurizer.pipeline_depth), 8))
  state = featurizer.__dict__.setdefault("_shard_nccl_state", {"lane": {}, "recv": None, "whole": None})
  lane_bufs = state["lane"]

  def device_out(lane, n_floats):
    buf = lane_bufs.get(lane)
    if buf is None or buf.numel() < n_floats:
      buf = lane_bufs[lane] = torch.empty(max(int(n_floats * 1.25), 1), dtype=torch.float32, device="cuda")
    return buf.data_ptr()

  # The coordinates: every rank has the trajectory in host memory, and every rank's kernels need (nearly) all of it -- the
  # crop reads every atom of every frame of the rank's windows.  Instead of `world` full uploads through the host's
  # memory system, every rank uploads 1 / world of the frames and an all-gather over NVLink completes the copy on every
  # device (configs[4] on eight GPUs: 60 MB instead of 480 MB per rank over PCIe).  The collective is the same on every
  # rank in the same order (same trajectories, same lanes); the lane's kernels are ordered behind it on the device.
  import ctypes
  xyz_bufs, stage_bufs = {}, {}
  L = _lib_module().lib()

  def device_xyz(lane, xyz):
    n_frames, per_frame = xyz.shape[0], int(xyz.shape[1]) * 3
    per = -(-n_frames // world)
    dtype = torch.float64 if xyz.dtype == np.float64 else torch.float32
    need = per * world * per_frame
    buf = xyz_bufs.get(lane)
    if buf is None or buf.numel() < need or buf.dtype != dtype:
      buf = xyz_bufs[lane] = torch.empty(need, dtype=dtype, device="cuda")
      stage_bufs[lane] = torch.empty(per * per_frame, dtype=dtype, device="cuda")
    stage = stage_bufs[lane]
    if stage.numel() < per * per_frame:
      stage = stage_bufs[lane] = torch.empty(per * per_frame, dtype=dtype, device="cuda")
    f0 = min(rank * per, n_frames)
    f1 = min(f0 + per, n_frames)
    if f1 > f0:
      stage[:(f1 - f0) * per_frame].copy_(torch.from_numpy(xyz[f0:f1]).reshape(-1), non_blocking=True)
    dist.all_gather_into_tensor(buf[:need], stage[:per * per_frame], group=group)
    # the lane starts this trajectory's kernels after what torch's stream holds now: the all-gather
    _lib_module().check(L.nrl_lane_fork(lane, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return buf.data_ptr()

  total = {"trajectories": 0, "tasks": 0, "native_tasks": 0, "frames": 0, "seconds": 0.0}
  for c0 in range(0, len(indices), depth):
    chunk = indices[c0:c0 + depth]          # one lane (and its device buffer) per trajectory of the chunk
    # (a lane's device buffer holds a trajectory's parts until they were sent: no lane may be used twice inside a chunk,
    #  so the byte budget that narrows the pipeline for large trajectories is lifted for this run)
    budget, featurizer.pipeline_bytes = featurizer.pipeline_bytes, 1 << 62
    stats = featurizer.run(trajectory_indices=chunk, task_shard=(rank, world, n_slabs), shard_device_out=device_out,
                           shard_device_xyz=device_xyz if gather_uploads else None)
    featurizer.pipeline_bytes = budget
    for key in total:
      total[key] += stats.get(key, 0)
    local = featurizer.SHARD_RESULTS
    featurizer.SHARD_RESULTS = {}
    generic = {}
    if any(info["has_generic"] for info in local.values()):      # the same answer on every rank: same trajectories, same features
      bucket = [None] * world if rank == 0 else None
      dist.gather_object({tid: info["generic"] for tid, info in local.items()}, bucket, dst=0, group=group)
      for part in bucket or []:
        for tid, items in part.items():
          generic.setdefault(tid, []).extend(items)
    parts = {}
    for tid in chunk:
      info = local.get(tid)
      if info is None:
        continue                              # skipped (no focal point): on every rank alike
      if rank != 0:
        if info["floats"]:
          dist.send(lane_bufs[info["lane"]][:info["floats"]], dst=0, group=group)
        continue
      # rank 0: whole grids of this trajectory in a page-locked buffer, every slab copied to its rows
      dims, native = info["dims"], info["native_mask"]
      offsets, n_whole = {}, 0
      for r in range(world):
        for (bid, pid, fidx) in info["all_parts"][r][0]:
          if native[fidx] and (bid, pid, fidx) not in offsets:
            offsets[(bid, pid, fidx)] = n_whole
            n_whole += int(np.prod(dims[fidx]))
      if state["whole"] is None or state["whole"].numel() < n_whole:
        state["whole"] = torch.empty(max(n_whole, 1), dtype=torch.float32).pin_memory()
      whole = state["whole"]
      for r in range(world):
        keys, slabs = info["all_parts"][r]
        sizes = [(xc or dims[fidx][0]) * dims[fidx][1] * dims[fidx][2] if native[fidx] else 0 for (_, _, fidx), (_, xc) in zip(keys, slabs)]
        n_r = int(sum(sizes))
        if n_r == 0:
          continue
        if r == 0:
          src = lane_bufs[info["lane"]]
        else:
          if state["recv"] is None or state["recv"].numel() < n_r:
            state["recv"] = torch.empty(int(n_r * 1.25), dtype=torch.float32, device="cuda")
          src = state["recv"]
          dist.recv(src[:n_r], src=r, group=group)
        pos = 0
        for (bid, pid, fidx), (xb, _), size in zip(keys, slabs, sizes):
          if size == 0:
            continue
          at = offsets[(bid, pid, fidx)] + xb * dims[fidx][1] * dims[fidx][2]
          whole[at:at + size].copy_(src[pos:pos + size], non_blocking=True)
          pos += size
        torch.cuda.current_stream().synchronize()       # `recv` is reused for the next rank
      host = whole.numpy()
      items = [(bid, pid, fidx, 0, host[at:at + int(np.prod(dims[fidx]))].reshape(dims[fidx])) for (bid, pid, fidx), at in offsets.items()]
      featurizer.deliver_shard_results({tid: items + generic.get(tid, [])})
    torch.cuda.synchronize()                  # the lanes' device buffers are written again by the next chunk
  return total
