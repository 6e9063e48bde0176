"""One process per GPU: featurise a sharded trajectory list, bring the per-trajectory results to rank 0 in
order, write ONE HDF5 file there and compare it with a single-GPU run of the whole list.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/check_multi_gpu.py"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import nearl_b200
from nearl_b200 import features, featurizer, io, sharding, utils
nearl_b200.update_config(quiet=True)
rank, world = sharding.rank_world()
local_rank = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

def make(outfile=None):
  trajs = [io.synthetic_trajectory(4000, 20, seed=500 + i, pin=False) for i in range(7)]
  fz = featurizer.Featurizer({"dimensions": 24, "lengths": 14, "time_window": 10, "cutoff": 4.0, "sigma": 1.3})
  fz.register_features([features.Mass(outkey="mass", outfile=outfile),
                        features.MarchingObservers(obs="distinct_count", agg="mean", weight_type="atomic_number", outkey="distinct", outfile=outfile),
                        features.DensityFlow(agg="standard_deviation", weight_type="mass", outkey="flow", outfile=outfile)])
  fz.register_focus([":ARG"], "mask")
  fz.register_trajloader(trajs)
  return fz

# sharded run: results kept per trajectory
fz = make()
per_traj = {}
fz.result_sink = lambda key, tid, bid, pid, grid: per_traj.setdefault(tid, []).append((key, np.array(grid)))
stats, mine = sharding.run_sharded(fz, rank, world, device=local_rank)
merged = sharding.gather_in_order(per_traj, fz.TRAJECTORYNUMBER)
ok = True
if rank == 0:
  out = os.path.join(tempfile.mkdtemp(), "gathered.h5")
  with utils.batched_output():
    for tid, rows in enumerate(merged):
      for key, grid in rows:
        utils.append_hdf_data(out, key, grid[None], dtype=np.float32, maxshape=(None, 24, 24, 24), chunks=(1, 24, 24, 24))
  ref_file = os.path.join(tempfile.mkdtemp(), "single.h5")
  single = make(outfile=ref_file)
  single.run()
  with utils.open_hdf(out, "r") as a, utils.open_hdf(ref_file, "r") as b:
    for key in ("mass", "distinct", "flow"):
      x, y = a[key][:], b[key][:]
      same = x.shape == y.shape and np.array_equal(x, y)
      ok &= same
      print(key, x.shape, "identical to the single-GPU file:", same)
  print("shards:", [sharding.shard_indices(7, r, world) for r in range(world)])

# ---- one trajectory over all GPUs: (window, focal point) units and grid slabs, parts sent to rank 0 over NCCL ----------
class Custom(features.Mass):                 # a user subclass: host protocol, goes whole to the owner of the unit's first slab
  def run(self, coords, weights):
    return features.Mass.run(self, coords, weights) * 2.0

def make_big(outfile=None):
  trajs = [io.synthetic_trajectory(6000, 12, seed=600 + i, pin=False) for i in range(3)]
  fz = featurizer.Featurizer({"dimensions": 48, "lengths": 24, "time_window": 6, "cutoff": 4.0, "sigma": 1.3,
                              **({"outfile": outfile} if outfile else {})})
  fz.register_features([features.Mass(outkey="mass"),
                        features.MarchingObservers(obs="existence", agg="mean", outkey="exist"),
                        features.MarchingObservers(obs="mean_distance", agg="standard_deviation", weight_type="mass", outkey="meandist"),
                        features.DensityFlow(agg="mean", weight_type="mass", outkey="flow"),
                        Custom(outkey="custom", dimensions=16, lengths=8.0)])
  fz.register_focus([":ARG"], "mask")
  fz.register_trajloader(trajs)
  return fz

tmp = [tempfile.mkdtemp() if rank == 0 else None]
dist.broadcast_object_list(tmp, src=0)
shard_file = os.path.join(tmp[0], "task_sharded.h5")
fzs = make_big(outfile=shard_file if rank == 0 else None)
for feat in fzs.FEATURESPACE:
  feat.outfile = shard_file if rank == 0 else None
sstats = sharding.run_task_sharded(fzs, rank, world, device=local_rank)
if rank == 0:
  ref_file = os.path.join(tmp[0], "task_single.h5")
  single = make_big(outfile=ref_file)
  single.run()
  with utils.open_hdf(shard_file, "r") as a, utils.open_hdf(ref_file, "r") as b:
    for key in ("mass", "exist", "meandist", "flow", "custom"):
      x, y = a[key][:], b[key][:]
      same = x.shape == y.shape and np.array_equal(x, y)
      ok &= same
      print("task-sharded", key, x.shape, "identical to the single-GPU file:", same)
  print("parts of rank 0..:", [sharding.task_parts(2, r, world) for r in range(world)])
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
