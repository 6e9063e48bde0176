"""Soak: many trajectories of ragged sizes through Featurizer.run with eight (then five) lanes in flight, every result compared bit for
bit with a serial single-lane run (buffer re-use, lane hand-over, arena growth).  python scripts/soak_pipeline.py [n_traj]"""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, nearl_b200
from nearl_b200 import features, featurizer, io
nearl_b200.update_config(quiet=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(5)
trajs = [io.synthetic_trajectory(int(rng.integers(200, 9000)), int(rng.choice([4, 8, 12, 20])), seed=100 + i, pin=bool(i % 2),
                                 dtype=(np.float64 if i % 7 == 0 else np.float32), ligand_atoms=int(rng.choice([0, 20]))) for i in range(n)]
def run(depth):
  fz = featurizer.Featurizer({"dimensions": 16, "lengths": 12, "time_window": 4, "cutoff": 3.5, "sigma": 1.2})
  fz.register_features([
    features.Mass(outkey="mass"), features.HeavyAtom(outkey="heavy", byres=BYRES, selection="!@H="),
    features.MarchingObservers(obs="existence", agg="mean", outkey="e"),
    features.MarchingObservers(obs="distinct_count", agg="max", weight_type="atomic_number", outkey="d"),
    features.MarchingObservers(obs="mean_distance", agg="standard_deviation", weight_type="mass", outkey="m"),
    features.MarchingObservers(obs="cumulative_weight", agg="information_entropy", weight_type="mass", outkey="c"),
    features.DensityFlow(agg="mean", weight_type="mass", outkey="f"), features.DensityFlow(agg="drift", weight_type="mass", selection=":ARG,LYS", outkey="g"),
  ])
  fz.register_focus([":ARG", "@CA"], "mask")
  fz.register_trajloader(trajs)
  fz.pipeline_depth = depth
  sums = []
  fz.result_sink = lambda key, tid, bid, pid, grid: sums.append((key, tid, bid, pid, zlib.crc32(np.ascontiguousarray(grid).tobytes())))
  fz.run()
  return sums
ok = True
for BYRES in (True, False):        # without a by-residue crop the batches take the streaming crop kernel (first crop kernel: crop_legacy)
  a = run(8); b = run(1); c = run(5)
  from nearl_b200 import _lib
  _lib.set_option("crop_legacy", 1); d = run(8); _lib.set_option("crop_legacy", 0)
  print("by-residue" if BYRES else "streaming crop", "results", len(a), "pipelined == serial:", a == b, "; pipelined twice:", a == c,
        "; first crop kernel gives the same:", a == d)
  ok &= a == b and a == c and a == d
sys.exit(0 if ok else 1)
