"""How many voxels of the single-frame observables (index-ordered scan kernel) are BIT-identical to the reference's own
kernels (oracle/_ref)?  A restatement that uses the reference's arithmetic op for op scores 100 %.  (GPU box)
  python scripts/bitwise_vs_reference.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import numpy as np
from nearl_b200 import all_actions as nb
import all_actions as ref
rng = np.random.default_rng(1)
D, s, c, sig = 20, 0.75, 3.5, 1.5
dims = np.array([D, D, D], dtype=np.int32)
names = {1: "existence", 2: "direct_count", 3: "distinct_count", 11: "mean_distance", 12: "cumulative_weight", 13: "density",
         14: "dispersion", 15: "eccentricity", 16: "radius_of_gyration"}
for N in (300, 2500):
  coords = rng.uniform(-c, D * s + c, size=(N, 3)).astype(np.float32)
  w = rng.uniform(0.5, 16.0, size=N).astype(np.float32)
  for obs, name in names.items():
    g = nb.frame_observation(coords, w, dims, s, c, obs); r = ref.frame_observation(coords, w, np.array(dims, dtype=int), s, c, obs)
    same = np.count_nonzero(g.view(np.int32) == r.view(np.int32)); close = np.count_nonzero(np.isclose(g, r, rtol=1e-5, atol=1e-6))
    print(f"N={N:5d} {name:20s} bit-identical {same / g.size:8.4%}   within 1e-5 {close / g.size:8.4%}  max|d| {np.abs(g - r).max():.3g}")
  g = nb.frame_voxelize(coords, w, dims, s, c, sig, 0); r = ref.frame_voxelize(coords, w, np.array(dims, dtype=int), s, c, sig, 0)
  print(f"N={N:5d} {'frame_voxelize':20s} bit-identical {np.count_nonzero(g.view(np.int32) == r.view(np.int32)) / g.size:8.4%} (the reference deposits with float atomics: its own order varies)")
F = 10
coords = rng.uniform(-c, D * s + c, size=(F, 600, 3)).astype(np.float32); w = rng.uniform(0.5, 16.0, size=F * 600).astype(np.float32)
x = nb.marching_observer(coords, w, dims, s, c, 12, 1)
series = np.stack([ref.frame_observation(coords[i], w[i * 600:(i + 1) * 600], np.array(dims, dtype=int), s, c, 11) for i in range(F)])
for agg, name in ((1, "mean"), (2, "standard_deviation"), (3, "median"), (4, "variance"), (5, "max"), (6, "min"), (7, "information_entropy"), (8, "drift")):
  g = nb.aggregate(series, agg); r = ref.aggregate(series, agg)
  print(f"aggregate {name:20s} bit-identical {np.count_nonzero(g.view(np.int32) == r.view(np.int32)) / g.size:8.4%}  max|d| {np.abs(g - r).max():.3g}")
