"""Experiment: what bounds Featurizer.run() end to end on the headline workload?  depth sweep; host timeline.
  python scripts/exp_e2e.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import nearl_b200
from nearl_b200 import _lib, workloads, featurizer
nearl_b200.update_config(quiet=True)
name = sys.argv[1] if len(sys.argv) > 1 else "dynamic"
make = {"dynamic": workloads.dynamic_observers, "misato": workloads.misato_mixed, "readme": workloads.readme_mass,
        "misato_production": workloads.misato_production}
fz, trajs = make[name](n_traj=8)
touched = [0.0]
def sink(outkey, tid, bid, pid, grid):
  touched[0] += float(grid[0, 0, 0]) + float(grid[-1, -1, -1])
fz.result_sink = sink
K = 120
idx = [i % 8 for i in range(K)]
fz.run(trajectory_indices=idx)
for depth in ((3, 4, 5, 6, 7, 8) if name == "dynamic" else (1, 4, 8)):
  fz.pipeline_depth = depth
  fz.run(trajectory_indices=idx[:16])
  best = []
  for rep in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fz.run(trajectory_indices=idx)
    best.append(time.perf_counter() - t0)
  print(f"depth={depth}: " + " ".join(f"{100 * K / dt / 1e3:.0f}k" for dt in best) + " frames/s", flush=True)
# host-only cost of submit / finish: time spent inside, and time blocked in the lane sync
fz.pipeline_depth = 3
T = {"submit": 0.0, "finish": 0.0, "sync": 0.0, "focus": 0.0, "execute": 0.0, "lower": 0.0, "split": 0.0, "spec": 0.0}
real_submit, real_finish = fz._submit, fz._finish
real_sync = _lib.lib().nrl_lane_synchronize
def timed(name, fn):
  def w(*a, **k):
    t = time.perf_counter()
    try:
      return fn(*a, **k)
    finally:
      T[name] += time.perf_counter() - t
  return w
fz._submit = timed("submit", real_submit)
fz._finish = timed("finish", real_finish)
fz.parse_focus = timed("focus", fz.parse_focus)
fz.lower_batch = timed("lower", fz.lower_batch)
real_exec = featurizer.LoweredBatch.execute
featurizer.LoweredBatch.execute = timed("execute", real_exec)
real_split = featurizer.LoweredBatch.split
featurizer.LoweredBatch.split = timed("split", real_split)
for f in fz.FEATURESPACE:
  f._native_spec = timed("spec", f._native_spec)
class SyncProxy:
  def __init__(self, lib): self._lib = lib
  def __getattr__(self, n):
    f = getattr(self._lib, n)
    return timed("sync", f) if n == "nrl_lane_synchronize" else f
real_lib = _lib._lib
_lib._lib = SyncProxy(real_lib)
t0 = time.perf_counter()
fz.run(trajectory_indices=idx)
dt = time.perf_counter() - t0
_lib._lib = real_lib
featurizer.LoweredBatch.execute = real_exec
featurizer.LoweredBatch.split = real_split
print(f"instrumented: {dt / K * 1e3:.4f} ms/traj; per trajectory (us): " + ", ".join(f"{k} {v / K * 1e6:.0f}" for k, v in T.items()))
