import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import numpy as np, nearl_b200
import all_actions as ref
from nearl_b200 import workloads
from oracle import host_ref
nearl_b200.update_config(quiet=True)
fz, trajs = workloads.readme_mass(n_traj=3)
k = host_ref.ModuleKernels(ref)
for traj in trajs:
  fz.traj = traj; fz.parse_focus()
  for f in fz.FEATURESPACE: f.cache(traj)
  descs = workloads.describe(fz); resids = fz.FEATURESPACE[0].resids
  t = time.perf_counter()
  host_ref.featurize(traj.xyz, resids, descs, fz.FOCALPOINTS.copy(), fz.time_window, kernels=k)
  print("featurize ms", (time.perf_counter() - t) * 1e3)
  for bid in range(3):
    t = time.perf_counter(); frames = traj.xyz[bid * 10:(bid + 1) * 10]; t1 = time.perf_counter()
    host_ref.run_task(descs[0], resids, frames, fz.FOCALPOINTS[bid, 0], kernels=k); t2 = time.perf_counter()
    print("  slice", bid, (t1 - t) * 1e3, (t2 - t1) * 1e3, type(traj.xyz), traj.xyz.flags.c_contiguous)
