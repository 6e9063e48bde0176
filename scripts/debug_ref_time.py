import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import numpy as np, nearl_b200
import all_actions as ref
from nearl_b200 import workloads
from oracle import host_ref
nearl_b200.update_config(quiet=True)
fz, trajs = workloads.readme_mass()
traj = trajs[0]; fz.traj = traj; fz.parse_focus()
for f in fz.FEATURESPACE: f.cache(traj)
descs = workloads.describe(fz); d = descs[0]; resids = fz.FEATURESPACE[0].resids
def T(fn, n=5):
  fn(); t = time.perf_counter()
  for _ in range(n): fn()
  return (time.perf_counter() - t) / n * 1e3
mask, coords = host_ref.base_query(d, resids, traj.xyz[0], fz.FOCALPOINTS[0, 0])
w = np.asarray(d["weights"], dtype=np.float32)[mask]
print("atoms", len(coords), coords.dtype, coords.min(0), coords.max(0))
print("base_query ms", T(lambda: host_ref.base_query(d, resids, traj.xyz[0], fz.FOCALPOINTS[0, 0])))
c32 = coords.astype(np.float32)
print("ref voxelize ms", T(lambda: ref.frame_voxelize(c32, w, np.array([32, 32, 32]), 0.5, 5.0, 1.5, 0)))
rng = np.random.default_rng(0)
cu = rng.uniform(-5, 21, size=(len(coords), 3)).astype(np.float32)
print("ref voxelize uniform ms", T(lambda: ref.frame_voxelize(cu, w, np.array([32, 32, 32]), 0.5, 5.0, 1.5, 0)))
print("run_task ms", T(lambda: host_ref.run_task(d, resids, traj.xyz[0:10], fz.FOCALPOINTS[0, 0], kernels=host_ref.ModuleKernels(ref))))
