import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, nearl_b200
from nearl_b200 import _lib, workloads
from oracle import host_ref
nearl_b200.update_config(quiet=True)
def run():
  fz, trajs = workloads.readme_mass()
  fz.keep_results = True
  fz.run()
  return np.array(fz.RESULTS["mass"]), fz, trajs
a, fz, trajs = run()
b, _, _ = run()
print("deterministic:", np.array_equal(a, b))
ref = host_ref.featurize(trajs[0].xyz, fz.FEATURESPACE[0].resids, workloads.describe(fz), fz.FOCALPOINTS, 10, buff_dim=_lib.buff_dim)["mass"]
scale = np.maximum(np.abs(ref), 1e-5 * np.abs(ref).max())
bad = np.argwhere(np.abs(a - ref) > 1e-5 * scale)
print("bad:", len(bad))
for ix in bad[:10]:
  ix = tuple(ix); print(ix, a[ix], ref[ix], (a[ix] - ref[ix]) / ref[ix])
err = np.abs(a - ref) / scale
print("max rel err", err.max(), "p99.99", np.quantile(err, 0.9999))
# ---- which atom? legacy API on the cropped atoms, one at a time near the bad voxel
from nearl_b200 import all_actions as nb
from oracle import oracle as orc
f0 = fz.FEATURESPACE[0]
fz.traj = trajs[0]; fz.parse_focus(); f0.cache(trajs[0])
coords, w = f0.query(trajs[0].top, trajs[0].xyz[0:10], fz.FOCALPOINTS[0, 0])
print("cropped atoms", len(coords))
dims = np.array([32, 32, 32], dtype=np.int32)
full_new = nb.frame_voxelize(coords, w, dims, 0.5, 5.0, 1.5, 0).reshape(32, 32, 32)
full_ref = orc.frame_voxelize(coords, w, dims, 0.5, 5.0, 1.5, buff_dim=11).reshape(32, 32, 32)
print("legacy path at bad voxel:", full_new[0, 8, 9], full_ref[0, 8, 9], "batch:", a[0, 0, 8, 9])
p = np.array([0.0, 4.0, 4.5])
d = np.linalg.norm(coords - p, axis=1)
near = np.argsort(d)[:60]
for i in near:
  if d[i] > 5.3: break
  one_new = nb.frame_voxelize(coords[i:i+1], w[i:i+1], dims, 0.5, 5.0, 1.5, 0).reshape(32, 32, 32)
  one_ref = orc.frame_voxelize(coords[i:i+1], w[i:i+1], dims, 0.5, 5.0, 1.5, buff_dim=11).reshape(32, 32, 32)
  if abs(one_new[0, 8, 9] - one_ref[0, 8, 9]) > 1e-6 * max(one_ref.max(), 1e-9):
    print("atom", i, coords[i], w[i], d[i], one_new[0, 8, 9], one_ref[0, 8, 9], one_new.sum(), one_ref.sum())
