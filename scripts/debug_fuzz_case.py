"""One case of tests/test_fuzz_gpu.py under the microscope (GPU box): the entropy of a marching-observer feature on the first
(window, focal point) task through this library, the CPU oracle and the reference's own kernels (oracle/_ref), with the per-frame
series of the voxels that differ.  python scripts/debug_fuzz_case.py <seed> <outkey>"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import nearl_b200
from nearl_b200 import featurizer, features, _lib, all_actions as nb
from oracle import host_ref, oracle as orc
import test_fuzz_gpu as T
nearl_b200.update_config(quiet=True)
seed = int(sys.argv[1]); fkey = sys.argv[2]
traj, parms, specs, focus, fmt = T.random_case(seed)
fz = featurizer.Featurizer(parms); fz.register_features([c(**k) for c, k in specs]); fz.register_focus(focus, fmt); fz.register_trajloader([traj]); fz.traj = traj
fz.parse_focus()
for f in fz.FEATURESPACE: f.cache(traj)
fi = [i for i, f in enumerate(fz.FEATURESPACE) if f.outkey == fkey][0]
f = fz.FEATURESPACE[fi]; d = T.descs_from(fz)[fi]
print(f, parms, "window", fz.time_window, "dtype", traj.xyz.dtype)
frames = traj.xyz[0:fz.time_window]
focal = fz.FOCALPOINTS[0, 0]
F = len(frames)
coords = np.full((F, 9999, 3), 99999.0, dtype=np.float32); weights = np.zeros((F, 9999), dtype=np.float32); mx = 0
for i, fr in enumerate(frames):
  mask, inbox = host_ref.base_query(d, f.resids, fr, focal)
  n = int(mask.sum()); coords[i, :n] = inbox; weights[i, :n] = np.asarray(d["weights"], dtype=np.float32)[mask]; mx = max(mx, n)
c = np.ascontiguousarray(coords[:, :mx]); w = np.ascontiguousarray(weights[:, :mx].reshape(-1))
dims = np.array(f.dims); s, cut = f.spacing, f.cutoff
g = nb.marching_observer(c, w, dims, s, cut, f.obs, 7); r = orc.marching_observer(c, w, dims, s, cut, f.obs, 7)
bad = np.flatnonzero(~np.isclose(g, r, atol=1e-3)); print("mismatch vs oracle", len(bad), "of", g.size)
sys.path.insert(0, "/root/repo/oracle/_ref")
import all_actions as ref
rr = ref.marching_observer(c, w, np.array(dims, dtype=int), float(s), float(cut), int(f.obs), 7)
print("mismatch vs the reference's own kernels", int(np.count_nonzero(~np.isclose(g, rr, atol=1e-3))), "; oracle vs reference kernels", int(np.count_nonzero(~np.isclose(r, rr, atol=1e-3))))
bad = np.flatnonzero(~np.isclose(g, rr, atol=1e-3))
pg = np.stack([nb.frame_observation(c[i], w[i*mx:(i+1)*mx], dims, s, cut, f.obs) for i in range(F)])
pr = np.stack([orc.frame_observation(c[i], w[i*mx:(i+1)*mx], dims, s, cut, f.obs) for i in range(F)])
for v in bad[:6]:
  print("voxel", v, "dev", g[v], "ora", r[v], "ref", rr[v]); print("  dev series", pg[:, v].tolist()); print("  ora series", pr[:, v].tolist())
