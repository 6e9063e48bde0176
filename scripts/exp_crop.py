"""Experiment: crop kernel alone, streaming form against the first form (device-resident trajectory, per-kernel events).
  python scripts/exp_crop.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import nearl_b200
from nearl_b200 import _lib, workloads
nearl_b200.update_config(quiet=True)
for name, make, kw in (("dynamic", workloads.dynamic_observers, {}), ("stress", workloads.stress_large, {})):
  fz, trajs = make(n_traj=1, **kw)
  traj = trajs[0]
  fz.traj = traj; fz.parse_focus()
  for f in fz.FEATURESPACE: f.cache(traj)
  specs = [f._native_spec() for f in fz.FEATURESPACE]
  batch = fz.lower_batch(list(range(fz.FEATURENUMBER)), specs, traj.n_atoms, False)
  xyz = torch.from_numpy(traj.xyz).cuda()
  res = torch.empty(batch.out_floats, dtype=torch.float32, device="cuda")
  n_crop_frames = sum(1 for _ in range(1))
  for legacy in (1, 0):
    _lib.set_option("crop_legacy", legacy)
    for _ in range(3):
      batch.execute(xyz.data_ptr(), traj.n_frames, res.data_ptr(), xyz_on_device=True, out_on_device=True)
    _lib.profile(True); _lib.profile_read(reset=True)
    for _ in range(10):
      batch.execute(xyz.data_ptr(), traj.n_frames, res.data_ptr(), xyz_on_device=True, out_on_device=True)
    prof = _lib.profile_read(reset=True); _lib.profile(False)
    ms, n = prof["crop_pack"]
    frames = traj.n_frames if name == "dynamic" else 200
    nbytes = frames * traj.n_atoms * 12
    print(json.dumps({"workload": name, "crop": "first form" if legacy else "streaming", "us_per_launch": ms / n * 1e3,
                      "GB/s": nbytes / (ms / n * 1e-3) / 1e9, "bytes": nbytes, "sum": float(res.sum())}), flush=True)
  _lib.set_option("crop_legacy", 0)
