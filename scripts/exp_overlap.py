"""Experiment: device-resident steps of the headline workload issued round-robin over L lanes of the library (each lane its own
stream and workspace), so that the crop / bin / tails of one step overlap the sweeps of another.
  python scripts/exp_overlap.py [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nearl_b200
from nearl_b200 import _lib, workloads
nearl_b200.update_config(quiet=True)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
fz, trajs = workloads.dynamic_observers(n_traj=8)
L = _lib.lib()
lowered, dev = [], []
for traj in trajs:
  fz.traj = traj
  fz.parse_focus()
  for f in fz.FEATURESPACE:
    f.cache(traj)
  specs = [f._native_spec() for f in fz.FEATURESPACE]
  lowered.append(fz.lower_batch(list(range(fz.FEATURENUMBER)), specs, traj.n_atoms, False))
  dev.append(torch.from_numpy(traj.xyz).cuda())
outs = [torch.empty(lowered[0].out_floats, dtype=torch.float32, device="cuda") for _ in range(4)]
n_frames = trajs[0].n_frames
for lanes in (1, 2, 3, 4):
  for rep in range(2):
    for l in range(lanes):
      _lib.check(L.nrl_lane_synchronize(l))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
      k, lane = i % len(trajs), i % lanes
      lowered[k].execute(dev[k].data_ptr(), n_frames, outs[lane].data_ptr(), xyz_on_device=True, out_on_device=True, no_sync=True, lane=lane)
    for l in range(lanes):
      _lib.check(L.nrl_lane_synchronize(l))
    dt = time.perf_counter() - t0
  print(f"lanes={lanes}: {dt / steps * 1e3:.4f} ms/step  {n_frames * steps / dt:.0f} frames/s", flush=True)
