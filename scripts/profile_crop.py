"""One device-resident step of a workload (for ncu -k regex:crop): python scripts/profile_crop.py dynamic|stress"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nearl_b200
from nearl_b200 import _lib, workloads
nearl_b200.update_config(quiet=True)
name = sys.argv[1] if len(sys.argv) > 1 else "dynamic"
fz, trajs = (workloads.stress_large if name == "stress" else workloads.dynamic_observers)(n_traj=1)
traj = trajs[0]
fz.traj = traj; fz.parse_focus()
for f in fz.FEATURESPACE: f.cache(traj)
specs = [f._native_spec() for f in fz.FEATURESPACE]
batch = fz.lower_batch(list(range(fz.FEATURENUMBER)), specs, traj.n_atoms, False)
xyz = torch.from_numpy(traj.xyz).cuda()
res = torch.empty(batch.out_floats, dtype=torch.float32, device="cuda")
for _ in range(2):
  batch.execute(xyz.data_ptr(), traj.n_frames, res.data_ptr(), xyz_on_device=True, out_on_device=True)
torch.cuda.synchronize()
print("ok")
