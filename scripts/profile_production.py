import os, sys
sys.path.insert(0, "/root/repo")
import torch, nearl_b200
from nearl_b200 import _lib, workloads
nearl_b200.update_config(quiet=True)
fz, trajs = workloads.misato_production(n_traj=1)
traj = trajs[0]; fz.traj = traj; fz.parse_focus()
for f in fz.FEATURESPACE: f.cache(traj)
specs = [f._native_spec() for f in fz.FEATURESPACE]
batch = fz.lower_batch(list(range(fz.FEATURENUMBER)), specs, traj.n_atoms, False)
xyz = torch.from_numpy(traj.xyz).cuda(); out = torch.empty(batch.out_floats, dtype=torch.float32, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
  batch.execute(xyz.data_ptr(), traj.n_frames, out.data_ptr(), xyz_on_device=True, out_on_device=True)
torch.cuda.synchronize(); print("ok")
