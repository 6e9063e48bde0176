"""Where does an end-to-end Featurizer.run step spend its time? (GPU box)
  python scripts/e2e_breakdown.py [workload]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import nearl_b200
from nearl_b200 import _lib, workloads
nearl_b200.update_config(quiet=True)
name = sys.argv[1] if len(sys.argv) > 1 else "dynamic"
make = {"dynamic": workloads.dynamic_observers, "misato": workloads.misato_mixed, "readme": workloads.readme_mass}
fz, trajs = make[name](n_traj=4)
fz.keep_results = True
for i in range(4): fz.run(trajectory_indices=[i])
T = {}
def tick(name, t0): T[name] = T.get(name, 0.0) + time.perf_counter() - t0
N = 20
for it in range(N):
  traj = trajs[it % 4]
  t0 = time.perf_counter(); fz.traj = traj; fz.parse_focus(); tick("focus", t0)
  t0 = time.perf_counter()
  for f in fz.FEATURESPACE: f.cache(traj)
  specs = [f._native_spec() for f in fz.FEATURESPACE]; tick("cache+spec", t0)
  t0 = time.perf_counter(); batch = fz.lower_batch(list(range(fz.FEATURENUMBER)), specs, traj.n_atoms, False); tick("lower", t0)
  out = fz._out_buffers[0][:batch.out_floats]
  t0 = time.perf_counter(); batch.execute(_lib.ptr(traj.xyz), traj.n_frames, _lib.ptr(out), no_sync=True); tick("native call: enqueue (host)", t0)
  t0 = time.perf_counter(); _lib.check(_lib.lib().nrl_lane_synchronize(0)); tick("native call: wait (H2D+kernels+D2H)", t0)
  t0 = time.perf_counter(); res = batch.split(out); tick("split", t0)
fz.keep_results = False
fz.result_sink = lambda *a: None
t0 = time.perf_counter()
fz.run(trajectory_indices=[it % 4 for it in range(N)])
tick("full run() per trajectory, pipelined", t0)
for k, v in T.items(): print(f"{k:44s} {v / N * 1e3:8.3f} ms")
x = torch.from_numpy(trajs[0].xyz); d = torch.empty_like(x, device="cuda"); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); print("H2D of one trajectory, pinned", (time.perf_counter() - t0) / 10 * 1e3, "ms", x.is_pinned(), x.numel() * 4 / 1e6, "MB")
