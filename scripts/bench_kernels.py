"""Device-resident micro-benchmarks of the HBM-bound kernels of the path against the measured copy
bandwidth (MEASURED_PEAKS.json): standalone K3 (`nrl_aggregate_dev`, all eight reducers), the crop
kernel (through nrl_featurize_batch's per-kernel events) and the superposition kernel.
  python scripts/bench_kernels.py > profiles/<round>_kernel_roofline.json"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import nearl_b200
from nearl_b200 import _lib, workloads
nearl_b200.update_config(quiet=True)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6650.0))
L = _lib.lib()
stream = torch.cuda.current_stream()
out = {"hbm_peak_gbs": hbm, "peak_source": "MEASURED_PEAKS.json (copy, read+write)" if peaks else "fallback", "kernels": []}

def timed(fn, reps=10):
  fn(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(stream)
  for _ in range(reps):
    fn()
  e1.record(stream); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps * 1e-3

# ---- K3 standalone: series [F, G] -> [G]
for F, D in ((50, 32), (100, 128)):
  G = D ** 3
  # several series so that consecutive launches do not find their input in the 126 MB L2
  n_rot = max(2, int(300e6 // (4 * F * G)) + 1)
  series = [torch.rand(F, G, device="cuda") for _ in range(min(n_rot, 6))]
  res = torch.empty(G, device="cuda")
  for agg, name in ((1, "mean"), (2, "standard_deviation"), (3, "median"), (5, "max"), (7, "information_entropy"), (8, "drift")):
    k = [0]
    def run():
      s = series[k[0] % len(series)]; k[0] += 1
      _lib.check(L.nrl_aggregate_dev(ctypes.c_void_p(s.data_ptr()), F, G, agg, ctypes.c_void_p(res.data_ptr()), ctypes.c_void_p(stream.cuda_stream)))
    t = timed(run, reps=6 if D == 128 else 30)
    nbytes = 4 * F * G + 4 * G
    out["kernels"].append({"kernel": "nrl_aggregate_kernel", "case": f"{name}, F={F}, G={D}^3", "algorithmic_bytes": nbytes, "seconds": t,
                           "achieved_gbs": nbytes / t / 1e9, "frac_of_hbm_peak": nbytes / t / 1e9 / hbm})
  del series

# ---- superposition: [F, N, 3] float32 in place
for F, N in ((100, 20000), (200, 200000)):
  xyz = torch.rand(F, N, 3, device="cuda") * 50
  idx = torch.arange(0, N, 16, dtype=torch.int32).numpy().copy()
  def run():
    _lib.check(L.nrl_superpose(ctypes.c_void_p(xyz.data_ptr()), 0, F, N, _lib.ptr(idx), len(idx), 0, 1, None, ctypes.c_void_p(stream.cuda_stream)))
  t = timed(run, reps=5)
  nbytes = 24 * F * N
  out["kernels"].append({"kernel": "nrl_superpose_kernel", "case": f"F={F}, N={N}, {len(idx)} fitted atoms", "algorithmic_bytes": nbytes, "seconds": t,
                         "achieved_gbs": nbytes / t / 1e9, "frac_of_hbm_peak": nbytes / t / 1e9 / hbm})
  del xyz

# ---- crop kernel: reads the frames of a trajectory once (12 bytes per atom per frame); device-resident, one lane
for name, n_atoms in (("dynamic", 20000), ("stress", 200000)):
  fz, trajs = (workloads.dynamic_observers(n_traj=2) if name == "dynamic" else workloads.stress_large(n_traj=1, n_frames=100))
  traj = trajs[0]
  fz.traj = traj
  fz.parse_focus()
  for f in fz.FEATURESPACE:
    f.cache(traj)
  specs = [f._native_spec() for f in fz.FEATURESPACE]
  batch = fz.lower_batch(list(range(fz.FEATURENUMBER)), specs, traj.n_atoms, False)
  xyz = torch.from_numpy(traj.xyz).cuda()
  res = torch.empty(batch.out_floats, dtype=torch.float32, device="cuda")
  go = lambda: batch.execute(xyz.data_ptr(), traj.n_frames, res.data_ptr(), xyz_on_device=True, out_on_device=True)
  go(); go()
  _lib.profile(True); _lib.profile_read(reset=True)
  for _ in range(5):
    go()
  prof = _lib.profile_read(reset=True); _lib.profile(False)
  ms, launches = prof["crop_pack"]
  nbytes = 12 * n_atoms * traj.n_frames
  t = ms / launches * 1e-3
  out["kernels"].append({"kernel": "nrl_crop_pack_cluster_kernel", "case": f"{name}: {n_atoms} atoms x {traj.n_frames} frames per launch",
                         "algorithmic_bytes": nbytes, "seconds": t, "achieved_gbs": nbytes / t / 1e9, "frac_of_hbm_peak": nbytes / t / 1e9 / hbm})
  del xyz, res
print(json.dumps(out, indent=1))
