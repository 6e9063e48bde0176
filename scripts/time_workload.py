"""Wall time and per-kernel breakdown of one workload through Featurizer.run():
  python scripts/time_workload.py misato|misato_production[_labels]|stress|stress_native|dynamic|readme|pdbbind|residue_single|residue_dual [n_traj] [reps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import nearl_b200
from nearl_b200 import _lib, workloads
nearl_b200.update_config(quiet=True)
name = sys.argv[1]
n_traj = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
make = {"misato": workloads.misato_mixed, "dynamic": workloads.dynamic_observers, "readme": workloads.readme_mass,
        "stress": workloads.stress_large, "stress_native": lambda **kw: workloads.stress_large(native_atoms=True, **kw)}
loop = "run"
if name in ("residue_single", "residue_dual"):
  fz, trajs = workloads.residue_classification(n_traj=n_traj)
  loop = name.split("_")[1]
elif name in ("misato_production", "misato_production_labels"):
  table = None
  if name.endswith("labels"):
    import tempfile
    table = os.path.join(tempfile.mkdtemp(), "table.csv")
    with open(table, "w") as f:
      f.write("pdbcode,pK\n" + "".join(f"{i}syn,{3.0 + 0.5 * i}\n" for i in range(1, 10)))
  fz, trajs = workloads.misato_production(n_traj=n_traj, label_table=table)
elif name == "pdbbind":
  fz, trajs = workloads.pdbbind_static(n_complex=n_traj)
else:
  fz, trajs = make[name](n_traj=n_traj)
sums = {}
def sink(outkey, tid, bid, pid, grid):
  sums[outkey] = sums.get(outkey, 0.0) + float(np.sum(grid, dtype=np.float64))
fz.result_sink = (lambda *a: None) if os.environ.get("NRL_NOSUM") else sink      # NRL_NOSUM=1: throughput without the checksums
run = (lambda: fz.run()) if loop == "run" else (lambda: fz.loop_by_residue(loop))
run()                          # warm-up: arenas, templates
_lib.check(_lib.lib().nrl_synchronize())
times = []
for r in range(reps):
  sums.clear()
  if r == reps - 1:
    _lib.profile(True)          # lane 0 only: every other trajectory
  t0 = time.perf_counter()
  stats = run()
  times.append(time.perf_counter() - t0)
prof = _lib.profile_read(reset=True)
frames = stats.get("frames", stats.get("blocks", 0))
best = min(times)
print(json.dumps({"workload": name, "n_traj": n_traj, "tasks": stats["tasks"], "frames": frames, "seconds_best": best,
                  "frames_per_s": frames / best, "tasks_per_s": stats["tasks"] / best,
                  "kernels_ms_last_rep": {k: round(v[0], 3) for k, v in prof.items()},
                  "sums": {k: round(v, 3) for k, v in sums.items()}}))
