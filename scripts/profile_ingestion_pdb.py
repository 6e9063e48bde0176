"""File-backed PDBBind-style run (BASELINE configs[2] from files): n single-frame complexes as PDB files through the eight
static channels, the loader's read-ahead threads off and on.  python scripts/profile_ingestion_pdb.py [n]"""
import json, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import nearl_b200
from nearl_b200 import io, workloads
nearl_b200.update_config(quiet=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
root = tempfile.mkdtemp(prefix="pdbbind_")
fz, trajs = workloads.pdbbind_static(n_complex=n)
paths = []
for i, traj in enumerate(trajs):
  paths.append(os.path.join(root, f"c{i:05d}.pdb"))
  io.write_pdb(paths[-1], traj.xyz, traj.top)
rows = {"complexes": n, "atoms": int(trajs[0].n_atoms), "pdb_kb": os.path.getsize(paths[0]) / 1e3}
t = time.perf_counter()
for p in paths[:64]:
  io.load_pdb(p)
rows["load_pdb_ms"] = (time.perf_counter() - t) / 64 * 1e3
t = time.perf_counter()
for p in paths[:64]:
  io._load_pdb_lines(p)
rows["load_pdb_line_parser_ms"] = (time.perf_counter() - t) / 64 * 1e3
fz.result_sink = lambda *a: None
for workers in (0, 4, 8, 0, 8):
  fz.register_trajloader(io.TrajectoryLoader([(p,) for p in paths], dtype=np.float32))
  fz.prefetch_workers = workers
  t = time.perf_counter()
  stats = fz.run()
  dt = time.perf_counter() - t
  rows.setdefault("runs", []).append({"loader_threads": workers, "ms_per_complex": dt / n * 1e3, "complexes_per_s": n / dt, "tasks": stats["tasks"]})
  print(rows["runs"][-1], file=sys.stderr, flush=True)
shutil.rmtree(root, ignore_errors=True)
print(json.dumps(rows, indent=1))
