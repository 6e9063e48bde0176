"""File-backed MISATO-style run (the step before the hot path): n complexes of 4 500 solute atoms, each with
a solvated parm7 (45 000 atoms, gzip) and 100 float32 frames in a gzip-chunked MD.hdf5, through the headline feature set with the
loader's read-ahead threads off and on.  python scripts/profile_ingestion.py [n] > profiles/<round>_ingestion.json"""
import gzip, json, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import nearl_b200
from nearl_b200 import featurizer, features, io, utils
nearl_b200.update_config(quiet=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(0)
root = tempfile.mkdtemp(prefix="misato_")


def parm7_text(solute, n_wat):
  top = solute.top
  n_at = top.n_atoms + 3 * n_wat
  names = [str(x)[:4] for x in top.names] + ["O", "H1", "H2"] * n_wat
  blocks = top.residue_blocks()
  first = [b[1] + 1 for b in blocks] + list(range(top.n_atoms + 1, n_at + 1, 3))
  labels = [b[0] for b in blocks] + ["WAT"] * n_wat
  z = [int(v) for v in top.atomic_numbers] + [8, 1, 1] * n_wat

  def sec(flag, fmt, vals, per, f):
    return [f"%FLAG {flag}", f"%FORMAT({fmt})"] + ["".join(f(v) for v in vals[i:i + per]) for i in range(0, len(vals), per)]
  I, E, A = (lambda v: f"{int(v):8d}"), (lambda v: f"{float(v):16.8E}"), (lambda v: f"{v:<4s}")
  ptr = [0] * 31
  ptr[0], ptr[11] = n_at, len(first)
  lines = ["%VERSION  VERSION_STAMP = V0001.000  DATE = 01/01/24  00:00:00", "%FLAG TITLE", "%FORMAT(20a4)", "synthetic"]
  lines += sec("POINTERS", "10I8", ptr, 10, I) + sec("ATOM_NAME", "20a4", names, 20, A)
  lines += sec("CHARGE", "5E16.8", rng.normal(size=n_at) * 5, 5, E) + sec("ATOMIC_NUMBER", "10I8", z, 10, I)
  lines += sec("MASS", "5E16.8", np.full(n_at, 12.01), 5, E) + sec("ATOM_TYPE_INDEX", "10I8", [1] * n_at, 10, I)
  lines += sec("NUMBER_EXCLUDED_ATOMS", "10I8", [3] * n_at, 10, I)
  lines += sec("RESIDUE_LABEL", "20a4", labels, 20, A) + sec("RESIDUE_POINTER", "10I8", first, 10, I)
  # the bonded terms and exclusion lists that make up most of a solvated system's file
  for flag, count in (("BONDS_INC_HYDROGEN", 3 * n_at), ("ANGLES_INC_HYDROGEN", 4 * top.n_atoms), ("DIHEDRALS_INC_HYDROGEN", 10 * top.n_atoms),
                      ("EXCLUDED_ATOMS_LIST", 3 * n_at)):
    lines += sec(flag, "10I8", rng.integers(0, 3 * n_at, size=count), 10, I)
  lines += sec("RADII", "5E16.8", np.full(n_at, 1.5), 5, E) + sec("SCREEN", "5E16.8", np.full(n_at, 0.8), 5, E)
  return "\n".join(lines) + "\n"


codes = [f"{1 + i % 9}{chr(97 + i // 9 % 26)}{chr(97 + i % 26)}{chr(97 + i // 26 % 26)}" for i in range(n)]
with utils.open_hdf(os.path.join(root, "MD.hdf5"), "a") as hdf:
  for i, code in enumerate(codes):
    solute = io.synthetic_trajectory(4500, 100, seed=100 + i, dtype=np.float32, pin=False)
    d = os.path.join(root, "parameter_restart_files_MD", code)
    os.makedirs(d)
    with gzip.open(os.path.join(d, "production.top.gz"), "wt") as fh:
      fh.write(parm7_text(solute, 13500))
    hdf.create_group(code.upper()).create_dataset("trajectory_coordinates", data=np.asarray(solute.xyz, dtype=np.float32),
                                                  chunks=(10, solute.n_atoms, 3), compression="gzip", compression_opts=4)

rows = {"complexes": n, "solute_atoms": 4500, "topology_atoms": 45000, "frames": 100,
        "md_hdf5_mb": os.path.getsize(os.path.join(root, "MD.hdf5")) / 1e6}
t = time.perf_counter()
for code in codes[:8]:
  io.load_amber_topology(os.path.join(root, "parameter_restart_files_MD", code, "production.top.gz"))
rows["topology_ms"] = (time.perf_counter() - t) / 8 * 1e3
t = time.perf_counter()
for code in codes[:8]:
  io.MisatoTraj.read_raw(code, root)
rows["read_raw_ms"] = (time.perf_counter() - t) / 8 * 1e3
for workers, to_file in ((0, False), (2, False), (4, False), (8, False), (0, False), (4, False), (0, True), (8, True)):
  loader = io.TrajectoryLoader([(c, root) for c in codes], trajtype=io.MisatoTraj, superpose=True)
  out = os.path.join(root, f"out_{workers}.h5")
  fz = featurizer.Featurizer({"dimensions": 32, "lengths": 16, "time_window": 10, "cutoff": 5.0, "sigma": 1.5, "outfile": out} if to_file else
                             {"dimensions": 32, "lengths": 16, "time_window": 10, "cutoff": 5.0, "sigma": 1.5})
  fz.register_features([features.MarchingObservers(obs="existence", agg="mean", outkey="e"),
                        features.MarchingObservers(obs="distinct_count", agg="mean", outkey="d"),
                        features.MarchingObservers(obs="mean_distance", agg="mean", outkey="m"),
                        features.DensityFlow(agg="mean", outkey="f")])
  fz.register_focus([":ARG"], "mask")
  fz.register_trajloader(loader)
  if not to_file:
    fz.result_sink = lambda *a: None
  fz.prefetch_workers = workers
  t = time.perf_counter()
  stats = fz.run()
  dt = time.perf_counter() - t
  rows.setdefault("runs", []).append({"loader_threads": workers, "sink": "HDF5 file (gzip level 0, the default)" if to_file else "none",
                                      "ms_per_trajectory": dt / n * 1e3, "frames_per_s": stats["frames"] / dt,
                                      "trajectories": stats["trajectories"], "tasks": stats["tasks"],
                                      "out_mb": os.path.getsize(out) / 1e6 if to_file else None})
  print(rows["runs"][-1], file=sys.stderr, flush=True)
if not os.environ.get("NRL_INGEST_KEEP"):
  shutil.rmtree(root, ignore_errors=True)
print(json.dumps(rows, indent=1))
