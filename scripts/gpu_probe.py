"""Ad-hoc per-call timing of the C-ABI entry points next to the reference's own kernels."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
from nearl_b200 import all_actions as nb
try:
  import all_actions as ref
except Exception as e:
  ref = None; print("no reference module", e)

def timeit(fn, n=20, warm=3):
  for _ in range(warm): fn()
  t = time.perf_counter()
  for _ in range(n): fn()
  return (time.perf_counter() - t) / n * 1e6

rng = np.random.default_rng(0)
dims = np.array([32, 32, 32], dtype=np.int32)
N, F, s, c, sig = 1750, 10, 0.5, 5.0, 1.5
coords = rng.uniform(-5, 21, size=(F, N, 3)).astype(np.float32)
w = rng.uniform(1, 16, size=F * N).astype(np.float32)
rows = {}
def both(name, f_new, f_ref, n=20):
  a = timeit(f_new, n)
  b = timeit(f_ref, max(3, n // 4)) if ref is not None else float("nan")
  rows[name] = {"new_us": a, "ref_us": b, "speedup": b / a}
  print(f"{name:28s} new {a:10.1f} us   ref {b:10.1f} us   x{b / a:6.1f}", flush=True)
both("V1 frame_voxelize", lambda: nb.frame_voxelize(coords[0], w[:N], dims, s, c, sig, 0), lambda: ref.frame_voxelize(coords[0], w[:N], dims, s, c, sig, 0))
for obs in (1, 2, 3, 11, 12):
  both(f"O marching_observer obs{obs}", lambda: nb.marching_observer(coords, w, dims, s, c, obs, 1), lambda: ref.marching_observer(coords, w, dims, s, c, obs, 1))
both("P1 density_flow agg1", lambda: nb.density_flow(coords, w, dims, s, c, sig, 1), lambda: ref.density_flow(coords, w, dims, s, c, sig, 1))
x = rng.random((50, 32768)).astype(np.float32)
both("A1 aggregate 50x32^3 agg1", lambda: nb.aggregate(x, 1), lambda: ref.aggregate(x, 1))
both("A1 aggregate 50x32^3 agg3", lambda: nb.aggregate(x, 3), lambda: ref.aggregate(x, 3))
c600 = coords[:, :600]
both("O6 dispersion N=600", lambda: nb.marching_observer(c600, w[:F*600], dims, s, c, 14, 1), lambda: ref.marching_observer(c600, w[:F*600], dims, s, c, 14, 1), n=4)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "probe_timing.json"), "w"), indent=1)
