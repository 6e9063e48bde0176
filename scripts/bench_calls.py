"""Per-call wall time of the six FFI twins against the reference's own kernels (oracle/_ref, built
unmodified from /root/reference/src for sm_100) at the shapes of BASELINE.md section 3, numpy in -> numpy out,
exactly as nearl/commands.py calls the module.  python scripts/bench_calls.py > profiles/<round>_calls.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import numpy as np
import torch
from nearl_b200 import all_actions as nb
try:
  import all_actions as ref
  probe = ref.frame_voxelize(np.full((4, 3), 4.0, dtype=np.float32), np.ones(4, dtype=np.float32), np.array([8, 8, 8]), 1.0, 2.0, 1.0, 0)
  if not np.any(probe):
    ref = None
except Exception:
  ref = None
_keep = torch.zeros(1 << 20, device="cuda")        # a live allocation keeps the reference's cudaMalloc/cudaFree cheap (its best case)
rng = np.random.default_rng(20240713)

def best(fn, reps):
  fn()
  ts = []
  for _ in range(reps):
    t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
  return min(ts)

def coords(F, N, D, s, c):
  x = rng.uniform(-c, D * s + c, size=(F, N, 3)).astype(np.float32)
  return x if F > 1 else x[0]

rows = []
def case(cid, name, call_new, call_ref, work, reps=7):
  t_new = best(call_new, reps)
  t_ref = best(call_ref, max(2, reps // 2)) if ref is not None else None
  rows.append({"id": cid, "call": name, "work_pairs": work, "nearl_b200_ms": t_new * 1e3, "reference_ms": None if t_ref is None else t_ref * 1e3,
               "speedup": None if t_ref is None else t_ref / t_new})

def dims(D):
  return np.array([D, D, D], dtype=np.int32)

# V1-V4: frame_voxelize
for cid, D, s, c, sg, N in (("V1", 32, 0.5, 5.0, 1.5, 1750), ("V2", 32, 1.0, 2.5, 1.0, 300), ("V3", 24, 20 / 24, 2.5, 1.5, 1560), ("V4", 128, 0.5, 5.0, 1.5, 9999)):
  x, w = coords(1, N, D, s, c), rng.uniform(1, 16, size=N).astype(np.float32)
  case(cid, f"frame_voxelize D={D} s={s:.3g} c={c} N={N}", lambda: nb.frame_voxelize(x, w, dims(D), s, c, sg, 0),
       lambda: ref.frame_voxelize(x, w, np.array([D, D, D]), s, c, sg, 0), D ** 3 * N, reps=5 if D == 128 else 9)
# O1-O5, O7, O8: marching_observer
for cid, obs, D, s, c, F, N in (("O1", 1, 32, 0.5, 5.0, 10, 1750), ("O2", 2, 32, 0.5, 5.0, 10, 1750), ("O3", 3, 32, 0.5, 5.0, 10, 1750),
                                ("O4", 11, 32, 0.5, 5.0, 10, 1750), ("O5", 12, 32, 0.5, 5.0, 10, 1750), ("O6", 14, 32, 0.5, 5.0, 10, 600),
                                ("O7", 1, 32, 1.0, 2.5, 50, 300), ("O8", 1, 128, 0.5, 5.0, 100, 9999)):
  x = coords(F, N, D, s, c)
  w = (rng.integers(1, 9, size=F * N) if obs == 3 else rng.uniform(1, 16, size=F * N)).astype(np.float32)
  case(cid, f"marching_observer obs={obs} D={D} F={F} N={N}", lambda: nb.marching_observer(x, w, dims(D), s, c, obs, 1),
       lambda: ref.marching_observer(x, w, np.array([D, D, D]), s, c, obs, 1), D ** 3 * N * F, reps=3 if D == 128 else 7)
# P1-P2: density_flow
for cid, agg in (("P1", 1), ("P2", 2)):
  x, w = coords(10, 1750, 32, 0.5, 5.0), rng.uniform(1, 16, size=17500).astype(np.float32)
  case(cid, f"density_flow agg={agg} D=32 F=10 N=1750", lambda: nb.density_flow(x, w, dims(32), 0.5, 5.0, 1.5, agg),
       lambda: ref.density_flow(x, w, np.array([32, 32, 32]), 0.5, 5.0, 1.5, agg), 32 ** 3 * 1750 * 10)
# A1: aggregate
for F, D in ((50, 32), (100, 128)):
  series = rng.random((F, D ** 3), dtype=np.float32)
  for agg in (1, 2, 7) if D == 128 else range(1, 9):
    case("A1", f"aggregate agg={agg} [{F}, {D}^3]", lambda: nb.aggregate(series, agg), lambda: ref.aggregate(series, agg), F * D ** 3, reps=3 if D == 128 else 7)
print(json.dumps({"reference": "oracle/_ref (reference kernels, sm_100 build)" if ref is not None else None,
                  "note": "host numpy arrays in -> host numpy array out per call; best of several calls; reference context kept warm", "rows": rows}, indent=1))
