"""K3 two-pass reducers: pipelined (ring of staged tiles) against the one-shot staged kernel, bit for bit and timed.  (GPU box)
  python scripts/k3_pipe_probe.py"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nearl_b200
from nearl_b200 import _lib
L = _lib.lib()
stream = torch.cuda.current_stream()
_lib.check(L.nrl_set_stream(ctypes.c_void_p(stream.cuda_stream), 1))

def timed(fn, reps=6):
  fn(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  best = 1e9
  for _ in range(3):
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / reps * 1e-3)
  return best

res = []
for F, D in ((100, 128), (200, 96), (300, 64), (512, 64), (100, 30), (65, 34)):
  G = D ** 3
  g = torch.Generator(device="cuda").manual_seed(F + D)
  s = torch.rand((F, G), device="cuda", generator=g) * 3.0
  s[:, ::7] = 1.5    # constant series (entropy range 0)
  nbytes = 4 * F * G + 4 * G
  for agg, name in ((2, "std"), (4, "var"), (7, "entropy"), (3, "median")):
    if agg == 3 and D > 64: continue
    out_ref = torch.empty(G, device="cuda"); out = torch.empty(G, device="cuda")
    def run(o): _lib.check(L.nrl_aggregate_dev(ctypes.c_void_p(s.data_ptr()), F, G, agg, ctypes.c_void_p(o.data_ptr()), ctypes.c_void_p(stream.cuda_stream)))
    _lib.set_option("agg_pipe_stages", -1)
    t_ref = timed(lambda: run(out_ref))
    row = {"F": F, "G": f"{D}^3", "agg": name, "one_shot_GBs": nbytes / t_ref / 1e9}
    for W, S in ((0, 0), (256, 2), (256, 1), (128, 4), (128, 3), (128, 2), (64, 8), (64, 4), (32, 8)):
      _lib.set_option("agg_pipe_w", W); _lib.set_option("agg_pipe_stages", S)
      out.zero_()
      try:
        t = timed(lambda: run(out))
      except Exception as e:
        row[f"W{W}S{S}"] = "n/a"; continue
      same = bool(torch.equal(out.view(torch.int32), out_ref.view(torch.int32)))
      row[f"W{W}S{S}"] = round(nbytes / t / 1e9) if same else "MISMATCH"
    _lib.set_option("agg_pipe_w", 0); _lib.set_option("agg_pipe_stages", 0)
    if agg == 7:
      _lib.set_option("agg_entropy_legacy", 1)
      out.zero_()
      t = timed(lambda: run(out))
      _lib.set_option("agg_entropy_legacy", 0)
      row["legacy_histogram_GBs"] = round(nbytes / t / 1e9)
      out2 = torch.empty_like(out); run(out2); torch.cuda.synchronize()
      row["legacy_histogram_same_bits"] = bool(torch.equal(out.view(torch.int32), out2.view(torch.int32)))
    res.append(row); print(json.dumps(row), flush=True)
