"""Experiment: where does a SHORT Featurizer.run() (20 trajectories) spend its time -- pipeline fill, steady state, drain?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import nearl_b200
from nearl_b200 import _lib, workloads, featurizer
nearl_b200.update_config(quiet=True)
fz, trajs = workloads.dynamic_observers(n_traj=8)
fz.result_sink = lambda outkey, tid, bid, pid, grid: None
K = 20
idx = [i % 8 for i in range(K)]
fz.run(trajectory_indices=idx); fz.run(trajectory_indices=idx)
ev = []
real_exec = featurizer.LoweredBatch.execute
real_finish = fz._finish
def ex(self, *a, **k):
  t = time.perf_counter(); r = real_exec(self, *a, **k); ev.append(("exec", t, time.perf_counter())); return r
def fin(job, stats):
  t = time.perf_counter(); r = real_finish(job, stats); ev.append(("finish", t, time.perf_counter())); return r
featurizer.LoweredBatch.execute = ex
fz._finish = fin
for rep in range(3):
  ev.clear()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  fz.run(trajectory_indices=idx)
  t1 = time.perf_counter()
  ex_t = [(a - t0, b - t0) for k, a, b in ev if k == "exec"]
  fi_t = [(a - t0, b - t0) for k, a, b in ev if k == "finish"]
  print(f"total {1e3 * (t1 - t0):.3f} ms; exec starts (ms): " + " ".join(f"{1e3 * a:.2f}" for a, _ in ex_t))
  print("   finish ends (ms): " + " ".join(f"{1e3 * b:.2f}" for _, b in fi_t))
