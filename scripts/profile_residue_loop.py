import os, sys, time, cProfile, pstats
sys.path.insert(0, "/root/repo")
import numpy as np, nearl_b200
from nearl_b200 import _lib, workloads
nearl_b200.update_config(quiet=True)
fz, trajs = workloads.residue_classification(n_traj=1)
fz.result_sink = lambda *a: None
fz.loop_by_residue("single")
for mode in ("single", "dual"):
  t0 = time.perf_counter(); st = fz.loop_by_residue(mode); dt = time.perf_counter() - t0
  print(mode, "seconds", round(dt, 4), st)
pr = cProfile.Profile(); pr.enable(); fz.loop_by_residue("single"); pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(25)
