"""cProfile of the host side of Featurizer.run (GPU box): python scripts/profile_host.py [workload] [n]"""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nearl_b200
from nearl_b200 import workloads
nearl_b200.update_config(quiet=True)
name = sys.argv[1] if len(sys.argv) > 1 else "dynamic"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
make = {"dynamic": workloads.dynamic_observers, "misato": workloads.misato_mixed, "readme": workloads.readme_mass, "misato_production": workloads.misato_production}
fz, trajs = workloads.pdbbind_static(n_complex=n) if name == "pdbbind" else make[name](n_traj=n)
fz.result_sink = lambda *a: None
fz.run()
reps = 5
pr = cProfile.Profile()
pr.enable()
for _ in range(reps):
  fz.run()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
