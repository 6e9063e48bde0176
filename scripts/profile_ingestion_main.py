"""cProfile of the calling thread of a file-backed run (see profile_ingestion.py): what the main thread does while the
loader threads read ahead.  python scripts/profile_ingestion_main.py [n] [loader threads]"""
import cProfile, os, pstats, runpy, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["NRL_INGEST_KEEP"] = "1"
n = sys.argv[1] if len(sys.argv) > 1 else "24"
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sys.argv = [sys.argv[0], n]
ns = runpy.run_path(os.path.join(ROOT, "scripts", "profile_ingestion.py"))
from nearl_b200 import featurizer, features, io
loader = io.TrajectoryLoader([(c, ns["root"]) for c in ns["codes"]], trajtype=io.MisatoTraj, superpose=True)
fz = featurizer.Featurizer({"dimensions": 32, "lengths": 16, "time_window": 10, "cutoff": 5.0, "sigma": 1.5, "outfile": os.path.join(ns["root"], "prof.h5")})
fz.register_features([features.MarchingObservers(obs="existence", agg="mean", outkey="e"), features.MarchingObservers(obs="distinct_count", agg="mean", outkey="d"),
                      features.MarchingObservers(obs="mean_distance", agg="mean", outkey="m"), features.DensityFlow(agg="mean", outkey="f")])
fz.register_focus([":ARG"], "mask")
fz.register_trajloader(loader)
fz.prefetch_workers = workers
pr = cProfile.Profile()
pr.enable()
fz.run()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
import shutil
shutil.rmtree(ns["root"], ignore_errors=True)
