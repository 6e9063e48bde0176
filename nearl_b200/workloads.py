"""The synthetic workloads of BASELINE.json `configs`, shared by bench.py,
the smoke test and the parity tests.  Everything is seeded; nothing is read from disk."""
import numpy as np

from . import features, featurizer, io

BASE_SEED = 20240713


def readme_mass(n_traj=1, outfile=None, seed_offset=0, pin=True):
  """configs[0]: README quick start -- Mass, 32^3, lengths 16, window 10, sigma 1.5, cutoff 5.0,
  focus ':ARG', 5 000 atoms x 100 frames."""
  trajs = [io.synthetic_trajectory(5000, 100, seed=BASE_SEED + 1 + 1000 * (i + seed_offset), pin=pin) for i in range(n_traj)]
  fz = featurizer.Featurizer({"dimensions": 32, "lengths": 16, "time_window": 10})
  fz.register_feature(features.Mass(outkey="mass", outfile=outfile, sigma=1.5, cutoff=5.0))
  fz.register_focus([":ARG"], "mask")
  fz.register_trajloader(trajs)
  return fz, trajs


def dynamic_observers(n_traj=1, n_atoms=20000, n_frames=100, seed_offset=0, outfile=None, pin=True):
  """configs[1]: existence / distinct_count(atomic_number) / mean_distance(mass) observers and a mass
  density flow, mean over a window of 10, 32^3 at 0.5 A, cutoff 5.0, sigma 1.5, focus ':ARG',
  20 000 atoms x 100 frames."""
  trajs = [io.synthetic_trajectory(n_atoms, n_frames, seed=BASE_SEED + 2 + 1000 * (i + seed_offset), pin=pin) for i in range(n_traj)]
  fz = featurizer.Featurizer({"dimensions": 32, "lengths": 16, "time_window": 10, "cutoff": 5.0, "sigma": 1.5})
  fz.register_features([
    features.MarchingObservers(obs="existence", agg="mean", weight_type="mass", outkey="obs_existence", outfile=outfile),
    features.MarchingObservers(obs="distinct_count", agg="mean", weight_type="atomic_number", outkey="obs_distinct", outfile=outfile),
    features.MarchingObservers(obs="mean_distance", agg="mean", weight_type="mass", outkey="obs_mean_distance", outfile=outfile),
    features.DensityFlow(agg="mean", weight_type="mass", outkey="pdf_mass", outfile=outfile),
  ])
  fz.register_focus([":ARG"], "mask")
  fz.register_trajloader(trajs)
  return fz, trajs


def pdbbind_static(n_complex=16, outfile=None, seed_offset=0, pin=True):
  """configs[2]: 8 static channels at 24^3 over a 20 A box, cutoff 2.5, sigma 1.5, focus ':LIG'."""
  trajs = [io.synthetic_complex(3000, seed=BASE_SEED + 3 + i + 100000 * seed_offset, pin=pin) for i in range(n_complex)]
  fz = featurizer.Featurizer({"dimensions": 24, "lengths": 20, "time_window": 1, "cutoff": 2.5, "sigma": 1.5})
  chans = [features.AtomType(focus_element=z, outkey=f"type_{z}", outfile=outfile) for z in (1, 6, 7, 8, 16)]
  chans += [features.HeavyAtom(outkey="heavy", outfile=outfile), features.Mass(outkey="mass", outfile=outfile),
            features.Selection(selection=":LIG", outkey="ligand", outfile=outfile)]
  fz.register_features(chans)
  fz.register_focus([":LIG"], "mask")
  fz.register_trajloader(trajs)
  return fz, trajs


def describe(fz):
  """Feature descriptors for oracle/host_ref.py from a featuriser whose features are cached."""
  descs = []
  for f in fz.FEATURESPACE:
    d = {"outkey": f.outkey, "dims": f.dims, "spacing": f.spacing, "cutoff": f.cutoff, "sigma": f.sigma,
         "padding": f.padding, "byres": f.byres, "selection": (f.selected if f.selection is not None else None),
         "weights": f.cached_array, "frame_offset": f.frame_offset}
    if isinstance(f, features.MarchingObservers):
      d.update(kind="marching_observers", obs=f.obs, agg=f.agg)
    elif isinstance(f, features.DensityFlow):
      d.update(kind="density_flow", agg=f.agg)
    else:
      d.update(kind="static")
    descs.append(d)
  return descs
