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


def misato_mixed(n_traj=1, n_atoms=6000, n_frames=100, seed_offset=0, outfile=None, pin=True):
  """configs[3]: MISATO-style batch -- 3 static + 2 marching-observer + 2 density-flow channels at
  32^3 over 20 A (spacing 0.625), window 10, cutoff 2.5 (static) / 5.0 (dynamic), sigma 1.5,
  focus ':LIG'-free protein systems -> ':ARG', 6 000 atoms x 100 frames per trajectory
  (/root/reference/scripts/slurm/featurize_misato.slurm:54-74)."""
  trajs = [io.synthetic_trajectory(n_atoms, n_frames, seed=BASE_SEED + 4 + 1000 * (i + seed_offset), pin=pin) for i in range(n_traj)]
  fz = featurizer.Featurizer({"dimensions": 32, "lengths": 20, "time_window": 10, "sigma": 1.5})
  fz.register_features([
    features.Mass(outkey="mass", cutoff=2.5, outfile=outfile),
    features.HeavyAtom(outkey="heavy", cutoff=2.5, outfile=outfile),
    features.Electronegativity(outkey="elecneg", cutoff=2.5, outfile=outfile),
    features.MarchingObservers(obs="existence", agg="mean", weight_type="mass", cutoff=5.0, outkey="obs_existence", outfile=outfile),
    features.MarchingObservers(obs="mean_distance", agg="standard_deviation", weight_type="mass", cutoff=5.0, outkey="obs_mean_distance", outfile=outfile),
    features.DensityFlow(agg="mean", weight_type="mass", cutoff=5.0, outkey="pdf_mass", outfile=outfile),
    features.DensityFlow(agg="standard_deviation", weight_type="electronegativity", cutoff=5.0, outkey="pdf_elecneg", outfile=outfile),
  ])
  fz.register_focus([":ARG"], "mask")
  fz.register_trajloader(trajs)
  return fz, trajs


def stress_large(n_traj=1, n_atoms=200000, n_frames=200, window=100, dim=128, seed_offset=0, outfile=None, pin=True,
                 native_atoms=False):
  """configs[4]: 128^3 grid at 0.5 A over a 200 000-atom system, window 100, cutoff 5.0; observers
  existence / distinct_count(atomic_number) / mean_distance(mass) and a mass density flow, mean and
  standard deviation over the window.  `native_atoms` lifts the reference's 9 999-atom cap on
  dynamic crops (/root/reference/nearl/features.py:1279,1353-1359); the default keeps it."""
  trajs = [io.synthetic_trajectory(n_atoms, n_frames, seed=BASE_SEED + 5 + 1000 * (i + seed_offset), pin=pin) for i in range(n_traj)]
  fz = featurizer.Featurizer({"dimensions": dim, "lengths": dim * 0.5, "time_window": window, "cutoff": 5.0, "sigma": 1.5})
  feats = [
    features.MarchingObservers(obs="existence", agg="mean", weight_type="mass", outkey="obs_existence", outfile=outfile),
    features.MarchingObservers(obs="distinct_count", agg="mean", weight_type="atomic_number", outkey="obs_distinct", outfile=outfile),
    features.MarchingObservers(obs="mean_distance", agg="standard_deviation", weight_type="mass", outkey="obs_mean_distance", outfile=outfile),
    features.DensityFlow(agg="mean", weight_type="mass", outkey="pdf_mass", outfile=outfile),
    features.DensityFlow(agg="standard_deviation", weight_type="mass", outkey="pdf_mass_std", outfile=outfile),
  ]
  if native_atoms:
    for f in feats:
      f.MAX_ALLOWED_ATOMS = 0
  fz.register_features(feats)
  fz.register_focus([":ARG"], "mask")
  fz.register_trajloader(trajs)
  return fz, trajs


def residue_classification(n_traj=1, n_residues=300, n_frames=100, seed_offset=0, outfile=None, pin=True):
  """Residue-classification data set (Featurizer.loop_by_residue, /root/reference/scripts/residue_dynamic_generation.py):
  every residue of a 300-residue chain is a sample per window of 10 frames, 2 static + 2 dynamic channels at 32^3 / 16 A."""
  trajs = [io.synthetic_peptide(n_residues, n_frames, seed=BASE_SEED + 6 + 1000 * (i + seed_offset), pin=pin) for i in range(n_traj)]
  fz = featurizer.Featurizer({"dimensions": 32, "lengths": 16, "time_window": 10, "cutoff": 4.0, "sigma": 1.5, "outfile": outfile} if outfile
                             else {"dimensions": 32, "lengths": 16, "time_window": 10, "cutoff": 4.0, "sigma": 1.5})
  fz.register_features([
    features.Mass(outkey="mass", outfile=outfile),
    features.HeavyAtom(outkey="heavy", outfile=outfile),
    features.MarchingObservers(obs="existence", agg="mean", weight_type="mass", outkey="obs_existence", outfile=outfile),
    features.DensityFlow(agg="standard_deviation", weight_type="mass", outkey="pdf_mass", outfile=outfile),
  ])
  fz.register_trajloader(trajs)
  return fz, trajs


def misato_production(n_traj=1, n_atoms=6000, n_frames=100, seed_offset=0, outfile=None, pin=True, label_table=None):
  """The feature set the reference's MISATO production script registers (/root/reference/scripts/featurize_misato.py:107-217):
  10 static atom-type channels (H/C/N/O/S x protein / ligand selection), 2 distinct_count observers and 4 density flows with
  uniformed weights on the protein / ligand selections (mean and standard deviation), and -- with `label_table`, a csv with
  `pdbcode` and `pK` columns holding the trajectories' identities -- the three labels; 32^3 over 20 A, window 10, focus on the
  ligand, 6 000 atoms x 100 frames with a 30-atom ligand."""
  trajs = [io.synthetic_trajectory(n_atoms, n_frames, seed=BASE_SEED + 7 + 1000 * (i + seed_offset), pin=pin, ligand_atoms=30,
                                   identity=f"{(i + seed_offset) % 9 + 1}syn")
           for i in range(n_traj)]
  kw = dict(cutoff=2.5, sigma=1.5, outfile=outfile)
  fz = featurizer.Featurizer({"dimensions": 32, "lengths": 20, "time_window": 10})
  feats = []
  for sel, tag in (("!:LIG", "prot"), (":LIG", "lig")):
    feats += [features.AtomType(selection=sel, focus_element=z, outkey=f"static_{sym}_{tag}", **kw)
              for z, sym in ((1, "H"), (6, "C"), (7, "N"), (8, "O"), (16, "S"))]
  feats += [features.MarchingObservers(selection=sel, weight_type="uniformed", obs="distinct_count", agg="standard_deviation",
                                       outkey=f"mo_PCount_{tag}", **kw) for sel, tag in ((":LIG", "lig"), ("!:LIG", "prot"))]
  feats += [features.DensityFlow(selection=sel, weight_type="uniformed", agg=agg, outkey=f"pdf_P{a}_{tag}", **kw)
            for sel, tag in ((":LIG", "lig"), ("!:LIG", "prot")) for agg, a in (("mean", "M"), ("standard_deviation", "S"))]
  if label_table is not None:
    feats += [features.LabelAffinity(baseline_map=label_table, outkey="pk_original", outfile=outfile),
              features.LabelStepping(baseline_map=label_table, outkey="pk_stepping", outfile=outfile),
              features.LabelPCDT(selection=":LIG", baseline_map=label_table, outkey="label_pcdt", outfile=outfile)]
  fz.register_features(feats)
  fz.register_focus([":LIG"], "mask")
  fz.register_trajloader(trajs)
  return fz, trajs
