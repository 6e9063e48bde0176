"""Generates tests/golden/host_vectors.npz by running the REFERENCE's own Python host half
(/root/reference/nearl/featurizer.py Featurizer.run / parse_focus, /root/reference/nearl/features.py
Feature.cache / query, DynamicFeature.query, cache_properties) in this container, CPU only:

  python tests/golden/make_host_golden.py

The reference package imports h5py, pytraj, rdkit and openbabel at module level and none of them exists on
this image, so they are replaced by inert stand-ins before `import nearl` (nothing of theirs is on the path
that is recorded here: the topology the reference sees is nearl_b200.io.Topology, whose atoms carry the
attributes the reference reads -- index, resid, atomic_number, name -- and whose `select` answers the masks).
The reference's compiled module `all_actions` is not built for this container; a recorder takes its place
behind the reference's unmodified `nearl.commands.*`, so what is pinned is everything the reference does up to
the FFI boundary: focal points, the order of the (slice, focal point, feature) tasks, and the exact coordinate /
weight arrays, grid dimensions and scalars every task hands to the compiled module.  tests/test_host_golden.py replays the same seeded inputs through
oracle/host_ref.py and through nearl_b200's own host protocol and demands identical bits."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


# ---- shared with the test: seeded inputs and the feature list -------------------------------------------------
def systems():
  """(name, trajectory kwargs) of the seeded synthetic systems (nearl_b200.io)."""
  return [("traj_a", dict(n_atoms=700, n_frames=13, seed=4242)),
          ("traj_b", dict(n_atoms=520, n_frames=8, seed=977))]


FEATURIZER_PARMS = {"dimensions": [12, 14, 16], "lengths": [9.0, 10.5, 12.0], "time_window": 4, "sigma": 1.3, "cutoff": 2.75}
FOCUS = [":ARG", ":LIG,GLY", "@CA"]


def feature_set(F):
  """The same constructor calls against the reference's `nearl.features` and against `nearl_b200.features`."""
  return [
    F.Mass(outkey="mass"),
    F.HeavyAtom(outkey="heavy", cutoff=3.5),
    F.AtomType(focus_element=8, outkey="type_o", sigma=0.9),
    F.Selection(selection=":ARG,LYS", outkey="sel_basic"),
    F.Electronegativity(outkey="elecneg", byres=True),
    F.Mass(outkey="mass_noh", selection="!@H=", padding=1.5),
    F.AtomicNumber(outkey="z_off", frame_offset=2),
    F.MarchingObservers(obs="existence", agg="mean", weight_type="mass", outkey="obs_exist"),
    F.MarchingObservers(obs="distinct_count", agg="max", weight_type="atomic_number", outkey="obs_distinct", cutoff=4.0),
    F.MarchingObservers(obs="mean_distance", agg="standard_deviation", weight_type="electronegativity", selection="!@H=", outkey="obs_md"),
    F.MarchingObservers(obs="density", agg="median", weight_type="heavy_atom", byres=True, outkey="obs_dens"),
    F.DensityFlow(agg="mean", weight_type="mass", outkey="flow_mass"),
    F.DensityFlow(agg="drift", weight_type="hydrophobicity", outkey="flow_hyd", sigma=2.0, cutoff=5.0),
    F.DensityFlow(agg="information_entropy", weight_type="uniformed", manual_weight=2.5, outkey="flow_uni"),
    F.DensityFlow(agg="min", weight_type="atom_type", element_type=7, outkey="flow_n"),
  ]


RESIDUE_PARMS = {"dimensions": [10, 10, 10], "lengths": [8.0, 8.0, 8.0], "time_window": 4, "sigma": 1.2, "cutoff": 3.0, "outfile": "/tmp/nearl_host_golden_unused.h5", "hdf_compress_level": 0}
RESIDUE_LOOPS = [("pep_single", "single", 2), ("pep_single_all", "single", 0), ("pep_dual", "dual", 0)]
PEPTIDE = dict(n_residues=26, n_frames=9, seed=31337, break_every=7)


def residue_feature_set(F):
  return [
    F.Mass(outkey="mass"),
    F.HeavyAtom(outkey="heavy", selection="!@H="),
    F.MarchingObservers(obs="existence", agg="mean", weight_type="mass", outkey="obs_exist"),
    F.DensityFlow(agg="standard_deviation", weight_type="electronegativity", outkey="flow_en"),
  ]


def make_peptide():
  from nearl_b200 import io as bio
  return bio.synthetic_peptide(PEPTIDE["n_residues"], PEPTIDE["n_frames"], seed=PEPTIDE["seed"], pin=False, break_every=PEPTIDE["break_every"])


# ---- result sink (Feature.dump -> utils.append_hdf_data, features.py:510-534, utils.py:491-570) --------------------------
SINK_PARMS = {"dimensions": [8, 9, 10], "lengths": [6.0, 6.75, 7.5], "time_window": 3, "sigma": 1.1, "cutoff": 2.5}
SINK_SYSTEM = dict(n_atoms=300, n_frames=7, seed=555)


def sink_feature_set(F, outfile):
  return [
    F.Mass(outkey="mass", outfile=outfile),
    F.HeavyAtom(outkey="heavy", outfile=outfile, hdf_compress_level=4),
    F.MarchingObservers(obs="existence", agg="mean", weight_type="mass", outkey="obs/exist", outfile=outfile),
    F.DensityFlow(agg="mean", weight_type="mass", outkey="flow", outfile=outfile),
  ]


def file_manifest(path):
  """{dataset path: (shape, dtype, chunks, maxshape, compression, digest)} of an HDF5 file read through hdf5_lite."""
  from nearl_b200 import hdf5_lite
  rows = {}
  def walk(group, prefix):
    for key in sorted(group.keys()):
      node = group[key]
      if hasattr(node, "keys"):
        walk(node, prefix + key + "/")
      else:
        val = np.asarray(node[()] if node.shape == () else node[:])
        rows[prefix + key] = (tuple(node.shape), str(val.dtype), node.chunks, tuple(node.maxshape), node.compression, digest(val) if val.dtype.kind in "fiub" else None, val if val.dtype.kind not in "fiub" else None)
  with hdf5_lite.File(path, "r") as hdf:
    walk(hdf, "")
  return rows


def pack_manifest(prefix, rows, out):
  out[f"{prefix}/keys"] = np.array(sorted(rows))
  for key, (shape, dtype, chunks, maxshape, compression, dig, val) in rows.items():
    out[f"{prefix}/{key}/meta"] = np.array([repr(shape), dtype, repr(chunks), repr(maxshape), repr(compression)])
    if dig is not None:
      out[f"{prefix}/{key}/digest"] = dig
    else:
      out[f"{prefix}/{key}/text"] = np.array([str(v) for v in np.asarray(val).reshape(-1)])


# ---- label features (features.py:1559-1775, utils.py:70-166,196-214) ---------------------------------------------------
LABEL_TABLE = "pdbcode,pK,resolution\n1abc,7.25,2.0\n6qs5,4.5,1.9\n2xyz,3.1,1.5\n1abc,9.99,3.0\n"
LABEL_SYSTEMS = [("1ABC", dict(n_atoms=700, n_frames=13, seed=4242)), ("4dgo", dict(n_atoms=520, n_frames=8, seed=977))]
LABEL_PARMS = {"dimensions": [8, 8, 8], "lengths": [8.0, 8.0, 8.0], "time_window": 4, "sigma": 1.0, "cutoff": 2.0}


def write_label_table(folder):
  path = os.path.join(folder, "nearl_golden_labels.csv")
  with open(path, "w") as f:
    f.write(LABEL_TABLE)
  return path


def label_feature_set(F, table):
  return [
    F.LabelAffinity(baseline_map=table, outkey="pk"),
    F.LabelStepping(baseline_map=table, outkey="step"),
    F.LabelAffinity(baseline_map=table, colname="resolution", outkey="res"),
    F.LabelPCDT(selection=":LYS&!@H=", baseline_map=table, outkey="pcdt"),
    F.LabelPCDT(selection=":ARG@CA", search_cutoff=5.5, baseline_map=table, outkey="pcdt_cut"),
  ]


# ---- dataset consumer (nearl/io/dataset.py) ----------------------------------------------------------------
DATASET_KEYS = ["mass", "obs"]
DATASET_CASES = [       # (name, Dataset kwargs, mini_batches kwargs, numpy / torch seed)
  ("ds_plain", dict(normalize=True), dict(batch_size=5, shuffle=False, process_nr=2), 11),
  ("ds_minmax_shuffle", dict(normalize="minmax"), dict(batch_size=4, shuffle=True, process_nr=2), 12),
  ("ds_zscore_nr", dict(normalize="zscore"), dict(batch_nr=3, shuffle=True, process_nr=2), 13),
  ("ds_augment", dict(normalize=False), dict(batch_size=6, shuffle=True, process_nr=2, augment=True, augment_translation=2), 14),
  ("ds_griddim", dict(normalize=False, grid_dim=6), dict(batch_size=7, shuffle=False, process_nr=2, augment=True, augment_translation=1), 15),
]


def write_dataset_files(folder):
  """Two small HDF5 files in the featuriser's layout (written through nearl_b200.utils, i.e. hdf5_lite here)."""
  from nearl_b200 import utils as butils
  rng = np.random.default_rng(2718)
  files = []
  for fi, n in enumerate((9, 8)):
    path = os.path.join(folder, f"nearl_golden_ds{fi}.h5")
    if os.path.exists(path):
      os.remove(path)
    butils.dump_dict(path, "featurizer_parms", {"dimensions": np.array([6, 6, 6]), "time_window": 4})
    for key in DATASET_KEYS:
      data = rng.random((n, 6, 6, 6)).astype(np.float32) * (fi + 1.5)
      data[1] = 0.75                                          # a constant sample: the normalisations return zeros
      butils.append_hdf_data(path, key, data, dtype=np.float32, maxshape=(None, 6, 6, 6), chunks=(1, 6, 6, 6))
    butils.append_hdf_data(path, "label", rng.integers(0, 20, size=n), dtype=int, maxshape=(None,), chunks=True)
    files.append(path)
  return files


def run_dataset_case(D, torch, files, ds_kw, mb_kw, seed):
  """Everything recorded for one case, from module `D` (the reference's nearl.io.dataset or nearl_b200.dataset)."""
  rec = {}
  ds = D.Dataset(files, feature_keys=list(DATASET_KEYS), **ds_kw)
  rec["len"] = np.array([len(ds)])
  rec["size"] = np.asarray(ds.size, dtype=np.int64)
  items = [ds[i] for i in (0, 1, 8, 9, len(ds) - 1)]
  rec["item_data"] = np.stack([digest(d.numpy()) for d, _ in items])
  rec["item_label"] = np.array([int(l.reshape(-1)[0]) for _, l in items])
  np.random.seed(seed)
  torch.manual_seed(seed)
  digs, labs, shapes = [], [], []
  for data, label in ds.mini_batches(**mb_kw):
    arr = np.ascontiguousarray(data.numpy())
    digs.append(digest(arr)); shapes.append(list(arr.shape)); labs.append(label.numpy().reshape(-1).astype(np.int64))
  rec["batch_digest"] = np.stack(digs)
  rec["batch_shape"] = np.array(shapes, dtype=np.int64)
  rec["batch_labels"] = np.concatenate(labs)
  rec["split"] = np.array([len(b) for n, bs in ((17, 5), (12, 4), (3, 8), (8, 8)) for b in D.split_array(np.arange(n), bs)])
  return rec


def make_system(kw):
  from nearl_b200 import io as bio
  return bio.synthetic_trajectory(kw["n_atoms"], kw["n_frames"], seed=kw["seed"], pin=False)


# ---- stand-ins for the third-party modules the reference imports at module level ---------------------------------
class _Inert:
  def __init__(self, *a, **k): pass
  def __getattr__(self, k): return _Inert()
  def __call__(self, *a, **k): return _Inert()


def _module(name, **attrs):
  m = types.ModuleType(name)
  m.__dict__.update(attrs)
  sys.modules[name] = m
  return m


class PtFrame:
  """pytraj.Frame as far as the label features read it."""
  def __init__(self, xyz=None):
    self.xyz = np.asarray(xyz)


class PtTrajectory:
  """pytraj.Trajectory as far as the recorded path reads it: xyz, top, n_frames, n_atoms, identity, traj[i]."""
  def __init__(self, xyz=None, top=None, identity="stub", **kw):
    self.xyz = np.asarray(xyz)
    self.top = top
    self.identity = identity
  @property
  def n_frames(self): return self.xyz.shape[0]
  @property
  def n_atoms(self): return self.xyz.shape[1]
  def __len__(self): return self.xyz.shape[0]
  def __getitem__(self, i): return PtFrame(np.array(self.xyz[int(i)]))


def install_stubs():
  _module("h5py", File=_Inert, Dataset=_Inert, Group=_Inert)
  _module("pytraj", Trajectory=PtTrajectory, Topology=type("Topology", (), {}), Frame=PtFrame, Atom=_Inert, Residue=_Inert)
  rd = _module("rdkit")
  rd.Chem = _module("rdkit.Chem", AllChem=_Inert(), rdFMCS=_Inert(), rdMolTransforms=_Inert())
  for sub in ("AllChem", "rdFMCS", "rdMolTransforms"):
    _module("rdkit.Chem." + sub)
  ob = _module("openbabel")
  ob.openbabel = _module("openbabel.openbabel")
  ob.pybel = _module("openbabel.pybel")


class Recorder:
  """Stands in for the compiled module `nearl.all_actions` (the drop-in boundary, /root/reference/src/actions_py.cpp:297-351):
  keeps the arguments of every FFI call exactly as nearl/commands.py hands them over, returns zeros of the grid size."""
  def __init__(self):
    self.calls = []
  def _grid(self, dims):
    return np.zeros(int(np.prod([int(v) for v in dims])), dtype=np.float32)
  def frame_voxelize(self, coords, weights, grid_dims, spacing, cutoff, sigma, auto_translate):
    self.calls.append(("frame_voxelize", np.array(coords), np.array(weights), np.array(grid_dims), [spacing, cutoff, sigma, -1, auto_translate]))
    return self._grid(grid_dims)
  def density_flow(self, coords, weights, grid_dims, spacing, cutoff, sigma, type_agg):
    self.calls.append(("density_flow", np.array(coords), np.array(weights), np.array(grid_dims), [spacing, cutoff, sigma, -1, type_agg]))
    return self._grid(grid_dims)
  def marching_observer(self, coords, weights, grid_dims, spacing, cutoff, type_obs, type_agg):
    self.calls.append(("marching_observer", np.array(coords), np.array(weights), np.array(grid_dims), [spacing, cutoff, -1, type_obs, type_agg]))
    return self._grid(grid_dims)
  def frame_observation(self, *a, **k):
    raise AssertionError("not on the recorded path")


class FakeKernels(Recorder):
  """Deterministic stand-in results (a ramp that depends on the call number, with NaNs the sink has to scrub)."""
  def _grid(self, dims):
    n = int(np.prod([int(v) for v in dims]))
    out = (np.arange(n, dtype=np.float32) % 17) * np.float32(0.25) + np.float32(len(self.calls))
    out[len(self.calls) % 5::23] = np.nan
    return out


def digest(a):
  """sha256 of dtype, shape and bytes of a C-contiguous copy -- what 'identical bits' means in the test."""
  import hashlib
  a = np.ascontiguousarray(a)
  h = hashlib.sha256()
  h.update(str(a.dtype).encode() + b"|" + ",".join(str(v) for v in a.shape).encode() + b"|")
  h.update(a.tobytes())
  return np.frombuffer(h.digest(), dtype=np.uint8).copy()


FULL_EVERY = 16       # every 16th call keeps its full arrays (the others their shape, digest and sum), so the file stays small


def pack(prefix, focal, calls, dumps, out):
  """Flatten one system's record into npz entries."""
  out[f"{prefix}/focal"] = np.asarray(focal, dtype=np.float32)
  out[f"{prefix}/n_calls"] = np.array([len(calls)], dtype=np.int64)
  out[f"{prefix}/kinds"] = np.array([c[0] for c in calls])
  out[f"{prefix}/dump_order"] = np.array(dumps)
  out[f"{prefix}/dims"] = np.array([c[3] for c in calls], dtype=np.int64)
  out[f"{prefix}/scalars"] = np.array([c[4] for c in calls], dtype=np.float64)
  out[f"{prefix}/coords_shape"] = np.array([list(c[1].shape) + [0] * (3 - c[1].ndim) for c in calls], dtype=np.int64)
  out[f"{prefix}/coords_dtype"] = np.array([str(c[1].dtype) for c in calls])
  out[f"{prefix}/weights_dtype"] = np.array([str(c[2].dtype) for c in calls])
  out[f"{prefix}/coords_digest"] = np.stack([digest(c[1]) for c in calls])
  out[f"{prefix}/weights_digest"] = np.stack([digest(c[2]) for c in calls])
  out[f"{prefix}/coords_sum"] = np.array([np.sum(c[1], dtype=np.float64) for c in calls])
  out[f"{prefix}/weights_sum"] = np.array([np.sum(c[2], dtype=np.float64) for c in calls])
  for i, (kind, coords, weights, dims, scalars) in enumerate(calls):
    if i % FULL_EVERY == 0:
      out[f"{prefix}/{i:03d}/coords"] = coords
      out[f"{prefix}/{i:03d}/weights"] = weights


def main(path):
  install_stubs()
  sys.path.insert(0, "/root/reference")
  cwd = os.getcwd()
  os.chdir("/tmp")                  # the reference appends to ./nearl.log-style files relative to its temp folder
  import nearl
  import nearl.features as RF
  import nearl.featurizer as RFz
  import nearl.commands as RC
  os.chdir(cwd)
  out = {}
  for name, kw in systems():
    sys_traj = make_system(kw)
    rec = Recorder()
    RC.all_actions = rec          # nearl/commands.py:6-9 -- the import failed (no GPU build of the module here)
    dumps = []
    feats = feature_set(RF)
    for f in feats:
      f.dump = (lambda key: (lambda result: dumps.append(key)))(f.outkey)
    fz = RFz.Featurizer(dict(FEATURIZER_PARMS))
    fz.register_features(feats)
    fz.register_focus(FOCUS, "mask")
    fz.register_trajloader([PtTrajectory(xyz=np.array(sys_traj.xyz), top=sys_traj.top, identity=name)])
    fz.run()
    pack(name, fz.FOCALPOINTS, rec.calls, dumps, out)
    print(name, "recorded", len(rec.calls), "kernel calls,", len(dumps), "dumps; focal points", fz.FOCALPOINTS.shape)
  # ---- the residue loops (featurizer.py:471-554, utils.py:573-659)
  import nearl.utils as RU
  for name, restype, tag_limit in RESIDUE_LOOPS:
    pep = make_peptide()
    rec = Recorder()
    RC.all_actions = rec
    dumps, labels = [], []
    feats = residue_feature_set(RF)
    for f in feats:
      f.dump = (lambda key: (lambda result: dumps.append(key)))(f.outkey)
    RU.append_hdf_data = lambda outfile, key, data, **kw: labels.append(np.array(data))
    RU.dump_dict = lambda *a, **k: None           # featurizer_parms of the run: an HDF5 write, not on the recorded path
    fz = RFz.Featurizer(dict(RESIDUE_PARMS))
    fz.register_features(feats)
    fz.register_trajloader([PtTrajectory(xyz=np.array(pep.xyz), top=pep.top, identity=name)])
    fz.loop_by_residue(restype, tag_limit=tag_limit) if restype == "single" else fz.loop_by_residue(restype)
    pack(name, np.zeros((0, 0, 3), dtype=np.float32), rec.calls, dumps, out)
    out[f"{name}/labels"] = np.concatenate(labels) if labels else np.zeros(0, dtype=int)
    print(name, "recorded", len(rec.calls), "kernel calls,", len(dumps), "dumps,", sum(len(l) for l in labels), "labels")
  # ---- label features: values dumped per (window, focal point) task, in task order
  table = write_label_table("/tmp")
  for ident, kw in LABEL_SYSTEMS:
    ltraj = make_system(kw)
    vals = []
    feats = label_feature_set(RF, table)
    for f in feats:
      f.dump = (lambda key: (lambda result: vals.append((key, float(result)))))(f.outkey)
    fz = RFz.Featurizer(dict(LABEL_PARMS))
    fz.register_features(feats)
    fz.register_focus([":ARG", "@CA"], "mask")
    fz.register_trajloader([PtTrajectory(xyz=np.array(ltraj.xyz), top=ltraj.top, identity=ident)])
    fz.run()
    out[f"labels_{ident}/keys"] = np.array([k for k, _ in vals])
    out[f"labels_{ident}/values"] = np.array([v for _, v in vals], dtype=np.float64)
    print("labels", ident, len(vals), "values", vals[:5])
  # ---- the reference's sink on this repo's self-contained HDF5 reader / writer standing in for h5py: dataset names,
  # ---- shapes, chunking, append semantics, NaN scrub, the parameter group (not the bytes of the file: no libhdf5 here)
  from nearl_b200 import hdf5_lite
  import torch
  sys.modules["h5py"].File = hdf5_lite.File
  RU.h5py = sys.modules["h5py"]
  import importlib
  importlib.reload(RU)                      # drops the recorders put into nearl.utils above
  RF.utils = RU; RFz.utils = RU
  sink_path = "/tmp/nearl_golden_sink_ref.h5"
  if os.path.exists(sink_path):
    os.remove(sink_path)
  sink_traj = make_system(SINK_SYSTEM)
  RC.all_actions = FakeKernels()
  fz = RFz.Featurizer(dict(SINK_PARMS, outfile=sink_path))
  fz.register_features(sink_feature_set(RF, sink_path))
  fz.register_focus([":ARG", "@CA"], "mask")
  fz.register_trajloader([PtTrajectory(xyz=np.array(sink_traj.xyz), top=sink_traj.top, identity="sink")] * 2)
  fz.run()
  rows = file_manifest(sink_path)
  pack_manifest("sink", rows, out)
  print("sink recorded", len(rows), "datasets:", {k: v[0] for k, v in rows.items()})
  # ---- the dataset consumer (nearl/io/dataset.py:40-393)
  import nearl.io.dataset as RD
  RD.h5py = sys.modules["h5py"]
  files = write_dataset_files("/tmp")
  for name, ds_kw, mb_kw, seed in DATASET_CASES:
    rec = run_dataset_case(RD, torch, files, ds_kw, mb_kw, seed)
    for k, v in rec.items():
      out[f"{name}/{k}"] = v
    print(name, "recorded", len(rec["batch_digest"]), "batches")
  np.savez_compressed(path, **out)
  print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
  main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_vectors.npz"))
