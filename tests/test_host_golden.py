"""The Python host half against the REFERENCE's own Python, CPU only.

tests/golden/host_vectors.npz was recorded by tests/golden/make_host_golden.py, which runs the reference's
unmodified nearl/featurizer.py + nearl/features.py + nearl/commands.py in the build container and keeps what
reaches the FFI boundary (`all_actions.*`): focal points, task order, and the coordinate / weight arrays, grid
dimensions and scalars of every call.  Here the same seeded inputs go through

  * oracle/host_ref.py  (the numpy restatement every GPU parity test of the batched path is checked against), and
  * nearl_b200's own host protocol (Featurizer.parse_focus, Feature.cache / query / run, commands.*),

and both must hand over identical bits.  Nothing here reads /root/reference."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import make_host_golden as G                      # noqa: E402  (inputs, feature list, digest -- no reference import at module level)
import nearl_b200                                  # noqa: E402
from nearl_b200 import commands, features, featurizer, workloads   # noqa: E402
from oracle import host_ref                        # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "host_vectors.npz"))


class BoundaryRecorder(G.Recorder):
  """Same recorder the generator put behind the reference's commands.py, here behind nearl_b200/commands.py."""
  def frame_voxelize(self, coords, weights, grid_dims, spacing, cutoff, sigma, auto_translate=0):
    return super().frame_voxelize(coords, weights, grid_dims, spacing, cutoff, sigma, auto_translate)


class OracleSideRecorder:
  """`kernels` back-end of oracle/host_ref.run_task: converts like host_ref.ModuleKernels, records, returns zeros."""
  def __init__(self):
    self.calls = []
  def frame_voxelize(self, coords, weights, dims, spacing, cutoff, sigma):
    self.calls.append(("frame_voxelize", np.array(coords), np.array(weights), np.array(dims, dtype=int), [float(spacing), float(cutoff), float(sigma), -1, 0]))
    return np.zeros(int(np.prod(dims)), dtype=np.float32)
  def density_flow(self, coords, weights, dims, spacing, cutoff, sigma, agg):
    self.calls.append(("density_flow", np.array(coords), np.array(weights), np.array(dims, dtype=int), [float(spacing), float(cutoff), float(sigma), -1, int(agg)]))
    return np.zeros(int(np.prod(dims)), dtype=np.float32)
  def marching_observer(self, coords, weights, dims, spacing, cutoff, obs, agg):
    self.calls.append(("marching_observer", np.array(coords), np.array(weights), np.array(dims, dtype=int), [spacing, cutoff, -1, int(obs), int(agg)]))
    return np.zeros(int(np.prod(dims)), dtype=np.float32)


def build(name, kw):
  traj = G.make_system(kw)
  traj.identity = name
  fz = featurizer.Featurizer(dict(G.FEATURIZER_PARMS))
  fz.register_features(G.feature_set(features))
  fz.register_focus(G.FOCUS, "mask")
  fz.register_trajloader([traj])
  fz.traj = traj
  assert fz.parse_focus()
  for f in fz.FEATURESPACE:
    f.cache(traj)
  return fz, traj


def compare(name, calls, what):
  n = int(GOLD[f"{name}/n_calls"][0])
  assert len(calls) == n, (what, len(calls), n)
  assert [c[0] for c in calls] == [str(k) for k in GOLD[f"{name}/kinds"]], what
  for i, (kind, coords, weights, dims, scalars) in enumerate(calls):
    where = (what, name, i, kind)
    assert np.array_equal(np.asarray(dims, dtype=np.int64), GOLD[f"{name}/dims"][i]), where
    assert np.array_equal(np.asarray(scalars, dtype=np.float64), GOLD[f"{name}/scalars"][i]), (where, scalars, GOLD[f"{name}/scalars"][i])
    assert str(coords.dtype) == str(GOLD[f"{name}/coords_dtype"][i]) and str(weights.dtype) == str(GOLD[f"{name}/weights_dtype"][i]), where
    shape = list(coords.shape) + [0] * (3 - coords.ndim)
    assert shape == list(GOLD[f"{name}/coords_shape"][i]), (where, shape, GOLD[f"{name}/coords_shape"][i])
    if f"{name}/{i:03d}/coords" in GOLD.files:              # full arrays kept for every 16th call: readable diagnostics first
      assert np.array_equal(coords, GOLD[f"{name}/{i:03d}/coords"]), where
      assert np.array_equal(weights, GOLD[f"{name}/{i:03d}/weights"]), where
    assert np.array_equal(G.digest(coords), GOLD[f"{name}/coords_digest"][i]), (where, float(np.sum(coords, dtype=np.float64)), float(GOLD[f"{name}/coords_sum"][i]))
    assert np.array_equal(G.digest(weights), GOLD[f"{name}/weights_digest"][i]), (where, float(np.sum(weights, dtype=np.float64)), float(GOLD[f"{name}/weights_sum"][i]))


@pytest.mark.parametrize("name,kw", G.systems())
def test_focal_points_have_the_reference_bits(name, kw):
  fz, traj = build(name, kw)
  gold = GOLD[f"{name}/focal"]
  assert fz.FOCALPOINTS.dtype == np.float32 and fz.FOCALPOINTS.shape == gold.shape
  assert np.array_equal(fz.FOCALPOINTS.view(np.int32), gold.view(np.int32))
  sel = [traj.top.select(m) for m in G.FOCUS]
  again = host_ref.focal_points_from_selections(traj.xyz, sel, fz.time_window)
  assert np.array_equal(again.view(np.int32), gold.view(np.int32))


@pytest.mark.parametrize("name,kw", G.systems())
def test_oracle_host_half_hands_over_the_reference_bits(name, kw):
  fz, traj = build(name, kw)
  rec = OracleSideRecorder()
  resids = np.array([a.resid for a in traj.top.atoms], dtype=int)
  host_ref.featurize(traj.xyz, resids, workloads.describe(fz), GOLD[f"{name}/focal"], fz.time_window, kernels=rec)
  compare(name, rec.calls, "oracle/host_ref.py")


@pytest.mark.parametrize("name,kw", G.systems())
def test_host_protocol_hands_over_the_reference_bits(name, kw, monkeypatch):
  fz, traj = build(name, kw)
  rec = BoundaryRecorder()
  monkeypatch.setattr(commands, "all_actions", rec)
  order = []
  for bid in range(fz.SLICENUMBER):                          # the reference's loop, /root/reference/nearl/featurizer.py:429-446
    frames = traj.xyz[fz.FRAMESLICES[bid]]
    for pid in range(fz.FOCALNUMBER):
      for feat in fz.FEATURESPACE:
        queried = feat.query(traj.top, frames, fz.FOCALPOINTS[bid, pid])
        feat.run(*queried)
        order.append(feat.outkey)
  assert order == [str(k) for k in GOLD[f"{name}/dump_order"]]
  compare(name, rec.calls, "nearl_b200 host protocol")


@pytest.mark.parametrize("name,restype,tag_limit", G.RESIDUE_LOOPS)
def test_residue_loops_hand_over_the_reference_bits(name, restype, tag_limit):
  # /root/reference/nearl/featurizer.py:471-554 + utils.py:573-659, recorded with the reference's own code: block order,
  # per-block re-caching rule (features.py:452-457), focal points (block centre in the window's first frame), labels
  pep = G.make_peptide()
  fz = featurizer.Featurizer({k: v for k, v in G.RESIDUE_PARMS.items() if k != "outfile"})
  fz.register_features(G.residue_feature_set(features))
  fz.register_trajloader([pep])
  fz.traj = pep
  for f in fz.FEATURESPACE:
    f.cache(pep)
  blocks = fz._residue_blocks(restype, tag_limit=tag_limit)
  rec = OracleSideRecorder()
  resids = np.array([a.resid for a in pep.top.atoms], dtype=int)
  _, labels = host_ref.featurize_by_residue(pep.xyz, resids, workloads.describe(fz), blocks, fz.time_window, kernels=rec)
  compare(name, rec.calls, "oracle/host_ref.py residue loop")
  assert np.array_equal(labels, GOLD[f"{name}/labels"])


@pytest.mark.parametrize("name,ds_kw,mb_kw,seed", G.DATASET_CASES)
def test_dataset_consumer_yields_the_reference_batches(name, ds_kw, mb_kw, seed, tmp_path):
  # /root/reference/nearl/io/dataset.py recorded with the reference's own code on the same files: sample reads, the
  # three normalisations, split_array, shuffle order, and the augmentation's use of numpy's and torch's generators
  import torch
  from nearl_b200 import dataset
  files = G.write_dataset_files(str(tmp_path))
  rec = G.run_dataset_case(dataset, torch, files, ds_kw, mb_kw, seed)
  for key, val in rec.items():
    assert np.array_equal(val, GOLD[f"{name}/{key}"]), (name, key, val, GOLD[f"{name}/{key}"])


def test_result_sink_writes_what_the_reference_sink_writes(tmp_path, monkeypatch):
  # The reference's Featurizer.__init__ (parameter group) + Feature.dump + utils.append_hdf_data
  # (/root/reference/nearl/featurizer.py:159-162, features.py:510-534, utils.py:491-570) were run on this repo's HDF5
  # writer standing in for h5py; the same run through nearl_b200 must leave the same datasets: names (nested keys
  # included), shapes, dtypes, chunking, extensibility, NaN scrub, append order, parameter group.
  path = str(tmp_path / "sink.h5")
  traj = G.make_system(G.SINK_SYSTEM)
  traj.identity = "sink"
  monkeypatch.setattr(commands, "all_actions", G.FakeKernels())
  fz = featurizer.Featurizer(dict(G.SINK_PARMS, outfile=path))
  fz.register_features(G.sink_feature_set(features, path))
  fz.register_focus([":ARG", "@CA"], "mask")
  for tid in range(2):
    fz.traj = traj
    assert fz.parse_focus()
    for f in fz.FEATURESPACE:
      f.cache(traj)
    results = []
    for bid in range(fz.SLICENUMBER):
      frames = traj.xyz[fz.FRAMESLICES[bid]]
      for pid in range(fz.FOCALNUMBER):
        for feat in fz.FEATURESPACE:
          results.append((feat, feat.run(*feat.query(traj.top, frames, fz.FOCALPOINTS[bid, pid]))))
    for feat, res in results:
      feat.dump(res)
  rows = G.file_manifest(path)
  assert sorted(rows) == [str(k) for k in GOLD["sink/keys"]]
  for key, (shape, dtype, chunks, maxshape, compression, dig, val) in rows.items():
    meta = [str(m) for m in GOLD[f"sink/{key}/meta"]]
    if key == "featurizer_parms/outfile":               # the path itself: a string scalar of another length
      assert dtype.startswith("<U") and meta[1].startswith("<U") and repr(shape) == meta[0]
      continue
    assert [repr(shape), dtype, repr(chunks), repr(maxshape), repr(compression)] == meta, (key, meta)
    if dig is not None:
      assert np.array_equal(dig, GOLD[f"sink/{key}/digest"]), key
    else:
      assert [str(v) for v in np.asarray(val).reshape(-1)] == [str(t) for t in GOLD[f"sink/{key}/text"]], key


@pytest.mark.parametrize("ident,kw", G.LABEL_SYSTEMS)
def test_label_features_give_the_reference_values(ident, kw, tmp_path):
  # LabelAffinity / LabelStepping / LabelPCDT (features.py:1559-1775, utils.compute_pcdt, utils.get_pdbcode incl. a superseded
  # PDB code and a duplicated table row) recorded from the reference; here through Featurizer.run() itself, which needs no
  # GPU when every registered feature is a label
  table = G.write_label_table(str(tmp_path))
  traj = G.make_system(kw)
  traj.identity = ident
  got = []
  fz = featurizer.Featurizer(dict(G.LABEL_PARMS))
  fz.register_features(G.label_feature_set(features, table))
  fz.register_focus([":ARG", "@CA"], "mask")
  fz.register_trajloader([traj])
  fz.result_sink = lambda outkey, tid, bid, pid, result: got.append((outkey, float(result)))
  fz.run()
  assert [k for k, _ in got] == [str(k) for k in GOLD[f"labels_{ident}/keys"]]
  assert np.allclose([v for _, v in got], GOLD[f"labels_{ident}/values"], rtol=1e-12, atol=0)
