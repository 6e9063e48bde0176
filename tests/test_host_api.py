"""CPU-only checks: the C-ABI library loads and exports every declared symbol, fails loudly
without a device, the host-side mirror of the reference API behaves like the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

import nearl_b200
from nearl_b200 import _lib, constants, features, featurizer, io, utils

nearl_b200.update_config(quiet=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
  header = open(os.path.join(ROOT, "include", "nearl_b200.h")).read()
  declared = set(re.findall(r"\b(nrl_[a-z0-9_]+)\s*\(", header))
  assert declared == set(_lib.EXPORTED), declared ^ set(_lib.EXPORTED)
  L = _lib.lib()
  for name in declared:
    assert hasattr(L, name), name
  assert b"sm_100a" in L.nrl_version()


def test_no_cpu_fallback():
  if _lib.device_count() > 0:
    pytest.skip("a CUDA device is present")
  from nearl_b200 import all_actions
  with pytest.raises(RuntimeError, match="no usable CUDA device"):
    all_actions.aggregate(np.zeros((2, 8), dtype=np.float32), 1)
  with pytest.raises(RuntimeError):
    all_actions.frame_voxelize(np.zeros((4, 3), np.float32), np.ones(4, np.float32), [8, 8, 8], 1.0, 2.0, 1.0, 0)


def test_product_never_imports_the_oracle():
  for dirpath, _, files in os.walk(os.path.join(ROOT, "nearl_b200")):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".inc", ".h")):
        text = open(os.path.join(dirpath, f)).read()
        assert "import oracle" not in text and "from oracle" not in text and "nearl_oracle" not in text, f


def test_validation_happens_before_any_device_work():
  from nearl_b200 import all_actions
  with pytest.raises(ValueError, match="3 dimensions"):
    all_actions.marching_observer(np.zeros((5, 3), np.float32), np.ones(5, np.float32), [8, 8, 8], 1.0, 2.0, 1, 1)
  L = _lib.lib()
  dims = (ctypes.c_int * 3)(8, 8, 8)
  buf = (ctypes.c_float * 512)()
  assert L.nrl_marching_observer(buf, buf, 1, 1, dims, 1.0, 2.0, 7, 1, buf) == _lib.ERR_OBSERVABLE
  assert L.nrl_marching_observer(buf, buf, 1, 1, dims, 1.0, 2.0, 1, 0, buf) == _lib.ERR_AGGREGATION
  assert L.nrl_marching_observer(buf, buf, 513, 1, dims, 1.0, 2.0, 1, 1, buf) == _lib.ERR_FRAMES
  assert b"exceeds the maximum number of frames allowed 512" in L.nrl_last_error()
  assert L.nrl_density_flow(buf, buf, 1, 1, dims, 1.0, 2.0, 1.0, 9, buf) == _lib.ERR_AGGREGATION


def test_constants_match_reference_tables():
  ref = "/root/reference/nearl/constants.py"
  assert constants.ATOMICMASS[6] == 12.011 and constants.ELECNEG[8] == 3.44 and len(constants.RES) == 20
  if not os.path.exists(ref):
    pytest.skip("reference tree not present")
  import importlib.util
  spec = importlib.util.spec_from_file_location("refconst", ref)
  m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
  assert constants.ATOMICMASS == m.ATOMICMASS and constants.ELECNEG == m.ELECNEG
  assert constants.COMMON_FEATURE_PARMS == m.COMMON_FEATURE_PARMS and constants.RES == m.RES


def test_amber_mask_subset():
  traj = io.synthetic_trajectory(320, 1, seed=0, pin=False)
  top = traj.top
  arg = top.select(":ARG")
  assert len(arg) == 16 and all(top.resnames[i] == "ARG" for i in arg)
  assert len(top.select("!:ARG")) == 304
  assert len(top.select(":*")) == 320
  assert np.array_equal(top.select(":1-2"), np.arange(32))
  assert set(top.select(":ARG,LYS")) == set(top.select(":ARG")) | set(top.select(":LYS"))
  assert set(top.select("!(:ARG,LYS)")) == set(range(320)) - set(top.select(":ARG,LYS"))
  assert np.array_equal(top.select("@CA"), np.arange(1, 320, 16))
  assert np.array_equal(top.select(":ARG@CA"), np.array([17]))
  assert set(top.select(":ARG|@CA")) == set(arg) | set(range(1, 320, 16))
  assert len(top.select(":XYZ")) == 0
  with pytest.raises(ValueError):
    top.select(":ARG<:5.0")


def test_featurizer_parameters_and_hook():
  fz = featurizer.Featurizer({"dimensions": 32, "lengths": 16, "time_window": 10, "cutoff": 5.0, "sigma": 1.5})
  assert list(fz.dims) == [32, 32, 32] and fz.spacing == 0.5 and list(fz.lengths) == [16.0, 16.0, 16.0]
  f = features.Mass(outkey="mass")
  fz.register_feature(f)
  assert f.cutoff == 5.0 and f.sigma == 1.5 and f.padding == 5.0 and f.frame_offset == 0
  assert list(f.dims) == [32, 32, 32] and f.spacing == 0.5 and f.center.dtype == np.float32 and list(f.center) == [8, 8, 8]
  g = features.Mass(outkey="mass")
  with pytest.raises(ValueError, match="unique"):
    fz.register_feature(g)
  fz2 = featurizer.Featurizer({"dimensions": [24, 24, 24], "spacing": 0.75}, sigma=2.0)
  assert list(fz2.lengths) == [18.0, 18.0, 18.0] and fz2.FEATURE_PARMS["sigma"] == 2.0
  assert fz2.parms["cutoff"] == 0 and fz2.parms["spacing"] == 0.75
  with pytest.raises(KeyError):
    nearl_b200.update_config(nonsense=1)
  mo = features.MarchingObservers(obs="distinct_count", agg="drift", weight_type="residue_id", outkey="x", cutoff=3.0)
  assert (mo.obs, mo.agg, mo.weight_type) == (3, 8, 2)
  with pytest.raises(ValueError):
    features.MarchingObservers(obs="nope", outkey="y").obs


def test_focus_parsing_and_frame_windows():
  traj = io.synthetic_trajectory(640, 25, seed=1, pin=False)
  fz = featurizer.Featurizer({"dimensions": 8, "lengths": 8, "time_window": 10})
  fz.register_focus([":ARG", ":LYS"], "mask")
  fz.traj = traj
  assert fz.SLICENUMBER == 2 and fz.FRAMESLICES[1] == slice(10, 20)
  assert fz.parse_focus() == 1 and fz.FOCALPOINTS.shape == (2, 2, 3) and fz.FOCALPOINTS.dtype == np.float32
  sel = traj.top.select(":LYS")
  assert np.array_equal(fz.FOCALPOINTS[1, 1], np.mean(traj.xyz[10][sel], axis=0).astype(np.float32))
  # the one-gather form gives the bits of the reference's per-frame np.mean(frame[selection], axis=0), f32 and f64
  for dtype in (np.float32, np.float64):
    big = io.synthetic_trajectory(5000, 40, seed=2, dtype=dtype, pin=False)
    fz.traj = big
    assert fz.parse_focus() == 1
    for midx, mask in enumerate((":ARG", ":LYS")):
      sel = big.top.select(mask)
      for w in range(fz.SLICENUMBER):
        assert np.array_equal(fz.FOCALPOINTS[w, midx], np.mean(big.xyz[w * 10][sel], axis=0).astype(np.float32))
  fz.traj = traj
  fz.register_focus([[0, 1, 2]], "index")
  assert fz.parse_focus() == 1 and np.allclose(fz.FOCALPOINTS[0, 0], traj.xyz[0][:3].mean(0))
  fz.register_focus([":NOPE"], "mask")
  assert not fz.parse_focus()


def test_host_query_matches_independent_restatement():
  from oracle import host_ref
  traj = io.synthetic_trajectory(2000, 3, seed=5, pin=False)
  fz = featurizer.Featurizer({"dimensions": 16, "lengths": 8, "time_window": 3, "cutoff": 3.0, "sigma": 1.0})
  feats = [features.Mass(outkey="a"), features.Selection(selection=":ARG", outkey="b", byres=True, padding=1.0),
           features.MarchingObservers(obs="direct_count", outkey="c", weight_type="atomic_id")]
  for f in feats:
    fz.register_feature(f)
    f.cache(traj)
  focal = np.mean(traj.xyz[0][traj.top.select(":GLY")], axis=0).astype(np.float32)
  for f in feats:
    desc = {"dims": f.dims, "spacing": f.spacing, "cutoff": f.cutoff, "padding": f.padding, "byres": f.byres,
            "selection": f.selected if f.selection is not None else None}
    if isinstance(f, features.DynamicFeature):
      coords, weights = f.query(traj.top, traj.xyz, focal)
      assert coords.shape[0] == 3 and coords.dtype == np.float32 and weights.shape == (3 * coords.shape[1],)
      for i in range(3):
        mask, inbox = host_ref.base_query(desc, f.resids, traj.xyz[i], focal)
        n = len(inbox)
        assert np.array_equal(coords[i, :n], inbox.astype(np.float32)) and np.all(coords[i, n:] == 99999.0)
        assert np.array_equal(weights.reshape(3, -1)[i, :n], f.cached_array[mask])
    else:
      coords, weights = f.query(traj.top, traj.xyz, focal)
      mask, inbox = host_ref.base_query(desc, f.resids, traj.xyz[0], focal)
      assert np.array_equal(coords, inbox) and np.array_equal(weights, f.cached_array[mask])


def test_weight_tables():
  traj = io.synthetic_trajectory(160, 1, seed=2, pin=False)
  z = traj.top.atomic_numbers
  def w(f):
    f.cache(traj); return f.cached_array
  assert np.array_equal(w(features.Mass()), np.array([constants.ATOMICMASS[i] for i in z], dtype=np.float32))
  assert np.array_equal(w(features.HeavyAtom(default_weight=3)), np.where(z > 1, 3, 0).astype(np.float32))
  assert np.array_equal(w(features.AtomType(focus_element=8)), (z == 8).astype(np.float32))
  bb = np.array([n in ("C", "O", "CA", "HA", "N", "HN") for n in traj.top.names])
  assert np.array_equal(w(features.Backbone()), bb.astype(np.float32))
  assert np.array_equal(w(features.Backbone(reverse=True)), (~bb).astype(np.float32))
  assert np.array_equal(features.cache_properties(traj, 26), (~bb).astype(np.float32))     # inverted in the reference too
  assert np.array_equal(features.cache_properties(traj, 27), (~bb).astype(np.float32))
  assert np.array_equal(features.cache_properties(traj, 1), np.arange(160, dtype=np.float32))
  assert np.array_equal(features.cache_properties(traj, 16, manual_weight=2.5), np.full(160, 2.5, np.float32))
  en = np.array([constants.ELECNEG[i] for i in z], dtype=np.float32)
  assert np.array_equal(w(features.Hydrophobicity()), np.abs(en - constants.ELECNEG[6]))
  with pytest.raises(NotImplementedError):
    features.cache_properties(traj, 22)


def test_value_ranks_for_distinct_count():
  traj = io.synthetic_trajectory(800, 2, seed=3, pin=False)
  fz = featurizer.Featurizer({"dimensions": 8, "lengths": 8, "time_window": 2, "cutoff": 3.0})
  few = features.MarchingObservers(obs="distinct_count", weight_type="atomic_number", outkey="few")
  many = features.MarchingObservers(obs="distinct_count", weight_type="residue_id", outkey="many")
  fz.register_features([few, many])
  few.cache(traj); many.cache(traj)
  ranks = few.value_ranks()
  w = np.asarray(few.cached_array, dtype=np.float32)
  assert ranks.dtype == np.float32 and ranks.max() == len(np.unique(w)) - 1
  for a, b in ((0, 1), (5, 700), (17, 333)):
    assert (ranks[a] == ranks[b]) == (w[a] == w[b]) and (ranks[a] < ranks[b]) == (w[a] < w[b])
  assert many.value_ranks() is None                      # 50 residues: more than 32 distinct values
  few.cached_array = np.array([0.0, -0.0, 99999.0, np.nan, np.nan, 2.0], dtype=np.float32)
  r = few.value_ranks()
  assert r[0] == r[1] and r[2] == -1 and r[3] != r[4] and r[5] > r[0]
  spec = few._dynamic_spec(3, 0.0)
  assert spec["ranks"] is r


def test_hdf5_layout_roundtrip(tmp_path):
  path = str(tmp_path / "o.h5")
  utils.dump_dict(path, "featurizer_parms", {"outfile": path, "cutoff": 5.0, "dimensions": [4, 4, 4], "time_window": 10, "byres": 0})
  rows = np.random.default_rng(0).random((70, 4, 4, 4)).astype(np.float32)
  for i in range(0, 70, 7):
    utils.append_hdf_data(path, "mass", rows[i:i + 7], dtype=np.float32, maxshape=(None, 4, 4, 4), chunks=(1, 4, 4, 4))
  with utils.batched_output():
    for i in range(70):
      utils.append_hdf_data(path, "gz", rows[i:i + 1], dtype=np.float32, maxshape=(None, 4, 4, 4), chunks=(1, 4, 4, 4),
                            compression="gzip", compression_opts=4)
  utils.dump_dict(path, "featurizer_parms", {"cutoff": 2.5, "name": "x"})     # recreated, not merged (utils.py:564-568)
  with utils.open_hdf(path, "r") as h:
    assert set(h.keys()) == {"featurizer_parms", "mass", "gz"}
    assert h["mass"].shape == (70, 4, 4, 4) and h["mass"].dtype == np.float32
    assert h["mass"].maxshape == (None, 4, 4, 4) and h["mass"].chunks == (1, 4, 4, 4)
    assert np.array_equal(h["mass"][:], rows) and np.array_equal(h["gz"][:], rows)
    assert np.array_equal(h["mass"][65], rows[65]) and np.array_equal(h["gz"][10:20], rows[10:20])
    assert set(h["featurizer_parms"].keys()) == {"cutoff", "name"}
    assert h["featurizer_parms/cutoff"][()] == 2.5 and h["featurizer_parms/name"][()] in ("x", b"x")
  if utils.hdf_backend() == "hdf5_lite":
    raw = open(path, "rb").read(16)
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 2
    from nearl_b200 import hdf5_lite
    assert hdf5_lite.lookup3(b"") == 0xDEADBEEF
    assert hdf5_lite.lookup3(b"Four score and seven years ago") == 0x17770551     # lookup3.c self-test vector


def test_focal_means_have_numpy_bits():
  """nrl_focal_means (host code of the native library) against np.mean(frame[selection], axis=0), the reference's focal point
  (featurizer.py:363-365): same order of additions, same precision, division in double -> identical float32 bits."""
  from nearl_b200 import _lib
  rng = np.random.default_rng(5)
  for dtype in (np.float32, np.float64):
    for trial in range(60):
      F, N = int(rng.integers(1, 9)), int(rng.integers(3, 3000))
      xyz = (rng.normal(0, 30, size=(F, N, 3)) + rng.normal(0, 100)).astype(dtype)
      n = int(rng.integers(1, N + 1))
      sel = np.sort(rng.choice(N, size=n, replace=False)) if trial % 2 else rng.choice(N, size=n, replace=True)
      W = int(rng.integers(1, F + 1))
      S = F // W
      want = np.array([np.mean(xyz[i * W][sel], axis=0) for i in range(S)], dtype=np.float32).reshape(S, 3)
      frames = np.arange(0, S * W, W, dtype=np.int32)
      s32 = np.ascontiguousarray(sel, dtype=np.int32)
      out = np.empty((S, 3), dtype=np.float32)
      _lib.check(_lib.lib().nrl_focal_means(_lib.ptr(xyz), int(dtype == np.float64), F, N, _lib.ptr(frames), S, _lib.ptr(s32), n, _lib.ptr(out)))
      assert np.array_equal(out, want), (dtype, trial)
  # through the featuriser: mask, index lists, float64 trajectory, a non-contiguous trajectory (numpy path)
  traj = io.synthetic_trajectory(900, 12, seed=4, pin=False)
  for xyz in (traj.xyz, traj.xyz.astype(np.float64), np.asfortranarray(traj.xyz)):
    t = io.Trajectory(xyz, traj.top)
    fz = featurizer.Featurizer({"dimensions": 8, "lengths": 8, "time_window": 5})
    fz.register_focus([":ARG", ":LYS,GLU"], "mask")
    fz.traj = t
    assert fz.parse_focus()
    for midx, mask in enumerate((":ARG", ":LYS,GLU")):
      sel = t.top.select(mask)
      want = np.array([np.mean(xyz[i * 5][sel], axis=0) for i in range(2)], dtype=np.float32)
      assert np.array_equal(fz.FOCALPOINTS[:, midx], want)
  with pytest.raises(ValueError):
    _lib.check(_lib.lib().nrl_focal_means(_lib.ptr(xyz), 0, 12, 900, _lib.ptr(np.array([12], dtype=np.int32)), 1, _lib.ptr(s32[:1]), 1, _lib.ptr(out)))
