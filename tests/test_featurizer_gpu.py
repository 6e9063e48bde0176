"""End-to-end parity of Featurizer.run (device-side focus boxes + batched kernels) against the
numpy restatement of the reference's host pipeline feeding the CPU oracle (oracle/host_ref.py),
and against this package's own per-task host protocol."""
import os

import numpy as np
import pytest

import nearl_b200
from nearl_b200 import _lib, features, featurizer, io, utils
from oracle import host_ref

pytestmark = pytest.mark.gpu
nearl_b200.update_config(quiet=True)


def rel_bad(new, ref, rtol=1e-5, atol=0.0):
  ref = np.asarray(ref, dtype=np.float64); new = np.asarray(new, dtype=np.float64)
  scale = np.maximum(np.abs(ref), 1e-5 * max(np.abs(ref).max(), 1e-30))
  return int((np.abs(new - ref) > rtol * scale + atol).sum())


def build(traj, tmp_path, feats, parms, focus=(":ARG",), native=True):
  out = str(tmp_path / ("native.h5" if native else "host.h5"))
  fz = featurizer.Featurizer({**parms, "outfile": out})
  fz.use_native_batch = native
  for f in feats():
    fz.register_feature(f)
  fz.register_focus(list(focus), "mask")
  fz.register_trajloader([traj])
  fz.main_loop()
  with utils.open_hdf(out, "r") as h:
    return {k: h[k][:] for k in h.keys() if k != "featurizer_parms"}, fz


def descs_from(fz):
  descs = []
  for f in fz.FEATURESPACE:
    d = {"outkey": f.outkey, "dims": f.dims, "spacing": f.spacing, "cutoff": f.cutoff, "sigma": f.sigma, "padding": f.padding,
         "byres": f.byres, "selection": (f.selected if f.selection is not None else None), "weights": f.cached_array,
         "frame_offset": f.frame_offset}
    if isinstance(f, features.MarchingObservers):
      d.update(kind="marching_observers", obs=f.obs, agg=f.agg)
    elif isinstance(f, features.DensityFlow):
      d.update(kind="density_flow", agg=f.agg)
    else:
      d.update(kind="static")
    descs.append(d)
  return descs


def test_readme_quickstart_mass(tmp_path):
  # README.rst:36-42 on a synthetic 5k-atom, 100-frame trajectory (BASELINE.json configs[0])
  traj = io.synthetic_trajectory(5000, 100, seed=20240713 + 1)
  parms = {"dimensions": 32, "lengths": 16, "time_window": 10}
  feats = lambda: [features.Mass(outkey="mass", sigma=1.5, cutoff=5.0)]
  got, fz = build(traj, tmp_path, feats, parms)
  assert got["mass"].shape == (10, 32, 32, 32) and got["mass"].dtype == np.float32
  ref = host_ref.featurize(traj.xyz, fz.FEATURESPACE[0].resids, descs_from(fz), fz.FOCALPOINTS, 10,
                           buff_dim=_lib.buff_dim)
  assert rel_bad(got["mass"], ref["mass"]) == 0
  host, _ = build(traj, tmp_path, feats, parms, native=False)
  assert rel_bad(got["mass"], host["mass"]) == 0
  with utils.open_hdf(str(tmp_path / "native.h5"), "r") as h:
    assert list(h["featurizer_parms/dimensions"][()].reshape(-1)) in ([32], [32, 32, 32]) or h["featurizer_parms/dimensions"][()] == 32
    assert h["mass"].chunks == (1, 32, 32, 32) and h["mass"].maxshape == (None, 32, 32, 32)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_mixed_static_dynamic_vs_host_pipeline(tmp_path, dtype):
  traj = io.synthetic_trajectory(3000, 20, seed=7, dtype=dtype)
  parms = {"dimensions": [24, 24, 24], "lengths": [12, 12, 12], "time_window": 5, "cutoff": 4.0, "sigma": 1.5}
  def feats():
    return [
      features.Mass(outkey="mass"),
      features.HeavyAtom(outkey="heavy", cutoff=2.5),
      features.AtomType(focus_element=6, outkey="carbon", frame_offset=2),
      features.Selection(selection=":ARG,LYS", outkey="sel", byres=True),
      features.Hydrophobicity(outkey="hyd", padding=2.0),
      features.MarchingObservers(obs="existence", agg="mean", outkey="obs_exist"),
      features.MarchingObservers(obs="direct_count", agg="standard_deviation", weight_type="mass", outkey="obs_count"),
      features.MarchingObservers(obs="distinct_count", agg="max", weight_type="atomic_number", outkey="obs_distinct"),
      features.MarchingObservers(obs="distinct_count", agg="mean", weight_type="residue_id", outkey="obs_distinct_res"),
      features.MarchingObservers(obs="mean_distance", agg="mean", weight_type="mass", outkey="obs_md"),
      features.MarchingObservers(obs="radius_of_gyration", agg="median", weight_type="mass", outkey="obs_rg"),
      features.DensityFlow(agg="mean", weight_type="mass", outkey="pdf_mean"),
      features.DensityFlow(agg="standard_deviation", weight_type="electronegativity", outkey="pdf_std", selection="!:ARG"),
    ]
  got, fz = build(traj, tmp_path, feats, parms, focus=(":ARG", ":1-5"))
  ref = host_ref.featurize(traj.xyz, fz.FEATURESPACE[0].resids, descs_from(fz), fz.FOCALPOINTS, 5,
                           buff_dim=_lib.buff_dim)
  host, _ = build(traj, tmp_path, feats, parms, focus=(":ARG", ":1-5"), native=False)
  for key in ref:
    assert got[key].shape == (8, 24, 24, 24), key
    if key == "obs_distinct":
      assert np.array_equal(got[key], ref[key]), key         # integer observable, max over frames: bit-exact
      assert np.array_equal(got[key], host[key]), key
    elif key in ("obs_exist", "obs_distinct_res"):
      # exact integers per frame; their mean goes through div.approx (<= 2 ulp from IEEE)
      assert np.allclose(got[key], ref[key], rtol=3e-7, atol=0), key
      assert np.array_equal(got[key], host[key]), key
    elif key in ("obs_count", "pdf_std"):
      scale = max(float(np.abs(ref[key]).max()), 1e-30)       # std cancels: tolerance relative to the grid scale
      assert rel_bad(got[key], ref[key], atol=1e-5 * scale) == 0, key
    else:
      assert rel_bad(got[key], ref[key]) == 0, (key, np.abs(got[key] - ref[key]).max())
      assert rel_bad(got[key], host[key]) == 0, key


def test_empty_focus_box_and_missing_selection(tmp_path):
  traj = io.synthetic_trajectory(800, 4, seed=3)
  fz = featurizer.Featurizer({"dimensions": 8, "lengths": 4, "time_window": 2, "cutoff": 2.0, "sigma": 1.0,
                              "outfile": str(tmp_path / "e.h5")})
  fz.register_feature(features.Mass(outkey="mass"))
  fz.register_feature(features.MarchingObservers(obs="eccentricity", agg="mean", outkey="ecc"))
  fz.register_focus(np.array([[500.0, 500.0, 500.0]]), "absolute")     # nothing near this point
  fz.register_trajloader([traj])
  fz.run()
  with utils.open_hdf(str(tmp_path / "e.h5"), "r") as h:
    assert np.all(h["mass"][:] == 0) and h["mass"].shape == (2, 8, 8, 8)
    assert np.all(h["ecc"][:] == np.float32(2.0))
  fz2 = featurizer.Featurizer({"dimensions": 8, "lengths": 4, "time_window": 2, "cutoff": 2.0, "sigma": 1.0})
  fz2.register_feature(features.Mass(outkey="mass", outfile=str(tmp_path / "f.h5")))
  fz2.register_focus([":XYZ"], "mask")
  fz2.register_trajloader([traj])
  stats = fz2.run()
  assert stats["trajectories"] == 0 and not os.path.exists(str(tmp_path / "f.h5"))


def test_custom_feature_subclass_uses_host_protocol(tmp_path):
  class Doubled(features.Mass):
    def run(self, coords, weights):
      return 2 * super().run(coords, weights)
  traj = io.synthetic_trajectory(1500, 6, seed=11)
  parms = {"dimensions": 16, "lengths": 8, "time_window": 3, "cutoff": 3.0, "sigma": 1.0}
  got, fz = build(traj, tmp_path, lambda: [features.Mass(outkey="m"), Doubled(outkey="d")], parms)
  assert fz.last_run_stats["native_tasks"] == 2 and fz.last_run_stats["tasks"] == 4
  assert rel_bad(got["d"], 2 * got["m"]) == 0


def test_atom_cap_and_sorted_emission(tmp_path):
  """More atoms in the focus box than DynamicFeature.MAX_ALLOWED_ATOMS: the reference keeps the
  first N in topology order (features.py:1353-1359); the device crop must do the same, also when
  it emits atoms in weight-value order for distinct_count."""
  traj = io.synthetic_trajectory(6000, 4, seed=21)
  parms = {"dimensions": 16, "lengths": 24, "time_window": 2, "cutoff": 3.0, "sigma": 1.5}
  def feats():
    fs = [features.MarchingObservers(obs="direct_count", agg="max", weight_type="mass", outkey="count"),
          features.MarchingObservers(obs="distinct_count", agg="max", weight_type="residue_id", outkey="distinct"),
          features.DensityFlow(agg="mean", weight_type="mass", outkey="flow")]
    for f in fs:
      f.MAX_ALLOWED_ATOMS = 700
    return fs
  got, fz = build(traj, tmp_path, feats, parms)
  assert fz.last_run_stats["native_tasks"] == 6
  from oracle import host_ref as hr
  old = hr.MAX_ALLOWED_ATOMS
  hr.MAX_ALLOWED_ATOMS = 700
  try:
    ref = hr.featurize(traj.xyz, fz.FEATURESPACE[0].resids, descs_from(fz), fz.FOCALPOINTS, 2, buff_dim=_lib.buff_dim)
  finally:
    hr.MAX_ALLOWED_ATOMS = old
  assert np.array_equal(got["count"], ref["count"])
  assert np.array_equal(got["distinct"], ref["distinct"])
  assert rel_bad(got["flow"], ref["flow"]) == 0
  uncapped = hr.featurize(traj.xyz, fz.FEATURESPACE[0].resids, descs_from(fz)[:1], fz.FOCALPOINTS, 2)
  assert not np.array_equal(uncapped["count"], ref["count"])      # the cap really bites on this system


@pytest.mark.parametrize("n_atoms", [16, 33, 480, 1500, 4097])
def test_ragged_atom_counts(tmp_path, n_atoms):
  """Device crop / compaction at atom counts around warp, word and segment boundaries."""
  traj = io.synthetic_trajectory(n_atoms, 4, seed=100 + n_atoms)
  parms = {"dimensions": 8, "lengths": 8, "time_window": 2, "cutoff": 3.0, "sigma": 1.0}
  feats = lambda: [features.Mass(outkey="m"),
                   features.MarchingObservers(obs="direct_count", agg="max", outkey="n"),
                   features.MarchingObservers(obs="distinct_count", agg="max", weight_type="residue_id", outkey="r", byres=True)]
  got, fz = build(traj, tmp_path, feats, parms, focus=(":1",))
  ref = host_ref.featurize(traj.xyz, fz.FEATURESPACE[0].resids, descs_from(fz), fz.FOCALPOINTS, 2, buff_dim=_lib.buff_dim)
  assert np.array_equal(got["n"], ref["n"]) and np.array_equal(got["r"], ref["r"])
  assert rel_bad(got["m"], ref["m"]) == 0


def test_shared_sweeps_multi_reducer_multi_weight(tmp_path):
  """Features that share a crop and a grid are lowered to multi-channel work items (several
  reducers of one observable, several weight tables per Gaussian sweep): results must equal the
  one-task-at-a-time oracle."""
  traj = io.synthetic_trajectory(2500, 12, seed=33)
  parms = {"dimensions": 16, "lengths": 8, "time_window": 6, "cutoff": 3.5, "sigma": 1.2}
  def feats():
    fs = [features.MarchingObservers(obs="direct_count", agg=a, outkey=f"count_{a}")
          for a in ("mean", "standard_deviation", "median", "max", "min", "drift", "information_entropy")]
    fs += [features.MarchingObservers(obs="mean_distance", agg=a, weight_type="mass", outkey=f"md_{a}") for a in ("median", "mean")]
    fs += [features.DensityFlow(agg="mean", weight_type="mass", outkey="pdf_mass_mean"),
           features.DensityFlow(agg="median", weight_type="mass", outkey="pdf_mass_median"),
           features.DensityFlow(agg="standard_deviation", weight_type="electronegativity", outkey="pdf_en_std"),
           features.DensityFlow(agg="max", weight_type="hydrophobicity", outkey="pdf_hyd_max"),
           features.DensityFlow(agg="mean", weight_type="atomic_number", outkey="pdf_z_mean"),
           features.DensityFlow(agg="min", weight_type="heavy_atom", outkey="pdf_heavy_min")]
    fs += [features.Mass(outkey="s_mass"), features.AtomType(focus_element=6, outkey="s_c"), features.AtomType(focus_element=8, outkey="s_o"),
           features.HeavyAtom(outkey="s_heavy"), features.Electronegativity(outkey="s_en"), features.AtomicNumber(outkey="s_z")]
    return fs
  got, fz = build(traj, tmp_path, feats, parms)
  ref = host_ref.featurize(traj.xyz, fz.FEATURESPACE[0].resids, descs_from(fz), fz.FOCALPOINTS, 6, buff_dim=_lib.buff_dim)
  for key in ref:
    if key in ("count_median", "count_max", "count_min"):
      assert np.array_equal(got[key], ref[key]), key
    elif key == "count_information_entropy":
      assert np.count_nonzero(~np.isclose(got[key], ref[key], atol=1e-3)) < 4, key
    elif key.startswith("count_") or key in ("pdf_en_std",):
      scale = max(float(np.abs(ref[key]).max()), 1e-30)
      assert rel_bad(got[key], ref[key], atol=2e-5 * scale) == 0, key
    else:
      assert rel_bad(got[key], ref[key]) == 0, (key, np.abs(got[key] - ref[key]).max())


def test_fused_observers_give_the_bits_of_separate_sweeps(tmp_path):
  """Observers of one focus box share a sweep whatever their observable (csrc/cell_kernels.cuh, NRL_OBS_FUSED);
  every accumulator is updated with the arithmetic of its single-observable kernel, so the results are bit-identical
  to one sweep per observable -- for streaming reducers and for reducers that keep a series."""
  traj = io.synthetic_trajectory(3000, 48, seed=77, pin=False)

  def feats(aggs):
    a = iter(aggs)
    return [features.MarchingObservers(obs="existence", agg=next(a), weight_type="mass", outkey="exist"),
            features.MarchingObservers(obs="direct_count", agg=next(a), weight_type="mass", outkey="count"),
            features.MarchingObservers(obs="distinct_count", agg=next(a), weight_type="atomic_number", outkey="distinct"),
            features.MarchingObservers(obs="mean_distance", agg=next(a), weight_type="mass", outkey="mean_distance"),
            features.MarchingObservers(obs="cumulative_weight", agg=next(a), weight_type="mass", outkey="cumw"),
            features.MarchingObservers(obs="density", agg=next(a), weight_type="mass", outkey="density"),
            features.MarchingObservers(obs="eccentricity", agg=next(a), weight_type="mass", outkey="ecc"),
            # a different weight table: cannot join the mass group's float channels
            features.MarchingObservers(obs="cumulative_weight", agg=next(a), weight_type="electronegativity", outkey="cumw_en")]

  for window, aggs in ((6, ["mean", "max", "mean", "mean", "min", "mean", "max", "mean"]),                       # all streaming
                       (6, ["standard_deviation", "median", "mean", "drift", "variance", "information_entropy", "mean", "max"]),
                       # 24 frames x 4 series rows x 512 voxels: the series of a fused tile leaves shared memory
                       (24, ["standard_deviation", "median", "variance", "drift", "information_entropy", "mean", "median", "max"])):
    parms = {"dimensions": 24, "lengths": 14, "time_window": window, "cutoff": 4.0, "sigma": 1.2}
    got = {}
    for fuse in (True, False):
      _lib.set_option("no_fuse", 0 if fuse else 1)
      try:
        fz = featurizer.Featurizer(parms)
        fz.register_features(feats(aggs))
        fz.register_focus([":ARG"], "mask")
        fz.register_trajloader([traj])
        fz.keep_results = True
        launches = _lib.kernel_launches()
        fz.run()
        got[fuse] = ({k: np.array(v) for k, v in fz.RESULTS.items()}, _lib.kernel_launches() - launches)
      finally:
        _lib.set_option("no_fuse", 0)
    assert got[True][1] < got[False][1]                        # fewer launches: the sweeps really were shared
    for key, want in got[False][0].items():
      assert np.array_equal(got[True][0][key], want, equal_nan=True), (aggs, key)
    # and against the oracle
    fz.traj = traj
    fz.parse_focus()
    for f in fz.FEATURESPACE:
      f.cache(traj)
    ref = host_ref.featurize(traj.xyz, fz.FEATURESPACE[0].resids, descs_from(fz), fz.FOCALPOINTS, window, buff_dim=_lib.buff_dim, scrub_nan=False,
                             max_slices=2)
    for key in ("exist", "count", "distinct"):
      r = ref[key]
      if aggs[("exist", "count", "distinct").index(key)] == "information_entropy":
        assert np.count_nonzero(~np.isclose(got[True][0][key][:len(r)], r, atol=1e-3)) < 10, key
      else:
        assert np.allclose(got[True][0][key][:len(r)], r, rtol=1e-5, atol=1e-5), key


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cluster_crop_byres_selection_and_cap(tmp_path, dtype):
  """Long frames in topology order go through the cluster form of the crop kernel (four CTAs per frame, counts
  over distributed shared memory): by-residue expansion across chunk borders, selection masks and the atom cap
  must come out exactly as the host pipeline's."""
  traj = io.synthetic_trajectory(9000, 6, seed=91, dtype=dtype)
  parms = {"dimensions": 20, "lengths": 14, "time_window": 3, "cutoff": 3.5, "sigma": 1.2}

  def feats():
    capped = features.MarchingObservers(obs="direct_count", agg="max", weight_type="mass", outkey="count_capped", byres=True)
    capped.MAX_ALLOWED_ATOMS = 300                         # the first 300 in-box atoms in topology order
    return [features.Mass(outkey="mass", byres=True, padding=2.0),
            features.Selection(selection=":ARG,LYS,GLU", outkey="sel"),
            features.MarchingObservers(obs="mean_distance", agg="mean", weight_type="mass", outkey="mean_distance", byres=True),
            features.MarchingObservers(obs="direct_count", agg="mean", weight_type="mass", outkey="count", selection=":ALA,GLY,SER,THR"),
            capped]

  got, fz = build(traj, tmp_path, feats, parms)
  descs = descs_from(fz)
  for d, f in zip(descs, fz.FEATURESPACE):
    d["max_atoms"] = getattr(f, "MAX_ALLOWED_ATOMS", None)
  host, _ = build(traj, tmp_path, feats, parms, native=False)
  for key in got:
    if key in ("count", "count_capped"):
      assert np.array_equal(got[key], host[key]), key
    else:
      assert rel_bad(got[key], host[key]) == 0, (key, np.abs(got[key] - host[key]).max())
  assert got["count_capped"].max() <= 300 and got["count"].max() > 0 and got["sel"].sum() > 0


def test_single_frame_systems_grouped_per_call(tmp_path):
  """Static featurisation of single-frame systems sends `static_group` systems per native call (each a slice of a
  concatenated pseudo-system with its own focus boxes); results and their order equal one call per system."""
  systems = [io.synthetic_complex(1200 + 37 * i, seed=300 + i) for i in range(11)]
  systems.insert(5, io.synthetic_trajectory(900, 2, seed=299, pin=False))       # a two-frame system: falls back, order kept

  def run(group):
    fz = featurizer.Featurizer({"dimensions": 20, "lengths": 16, "time_window": 1, "cutoff": 3.0, "sigma": 1.2})
    fz.static_group = group
    fz.register_features([features.Mass(outkey="mass"), features.HeavyAtom(outkey="heavy", byres=True, padding=1.5),
                          features.AtomType(focus_element=8, outkey="oxygen"), features.Selection(selection=":ARG,LYS", outkey="sel")])
    fz.register_focus([":ARG", ":LYS"], "mask")                 # two focal points per system
    fz.register_trajloader(systems)
    fz.keep_results = True
    launches = _lib.kernel_launches()
    stats = fz.run()
    return {k: np.array(v) for k, v in fz.RESULTS.items()}, stats, _lib.kernel_launches() - launches

  grouped, gstats, glaunch = run(4)              # three groups of four; the one holding the two-frame system falls back
  single, sstats, slaunch = run(1)
  assert gstats["tasks"] == sstats["tasks"] == (11 * 2 + 2 * 2) * 4 and gstats["native_tasks"] == gstats["tasks"]
  assert glaunch < slaunch
  for key, want in single.items():
    assert grouped[key].shape == want.shape == (26, 20, 20, 20)
    assert np.array_equal(grouped[key], want), key
  assert grouped["mass"].sum() > 0 and grouped["sel"].sum() > 0


def test_recorded_reference_feature_set_on_the_device_path(tmp_path):
  # The systems and the 15-feature set whose FFI-boundary vectors were recorded from the reference's own Python
  # (tests/golden/make_host_golden.py; non-cubic 12 x 14 x 16 grid, per-feature cutoff / sigma / padding, by-residue,
  # selections, frame offset, three focal masks): the device path (crop, cell lists, fused sweeps, fused reducers)
  # against oracle/host_ref.py -- which tests/test_host_golden.py pins to that record -- feeding the C oracle.
  import sys
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
  import make_host_golden as G
  gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_vectors.npz"))
  for name, kw in G.systems():
    traj = G.make_system(kw)
    (tmp_path / name).mkdir()
    got, fz = build(traj, tmp_path / name, lambda: G.feature_set(features), dict(G.FEATURIZER_PARMS), focus=G.FOCUS)
    assert np.array_equal(fz.FOCALPOINTS.view(np.int32), gold[f"{name}/focal"].view(np.int32))
    ref = host_ref.featurize(traj.xyz, fz.FEATURESPACE[0].resids, descs_from(fz), fz.FOCALPOINTS, fz.time_window, buff_dim=_lib.buff_dim)
    for f in fz.FEATURESPACE:
      key = f.outkey
      assert got[key].shape == ref[key].shape, key
      if isinstance(f, features.MarchingObservers) and f.obs in (1, 2, 3) and f.agg in (3, 5, 6):
        assert np.array_equal(got[key], ref[key]), key                                   # integer observable, order statistic
      elif getattr(f, "agg", 0) in (2, 4, 8):
        peak = max(float(np.abs(ref[key]).max()), 1e-30)
        assert rel_bad(got[key], ref[key], atol=2e-5 * peak) == 0, key                    # cancelling reducers
      elif getattr(f, "agg", 0) == 7:
        assert np.count_nonzero(~np.isclose(got[key], ref[key], atol=1e-3)) <= max(2, got[key].size // 3000), key
      else:
        assert rel_bad(got[key], ref[key]) == 0, (key, np.abs(got[key] - ref[key]).max())


def test_label_features_ride_along_pipelined_and_grouped(tmp_path):
  # The reference's production scripts register LabelAffinity / LabelStepping / LabelPCDT next to the voxel features
  # (scripts/featurize_misato.py:214-217, featurize_pdbbind.py:92).  Labels are host-protocol tasks: they must not stop
  # the trajectory pipeline nor the grouping of single-frame systems, must land in task order, and must leave the grids
  # exactly as a run without them.
  table = str(tmp_path / "table.csv")
  codes = [f"{i:d}abc" for i in range(1, 10)]
  with open(table, "w") as f:
    f.write("pdbcode,pK\n" + "".join(f"{c},{2.0 + 0.75 * i}\n" for i, c in enumerate(codes)))

  def run(trajs, parms, with_labels, group=32, depth=3):
    fz = featurizer.Featurizer(parms)
    feats = [features.Mass(outkey="mass"), features.HeavyAtom(outkey="heavy", byres=True)]
    if parms["time_window"] > 1:
      feats.append(features.MarchingObservers(obs="existence", agg="mean", weight_type="mass", outkey="obs"))
    if with_labels:
      feats.insert(1, features.LabelAffinity(baseline_map=table, outkey="pk"))
      feats.append(features.LabelStepping(baseline_map=table, outkey="step"))
      if parms["time_window"] > 1:
        feats.append(features.LabelPCDT(selection=":LYS&!@H=", baseline_map=table, outkey="pcdt"))
    fz.register_features(feats)
    fz.register_focus([":ARG", "@CA"], "mask")
    fz.register_trajloader(trajs)
    fz.static_group, fz.pipeline_depth = group, depth
    rows = []
    fz.result_sink = lambda outkey, tid, bid, pid, result: rows.append((outkey, tid, bid, pid, np.array(result)))
    fz.run()
    return rows, fz

  # multi-frame trajectories: pipelined over three lanes
  trajs = [io.synthetic_trajectory(1500 + 100 * i, 12, seed=50 + i) for i in range(5)]
  for t, c in zip(trajs, codes):
    t.identity = c.upper()
  parms = {"dimensions": 16, "lengths": 12, "time_window": 4, "cutoff": 3.0, "sigma": 1.2}
  plain, _ = run(trajs, parms, False)
  mixed, fz = run(trajs, parms, True)
  assert fz.last_run_stats["native_tasks"] == len(plain)
  grids = [r for r in mixed if r[0] in ("mass", "heavy", "obs")]
  assert len(grids) == len(plain)
  for a, b in zip(grids, plain):
    assert a[:4] == b[:4] and np.array_equal(a[4], b[4])
  order = [r[0] for r in mixed[:6]]
  assert order == ["mass", "pk", "heavy", "obs", "step", "pcdt"]                      # registration order inside a task group
  for r in mixed:
    if r[0] == "pk":
      assert float(r[4]) == 2.0 + 0.75 * r[1]
    if r[0] == "step":
      assert float(r[4]) == min((2.0 + 0.75 * r[1]) // 2, 5)
  # PCDT against a serial, unpipelined run of the same featuriser
  serial, _ = run(trajs, parms, True, depth=1)
  assert [float(r[4]) for r in mixed if r[0] == "pcdt"] == [float(r[4]) for r in serial if r[0] == "pcdt"]

  # single-frame systems: grouped per native call
  systems = [io.synthetic_complex(900 + 50 * i, seed=70 + i) for i in range(9)]
  for t, c in zip(systems, codes):
    t.identity = c
  sparms = {"dimensions": 12, "lengths": 10, "time_window": 1, "cutoff": 2.5, "sigma": 1.0}
  launches0 = _lib.kernel_launches()
  plain, _ = run(systems, sparms, False, group=4)
  launches_plain = _lib.kernel_launches() - launches0
  launches0 = _lib.kernel_launches()
  mixed, fz = run(systems, sparms, True, group=4)
  launches_mixed = _lib.kernel_launches() - launches0
  assert launches_mixed == launches_plain                                             # labels did not break the grouping
  one_by_one, _ = run(systems, sparms, True, group=1)
  assert [(r[0], r[1], r[2], r[3]) for r in mixed] == [(r[0], r[1], r[2], r[3]) for r in one_by_one]
  for a, b in zip(mixed, one_by_one):
    assert np.array_equal(a[4], b[4]), a[:4]
  grids = [r for r in mixed if r[0] in ("mass", "heavy")]
  for a, b in zip(grids, plain):
    assert a[:4] == b[:4] and np.array_equal(a[4], b[4])


def test_density_flow_selections_share_one_sweep(tmp_path):
  # scripts/featurize_misato.py:163-166: density flows on the ligand and on the protein selection, mean and standard deviation.
  # Their selections are folded into the weights (an unselected atom = a zero-weight atom for the voxeliser), so the four
  # channels share one focus box and one multi-channel sweep; by-residue flows and systems above the 9 999-atom cap of the
  # dynamic query (no folding there) take the separate-crop path.  All against the host pipeline feeding the C oracle.
  def feats():
    out = []
    for sel, tag in ((":LIG", "lig"), ("!:LIG", "prot")):
      for agg in ("mean", "standard_deviation"):
        out.append(features.DensityFlow(selection=sel, weight_type="uniformed", agg=agg, outkey=f"pdf_{agg[:2]}_{tag}"))
    out.append(features.DensityFlow(selection=":ARG,LYS", weight_type="mass", agg="max", byres=True, outkey="pdf_byres"))
    out.append(features.MarchingObservers(selection=":LIG", weight_type="uniformed", obs="distinct_count", agg="standard_deviation", outkey="mo_lig"))
    return out
  parms = {"dimensions": 20, "lengths": 15, "time_window": 4, "cutoff": 2.5, "sigma": 1.5}
  for n_atoms, expect_folded in ((2500, True), (10400, False)):
    traj = io.synthetic_trajectory(n_atoms, 8, seed=300 + n_atoms, ligand_atoms=30)
    (tmp_path / str(n_atoms)).mkdir()
    launches0 = _lib.kernel_launches()
    got, fz = build(traj, tmp_path / str(n_atoms), feats, parms, focus=(":LIG",))
    launches = _lib.kernel_launches() - launches0
    spec = fz.FEATURESPACE[0]._native_spec()
    assert spec["selection_folded"] is expect_folded
    ref = host_ref.featurize(traj.xyz, fz.FEATURESPACE[0].resids, descs_from(fz), fz.FOCALPOINTS, 4, buff_dim=_lib.buff_dim)
    for f in fz.FEATURESPACE:
      key = f.outkey
      if f.agg == 2:
        peak = max(float(np.abs(ref[key]).max()), 1e-30)
        assert rel_bad(got[key], ref[key], atol=2e-5 * peak) == 0, (n_atoms, key)
      else:
        assert rel_bad(got[key], ref[key]) == 0, (n_atoms, key, np.abs(got[key] - ref[key]).max())
    assert launches > 0


def test_fused_presence_bit_and_its_exclusions(tmp_path):
  # A fused sweep with a distinct_count channel keeps no in-range counter: bit 31 of every atom's rank mask says "an atom
  # is in range".  So (1) a rank table that uses all 32 ranks stays on the single-observable kernel, (2) a direct count
  # never shares a group with a rank mask, (3) atoms whose value distinct_count hides (99999.0) still count for existence.
  # All three against the host pipeline feeding the C oracle; integer observables with max: bit-exact.
  traj = io.synthetic_trajectory(2400, 6, seed=91)
  parms = {"dimensions": 16, "lengths": 12, "time_window": 3, "cutoff": 3.5, "sigma": 1.0}

  class Ranks32(features.MarchingObservers):          # 32 distinct weight values: ranks 0..31
    def cache(self, trajectory):
      super().cache(trajectory)
      self.cached_array = (np.arange(len(self.resids)) % 32).astype(np.float32)

  class Hidden(features.MarchingObservers):           # every third atom carries the placeholder distinct_count skips
    def cache(self, trajectory):
      super().cache(trajectory)
      w = (np.arange(len(self.resids)) % 7).astype(np.float32)
      w[::3] = 99999.0
      self.cached_array = w

  def feats():
    return [
      features.MarchingObservers(obs="existence", agg="max", outkey="exist"),
      Ranks32(obs="distinct_count", agg="max", weight_type="atomic_number", outkey="d32"),
      Hidden(obs="distinct_count", agg="max", weight_type="atomic_number", outkey="dhid"),
      features.MarchingObservers(obs="direct_count", agg="max", outkey="count"),
      features.MarchingObservers(obs="distinct_count", agg="min", weight_type="atomic_number", outkey="dz"),
      features.MarchingObservers(obs="mean_distance", agg="mean", weight_type="mass", outkey="md"),
    ]
  got, fz = build(traj, tmp_path, feats, parms, focus=(":ARG", "@CA"))
  assert fz.FEATURESPACE[1].value_ranks().max() == 31
  ref = host_ref.featurize(traj.xyz, fz.FEATURESPACE[0].resids, descs_from(fz), fz.FOCALPOINTS, 3, buff_dim=_lib.buff_dim)
  for key in ("exist", "d32", "dhid", "count", "dz"):
    assert np.array_equal(got[key], ref[key]), key
  assert got["d32"].max() > 20 and got["dhid"].max() <= 7 and got["exist"].max() == 1.0
  assert rel_bad(got["md"], ref["md"]) == 0
  # and the same bits as one sweep per observable
  _lib.set_option("no_fuse", 1)
  try:
    (tmp_path / "nofuse").mkdir()
    sep, _ = build(traj, tmp_path / "nofuse", feats, parms, focus=(":ARG", "@CA"))
  finally:
    _lib.set_option("no_fuse", 0)
  for key in got:
    assert np.array_equal(got[key].view(np.int32), sep[key].view(np.int32)), key


def test_skipped_trajectories_do_not_alias_pipeline_lanes(tmp_path):
  """A trajectory without focal point is skipped and takes no lane: with three lanes and the second of seven trajectories
  skipped, the fourth one used to land on the lane of the first while that was still pending (its page-locked result
  buffer was overwritten: the first trajectory came back with the fourth's grids)."""
  trajs = [io.synthetic_trajectory(1500 + 50 * i, 4, seed=700 + i, ligand_atoms=(0 if i in (1, 4) else 20)) for i in range(7)]

  def run(depth, budget=None):
    fz = featurizer.Featurizer({"dimensions": 12, "lengths": 9, "time_window": 2, "cutoff": 3.0, "sigma": 1.0})
    fz.pipeline_depth = depth
    if budget is not None:
      fz.pipeline_bytes = budget
    fz.register_features([features.Mass(outkey="mass"), features.MarchingObservers(obs="direct_count", agg="max", outkey="count")])
    fz.register_focus([":LIG"], "mask")
    fz.register_trajloader(trajs)
    seen, lanes = [], []
    real = featurizer.LoweredBatch.execute
    featurizer.LoweredBatch.execute = lambda self, *a, **kw: lanes.append(kw.get("lane")) or real(self, *a, **kw)
    fz.result_sink = lambda outkey, tid, bid, pid, grid: seen.append((outkey, tid, bid, np.array(grid)))
    try:
      stats = fz.run()
    finally:
      featurizer.LoweredBatch.execute = real
    stats["lanes"] = lanes
    return seen, stats

  serial, s1 = run(1)
  piped, s3 = run(3)
  # a byte budget of two of the largest trajectories narrows the pipeline to two lanes, of less than one to a single lane
  two, s2 = run(8, budget=2 * max(t.xyz.nbytes for t in trajs))
  one, s0 = run(8, budget=1000)
  assert set(s2["lanes"]) == {0, 1} and set(s0["lanes"]) == {0} and set(s3["lanes"]) == {0, 1, 2}
  for a, b, c in zip(serial, two, one):
    assert a[:3] == b[:3] == c[:3] and np.array_equal(a[3], b[3]) and np.array_equal(a[3], c[3])
  assert s1["trajectories"] == s3["trajectories"] == 5 and [t for _, t, _, _ in serial] == [t for _, t, _, _ in piped]
  assert sorted({t for _, t, _, _ in piped}) == [0, 2, 3, 5, 6]
  for (k1, t1, b1, g1), (k3, t3, b3, g3) in zip(serial, piped):
    assert (k1, t1, b1) == (k3, t3, b3) and np.array_equal(g1, g3), (k1, t1, b1)
  assert len({g.tobytes() for k, _, _, g in piped if k == "mass"}) == 10          # ten different grids, none delivered twice


def test_static_groups_then_multi_frame_trajectories_stay_pipelined(tmp_path):
  """Single-frame systems first, multi-frame trajectories after: grouping is switched off at the first multi-frame
  trajectory and the REST of the list goes through the pipelined loop (it used to run on lane 0, one at a time)."""
  systems = [io.synthetic_complex(1000 + 31 * i, seed=800 + i) for i in range(6)]
  systems += [io.synthetic_trajectory(1100 + 17 * i, 3, seed=820 + i, pin=False, ligand_atoms=30) for i in range(5)]
  lanes = []
  real = featurizer.LoweredBatch.execute

  def run(group, depth):
    fz = featurizer.Featurizer({"dimensions": 16, "lengths": 12, "time_window": 1, "cutoff": 3.0, "sigma": 1.2})
    fz.static_group, fz.pipeline_depth = group, depth
    fz.register_features([features.Mass(outkey="mass"), features.AtomType(focus_element=6, outkey="carbon")])
    fz.register_focus([":LIG"], "mask")
    fz.register_trajloader(systems)
    fz.keep_results = True
    stats = fz.run()
    return {k: np.array(v) for k, v in fz.RESULTS.items()}, stats

  featurizer.LoweredBatch.execute = lambda self, *a, **kw: lanes.append(kw.get("lane")) or real(self, *a, **kw)
  try:
    grouped, gs = run(4, 3)
  finally:
    featurizer.LoweredBatch.execute = real
  plain, ps = run(1, 1)
  assert gs["tasks"] == ps["tasks"] == (6 + 5 * 3) * 2
  for key in plain:
    assert np.array_equal(grouped[key], plain[key]), key
  assert len(set(lanes[-4:])) == 3                                       # the tail of the list uses all three lanes


@pytest.mark.parametrize("dtype,n_atoms", [(np.float32, 9000), (np.float64, 9000), (np.float32, 2051), (np.float32, 30001)])
def test_streaming_crop_keeps_the_same_atoms_in_the_same_order(tmp_path, dtype, n_atoms):
  """Whole frames in topology order without by-residue expansion go through the streaming crop kernel (every warp a
  contiguous run, 16-byte loads, kept atoms compacted on the way): selections, the atom cap, float64 coordinates, atom
  counts that are no multiple of 4 (unaligned frames: scalar loads) and of 128 -- results and per-frame atom counts are
  those of the first crop kernel, bit for bit, and the host pipeline's."""
  traj = io.synthetic_trajectory(n_atoms, 6, seed=93, dtype=dtype)
  parms = {"dimensions": 20, "lengths": 14, "time_window": 3, "cutoff": 3.5, "sigma": 1.2}

  def feats():
    capped = features.MarchingObservers(obs="direct_count", agg="max", weight_type="mass", outkey="count_capped")
    capped.MAX_ALLOWED_ATOMS = 300                         # the first 300 in-box atoms in topology order
    return [features.Mass(outkey="mass", padding=2.0),
            features.Selection(selection=":ARG,LYS,GLU", outkey="sel"),
            features.MarchingObservers(obs="mean_distance", agg="mean", weight_type="mass", outkey="mean_distance"),
            features.MarchingObservers(obs="direct_count", agg="mean", weight_type="mass", outkey="count", selection=":ALA,GLY,SER,THR"),
            features.DensityFlow(agg="standard_deviation", weight_type="mass", outkey="flow", frame_offset=1),
            capped]

  out = {}
  for legacy in (0, 1):
    _lib.set_option("crop_legacy", legacy)
    try:
      sub = tmp_path / f"legacy{legacy}"
      sub.mkdir()
      out[legacy], fz = build(traj, sub, feats, parms)
    finally:
      _lib.set_option("crop_legacy", 0)
  for key in out[0]:
    assert np.array_equal(out[0][key], out[1][key]), key
  host, _ = build(traj, tmp_path, feats, parms, native=False)
  for key in out[0]:
    if key in ("count", "count_capped"):
      assert np.array_equal(out[0][key], host[key]), key
    else:
      scale = max(float(np.abs(host[key]).max()), 1e-30)
      assert rel_bad(out[0][key], host[key], atol=1e-5 * scale if key == "flow" else 0.0) == 0, (key, np.abs(out[0][key] - host[key]).max())
  assert out[0]["count_capped"].max() <= 300 and out[0]["count"].max() > 0 and out[0]["sel"].sum() > 0


def test_partial_charge_channels_on_the_device_path(tmp_path):
  """PartialCharge deposits signed weights: the batched device path against the per-task host protocol, and
  q == max(q, 0) + min(q, 0) voxel by voxel (the channels share one focus box and one sweep)."""
  traj = io.synthetic_trajectory(3000, 4, seed=77, pin=False, ligand_atoms=24)
  traj.top.charge = np.round(np.random.default_rng(7).normal(0.0, 0.35, size=traj.n_atoms), 4)
  traj.top.charge[::9] = 0.0                                       # exactly neutral atoms are skipped like any zero weight
  parms = {"dimensions": 24, "lengths": 14, "time_window": 2, "cutoff": 3.5, "sigma": 1.3}

  def feats():
    return [features.PartialCharge(outkey="q"), features.PartialCharge(keep_sign="p", outkey="qp"),
            features.PartialCharge(keep_sign="n", outkey="qn", selection="!:LIG"),
            features.PartialCharge(charge_type="external", charge_parm=lambda t: np.abs(t.top.charge), outkey="qa")]

  got, fz = build(traj, tmp_path, feats, parms, focus=(":LIG",))
  assert fz.last_run_stats["native_tasks"] == fz.last_run_stats["tasks"] == 2 * 4
  host, _ = build(traj, tmp_path, feats, parms, focus=(":LIG",), native=False)
  for key in ("q", "qp", "qn", "qa"):
    scale = max(float(np.abs(host[key]).max()), 1e-30)
    assert np.abs(got[key] - host[key]).max() <= 2e-5 * scale, (key, np.abs(got[key] - host[key]).max())
  assert got["qn"].max() <= 0.0 and got["qp"].min() >= 0.0 and got["q"].min() < 0 < got["q"].max()
  ref = descs_from(fz)
  want = host_ref.featurize(traj.xyz, fz.FEATURESPACE[0].resids, ref, fz.FOCALPOINTS, 2, buff_dim=_lib.buff_dim)
  for key in ("q", "qp", "qa"):
    scale = max(float(np.abs(want[key]).max()), 1e-30)
    assert np.abs(got[key] - want[key]).max() <= 2e-5 * scale, key
