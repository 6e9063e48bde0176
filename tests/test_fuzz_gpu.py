"""Seeded random configurations of the whole path -- grid shape and spacing, cutoff, sigma, padding, window, selections,
by-residue crops, frame offsets, every observable and reducer, focal points on atoms / in empty space / at the edge of the
system, float32 and float64 trajectories -- through Featurizer.run (device crop, cell lists or full-frame scan, fused
sweeps, fused reducers) against oracle/host_ref.py feeding the C oracle.  Each case is small enough for the oracle to
finish in about a second; the seeds are fixed, so a failure names a reproducible case."""
import os

import numpy as np
import pytest

import nearl_b200
from nearl_b200 import _lib, features, featurizer, io
from oracle import host_ref

pytestmark = pytest.mark.gpu
nearl_b200.update_config(quiet=True)

OBS = ["existence", "direct_count", "distinct_count", "mean_distance", "cumulative_weight", "density", "dispersion",
       "eccentricity", "radius_of_gyration"]
AGG = ["mean", "standard_deviation", "median", "variance", "max", "min", "information_entropy", "drift"]
WEIGHTS = ["mass", "atomic_number", "electronegativity", "hydrophobicity", "heavy_atom", "uniformed", "residue_id", "sidechainness"]
SELECTIONS = [None, None, "!@H=", ":ARG,LYS,GLU", ":LIG", "!:LIG", "@CA,C,N,O"]
INTEGER_OBS = (1, 2, 3)


def random_case(seed):
  rng = np.random.default_rng(seed)
  n_atoms = int(rng.integers(300, 2200))
  window = int(rng.choice([1, 2, 3, 5, 8, 17, 24]))           # > 16 frames: the series leaves shared memory
  n_frames = window * int(rng.integers(1, 4 if window < 17 else 2)) + int(rng.integers(0, 2))
  dtype = np.float64 if rng.random() < 0.25 else np.float32
  traj = io.synthetic_trajectory(n_atoms, n_frames, seed=seed * 7 + 1, dtype=dtype, pin=False, ligand_atoms=int(rng.choice([0, 12, 30])),
                                 step=float(rng.choice([0.1, 0.3, 0.8])))
  dims = [int(v) for v in rng.integers(6, 21, size=3)] if rng.random() < 0.5 else [int(rng.integers(6, 21))] * 3
  if rng.random() < 0.1:
    dims = [int(rng.integers(24, 41))] * 3                    # several tiles per axis
  spacing = float(rng.choice([0.35, 0.5, 0.75, 1.0, 1.3]))
  lengths = [d * spacing for d in dims]
  cutoff = float(rng.choice([1.5, 2.5, 3.25, 4.0, 5.0]))
  sigma = float(rng.choice([0.7, 1.0, 1.5, 2.2]))
  parms = {"dimensions": dims, "lengths": lengths, "time_window": window, "cutoff": cutoff, "sigma": sigma}

  def kw():
    out = {}
    sel = SELECTIONS[int(rng.integers(0, len(SELECTIONS)))]
    if sel is not None:
      out["selection"] = sel
    if rng.random() < 0.2:
      out["byres"] = True
    if rng.random() < 0.2:
      out["padding"] = float(rng.choice([0.0, 1.0, cutoff + 1.5]))
    if rng.random() < 0.2:
      out["cutoff"] = float(rng.choice([2.0, 3.0, 4.5]))
    return out

  specs = []
  for i in range(int(rng.integers(3, 8))):
    kind = rng.random()
    k = kw()
    if kind < 0.3:
      cls = [features.Mass, features.HeavyAtom, features.AtomicNumber, features.Electronegativity, features.Hydrophobicity][int(rng.integers(0, 5))]
      if window > 1 and rng.random() < 0.3:
        k["frame_offset"] = int(rng.integers(0, window))
      if rng.random() < 0.2:
        k["sigma"] = float(rng.choice([0.8, 1.8]))
      specs.append((cls, dict(k, outkey=f"f{i}")))
    elif kind < 0.75:
      k.update(obs=OBS[int(rng.integers(0, len(OBS)))], agg=AGG[int(rng.integers(0, len(AGG)))],
               weight_type=WEIGHTS[int(rng.integers(0, len(WEIGHTS)))], outkey=f"f{i}")
      specs.append((features.MarchingObservers, k))
    else:
      k.update(agg=AGG[int(rng.integers(0, len(AGG)))], weight_type=WEIGHTS[int(rng.integers(0, len(WEIGHTS)))], outkey=f"f{i}")
      specs.append((features.DensityFlow, k))
  # focal points: a mask, an atom index list, or absolute points (one of them far outside the system)
  choice = rng.random()
  side = (n_atoms / 0.1) ** (1.0 / 3.0)
  if choice < 0.4:
    focus, fmt = [":ARG", "@CA"][: int(rng.integers(1, 3))], "mask"
  elif choice < 0.7:
    focus, fmt = [rng.choice(n_atoms, size=int(rng.integers(1, 6)), replace=False) for _ in range(int(rng.integers(1, 3)))], "index"
  else:
    pts = rng.uniform(0.0, side, size=(int(rng.integers(1, 3)), 3))
    if rng.random() < 0.5:
      pts = np.concatenate([pts, [[side + 40.0, -25.0, side / 2]]])
    focus, fmt = pts.astype(np.float32), "absolute"
  return traj, parms, specs, focus, fmt


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


SEEDS = list(range(1, 1 + int(os.environ.get("NRL_FUZZ_CASES", "48"))))      # NRL_FUZZ_CASES=600 for a soak run


def reference_module():
  """The reference's own kernels (oracle/_ref, GPU box), or None."""
  import sys
  ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
  if not os.path.isdir(ref_dir):
    return None
  if ref_dir not in sys.path:
    sys.path.insert(0, ref_dir)
  try:
    import all_actions
    return all_actions
  except Exception:
    return None


@pytest.mark.parametrize("seed", SEEDS)
def test_random_configuration_against_the_oracle(seed):
  traj, parms, specs, focus, fmt = random_case(seed)
  fz = featurizer.Featurizer(parms)
  fz.register_features([cls(**k) for cls, k in specs])
  fz.register_focus(focus, fmt)
  fz.register_trajloader([traj])
  fz.keep_results = True
  if seed % 5 == 0:
    _lib.set_option("force_scan", 1)             # every fifth case on the full-frame scan kernels
  if seed % 2 == 1:
    _lib.set_option("crop_stream_min", 1)        # every other case: the streaming crop whatever the size of the system
  try:
    fz.run()
  finally:
    _lib.set_option("force_scan", 0)
    _lib.set_option("crop_stream_min", 0)
  got = {k: np.array(v) for k, v in fz.RESULTS.items()}
  ref = host_ref.featurize(traj.xyz, fz.FEATURESPACE[0].resids, descs_from(fz), fz.FOCALPOINTS, fz.time_window, buff_dim=_lib.buff_dim)
  for f in fz.FEATURESPACE:
    key, where = f.outkey, (seed, f.outkey, str(f))
    g, r = got[key].reshape(ref[key].shape), ref[key]
    obs, agg = getattr(f, "obs", 0) if isinstance(f, features.MarchingObservers) else 0, getattr(f, "agg", 0) if isinstance(f, features.DynamicFeature) else 0
    scale = max(float(np.abs(r).max()), 1e-30)
    if obs in INTEGER_OBS and agg in (3, 5, 6):
      if agg == 3 and fz.time_window % 2 == 0:
        assert np.allclose(g, r, rtol=3e-7, atol=0), where                  # even window: the median halves a sum (div.approx)
      else:
        assert np.array_equal(g, r), where                                  # integers, order statistics: bit-exact
    elif agg == 7:
      # The entropy is discontinuous in its samples (a histogram), so entropy tasks over float observables run on the scan
      # kernel, which adds the atoms in index order like the reference and is bit-identical to it.  Observer entropies are
      # therefore compared with the REFERENCE'S OWN KERNELS, on every voxel (the IEEE oracle cannot follow the reference's
      # fast-math division where a sample sits exactly on a bin edge -- the rule for integer samples and sums of discrete
      # weights); the entropy of a density flow (the reference deposits with float atomics: its own samples vary) with the
      # oracle by the reference's own criterion, a few voxels in a thousand off (test_features.py:119-125).
      ref_mod = reference_module() if obs != 0 else None          # obs == 0: density flows, compared with the oracle below
      if ref_mod is not None:
        # samples that are integers or sums of discrete weights sit exactly on bin edges, where the reference's fast-math
        # division decides and the IEEE oracle cannot follow: the arbiter for these is the reference's own kernels
        d1 = [d for d in descs_from(fz) if d["outkey"] == key]
        rk = host_ref.featurize(traj.xyz, fz.FEATURESPACE[0].resids, d1, fz.FOCALPOINTS, fz.time_window,
                                kernels=host_ref.ModuleKernels(ref_mod))[key]
        assert np.count_nonzero(~np.isclose(g, rk, rtol=1e-5, atol=1e-6)) == 0, (where, "vs the reference's own kernels")
        continue
      bad = np.count_nonzero(~np.isclose(g, r, atol=1e-3))
      assert bad <= max(2, r.size // 200), (where, bad)
    elif agg in (2, 4, 8) or obs in (14, 15, 16):
      # reducers that cancel (std, variance, drift) and observables that subtract positions (dispersion, eccentricity,
      # radius of gyration): error relative to the magnitude of what is subtracted
      tol = 2e-5 * scale + (8 * np.finfo(np.float32).eps * (max(fz.FEATURESPACE[0].lengths) + f.cutoff) if obs in (14, 15, 16) else 0.0)
      if agg in (2, 4, 8):
        # the magnitude of the SERIES, not of the (cancelled) result, sets the size of the rounding error: the same task reduced
        # with max gives it
        d5 = [dict(d, agg=5) for d in descs_from(fz) if d["outkey"] == key]
        peak = float(np.abs(host_ref.featurize(traj.xyz, fz.FEATURESPACE[0].resids, d5, fz.FOCALPOINTS, fz.time_window,
                                               buff_dim=_lib.buff_dim)[key]).max())
        tol += 2e-5 * (peak * peak if agg == 4 else peak)
      assert np.count_nonzero(np.abs(g.astype(np.float64) - r) > 1e-5 * np.abs(r) + tol) == 0, (where, np.abs(g - r).max())
    else:
      bad = np.abs(g.astype(np.float64) - r) > 1e-5 * np.maximum(np.abs(r), 1e-5 * scale)
      assert bad.sum() == 0, (where, int(bad.sum()), float(np.abs(g - r).max()))


@pytest.mark.parametrize("seed", list(range(1, 13)))
def test_random_single_frame_systems_grouped(seed):
  """Static channels over a list of single-frame systems of different sizes: the grouped path (several systems per native
  call, each a slice of one pseudo-system with its own focal points and residue numbering) against the oracle, system by
  system."""
  rng = np.random.default_rng(1000 + seed)
  systems = [io.synthetic_complex(int(rng.integers(200, 1500)), seed=seed * 31 + i, ligand_atoms=int(rng.choice([8, 30])),
                                  dtype=(np.float64 if seed % 4 == 0 else np.float32)) for i in range(int(rng.integers(3, 8)))]
  dims = [int(v) for v in rng.integers(6, 19, size=3)]
  spacing = float(rng.choice([0.5, 0.8, 1.0]))
  parms = {"dimensions": dims, "lengths": [d * spacing for d in dims], "time_window": 1,
           "cutoff": float(rng.choice([1.5, 2.5, 4.0])), "sigma": float(rng.choice([0.8, 1.5]))}
  classes = [features.Mass, features.HeavyAtom, features.AtomicNumber, features.Electronegativity, features.Hydrophobicity]
  specs = []
  for i in range(int(rng.integers(2, 7))):
    k = {"outkey": f"s{i}"}
    sel = SELECTIONS[int(rng.integers(0, len(SELECTIONS)))]
    if sel is not None:
      k["selection"] = sel
    if rng.random() < 0.3:
      k["byres"] = True
    if rng.random() < 0.3:
      k["cutoff"] = float(rng.choice([2.0, 3.0]))
    if rng.random() < 0.2:
      k["padding"] = float(rng.choice([0.0, 5.0]))
    specs.append((classes[int(rng.integers(0, len(classes)))], k))
  specs.append((features.AtomType, {"focus_element": int(rng.choice([1, 6, 7, 8])), "outkey": "type"}))
  fz = featurizer.Featurizer(parms)
  fz.register_features([cls(**k) for cls, k in specs])
  fz.register_focus([":LIG", ":ARG"][: int(rng.integers(1, 3))], "mask")
  fz.register_trajloader(systems)
  fz.static_group = int(rng.choice([2, 3, 32]))
  _lib.set_option("crop_stream_min", 1 if seed % 2 else 0)      # odd seeds: member slices (unaligned starts) on the streaming crop
  rows = {}
  fz.result_sink = lambda outkey, tid, bid, pid, grid: rows.setdefault((tid, outkey), []).append(np.array(grid))
  try:
    fz.run()
  finally:
    _lib.set_option("crop_stream_min", 0)
  for tid, traj in enumerate(systems):
    one = featurizer.Featurizer(parms)
    one.register_features([cls(**k) for cls, k in specs])
    one.register_focus(fz.FOCALPOINTS_PROTOTYPE, "mask")
    one.traj = traj
    assert one.parse_focus()
    for f in one.FEATURESPACE:
      f.cache(traj)
    ref = host_ref.featurize(traj.xyz, one.FEATURESPACE[0].resids, descs_from(one), one.FOCALPOINTS, 1, buff_dim=_lib.buff_dim)
    for f in one.FEATURESPACE:
      g, r = np.array(rows[(tid, f.outkey)]).reshape(ref[f.outkey].shape), ref[f.outkey]
      scale = max(float(np.abs(r).max()), 1e-30)
      bad = np.abs(g.astype(np.float64) - r) > 1e-5 * np.maximum(np.abs(r), 1e-5 * scale)
      assert bad.sum() == 0, (seed, tid, f.outkey, int(bad.sum()), float(np.abs(g - r).max()))


@pytest.mark.parametrize("seed", list(range(1, 1 + max(40, int(os.environ.get("NRL_FUZZ_CASES", "40"))))))
def test_random_calls_against_the_reference_kernels(seed):
  """The six FFI twins on random shapes against the REFERENCE'S OWN KERNELS (oracle/_ref, built unmodified from the
  reference's sources): the arbiter that shares the reference's fast-math arithmetic, which the IEEE oracle cannot."""
  ref = reference_module()
  if ref is None:
    pytest.skip("oracle/_ref not built")
  from nearl_b200 import all_actions as nb
  rng = np.random.default_rng(5000 + seed)
  D = int(rng.integers(8, 25))
  dims = np.array([D, D, D], dtype=np.int32)                  # the reference's index math needs cubic grids
  s = float(rng.choice([0.5, 0.75, 1.0]))
  c = float(rng.choice([1.5, 2.5, 3.5, 5.0]))
  sig = float(rng.choice([0.8, 1.5, 2.0]))
  F = int(rng.choice([1, 2, 5, 10, 17]))
  N = int(rng.integers(1, 700))
  coords = rng.uniform(-c, D * s + c, size=(F, N, 3)).astype(np.float32)
  if rng.random() < 0.3:
    coords[:, N - N // 5:] = 99999.0                          # the padding of DynamicFeature.query
  integer_w = rng.random() < 0.5
  w = (rng.integers(0, 9, size=F * N) if integer_w else rng.uniform(0.0, 16.0, size=F * N)).astype(np.float32)
  if rng.random() < 0.3:
    w[rng.random(F * N) < 0.2] = 0.0
  obs = int(rng.choice([1, 2, 3, 11, 12, 13, 14, 15, 16]))
  agg = int(rng.integers(1, 9))
  where = (seed, D, s, c, F, N, obs, agg, integer_w)
  g = nb.marching_observer(coords, w, dims, s, c, obs, agg)
  r = ref.marching_observer(coords, w, np.array(dims, dtype=int), s, c, obs, agg)
  scale = max(float(np.abs(r).max()), 1e-30)
  if obs in INTEGER_OBS and agg in (5, 6):
    assert np.array_equal(g, r), where
  elif agg == 7:
    # entropy tasks run where the arithmetic is the reference's to the last bit (index-ordered sums): every voxel
    bad = np.count_nonzero(~np.isclose(g, r, rtol=1e-5, atol=1e-5))
    assert bad == 0, (where, bad)
  elif agg in (2, 4, 8) or obs in (14, 15, 16):
    tol = 2e-5 * scale + (8 * np.finfo(np.float32).eps * (D * s + c) if obs in (14, 15, 16) else 0.0)
    assert np.count_nonzero(np.abs(g.astype(np.float64) - r) > 1e-5 * np.abs(r) + tol) == 0, (where, float(np.abs(g - r).max()))
  else:
    assert np.count_nonzero(np.abs(g.astype(np.float64) - r) > 1e-5 * np.maximum(np.abs(r), 1e-5 * scale)) == 0, (where, float(np.abs(g - r).max()))
  # single frame, voxeliser and flow
  g1 = nb.frame_observation(coords[0], w[:N], dims, s, c, obs)
  r1 = ref.frame_observation(coords[0], w[:N], np.array(dims, dtype=int), s, c, obs)
  tol1 = 8 * np.finfo(np.float32).eps * (D * s + c) if obs in (14, 15, 16) else 0.0
  assert np.count_nonzero(np.abs(g1.astype(np.float64) - r1) > 1e-5 * np.abs(r1) + tol1) == 0, (where, "frame_observation")
  gv = nb.frame_voxelize(coords[0], w[:N], dims, s, c, sig, 0)
  rv = ref.frame_voxelize(coords[0], w[:N], np.array(dims, dtype=int), s, c, sig, 0)
  sv = max(float(np.abs(rv).max()), 1e-30)
  assert np.count_nonzero(np.abs(gv.astype(np.float64) - rv) > 1e-5 * np.maximum(np.abs(rv), 1e-5 * sv)) == 0, (where, "frame_voxelize")
  if agg in (1, 2, 5, 6):
    gf = nb.density_flow(coords, w, dims, s, c, sig, agg)
    rf = ref.density_flow(coords, w, np.array(dims, dtype=int), s, c, sig, agg)
    sf = max(float(np.abs(rf).max()), 1e-30)
    assert np.count_nonzero(np.abs(gf.astype(np.float64) - rf) > 1e-5 * np.abs(rf) + (2e-5 * sf if agg == 2 else 1e-10 * sf)) == 0, (where, "density_flow")


@pytest.mark.parametrize("seed", list(range(1, 1 + max(24, int(os.environ.get("NRL_FUZZ_CASES", "24"))))))
def test_random_series_against_the_reference_aggregate(seed):
  """Standalone K3 on random series shapes (1..512 frames, ragged voxel counts, integer / float / constant / negative series)
  against the reference's own `aggregate`: streaming, register, pipelined (tiled TMA) and one-shot staged kernels."""
  ref = reference_module()
  if ref is None:
    pytest.skip("oracle/_ref not built")
  from nearl_b200 import all_actions as nb
  rng = np.random.default_rng(9000 + seed)
  F = int(rng.choice([1, 2, 3, 7, 10, 33, 64, 65, 100, 129, 257, 300, 512]))
  G = int(rng.choice([1, 5, 64, 1000, 1003, 4096, 4 * 3001, 20000]))
  kind = rng.random()
  if kind < 0.35:
    x = rng.integers(-3, 12, size=(F, G)).astype(np.float32)
  elif kind < 0.8:
    x = (rng.random((F, G)) * float(rng.choice([1.0, 50.0, 1e-3]))).astype(np.float32)
  else:
    x = (rng.standard_normal((F, G)) * 4).astype(np.float32)
  if G > 4:
    x[:, ::5] = float(rng.choice([0.0, 1.25, -2.0]))            # constant columns
  with_nan = seed % 6 == 0 and F > 1
  if with_nan:                                                    # a NaN compares false with everything: the order statistics
    x[rng.integers(0, F, 40), rng.integers(0, G, 40)] = np.nan    # then depend on the reference's sort itself
  for agg in range(1, 9):
    if with_nan and agg not in (3, 5, 6):
      continue
    g = nb.aggregate(x, agg)
    r = ref.aggregate(x, agg)
    where = (seed, F, G, agg, round(kind, 2))
    if agg in (3, 5, 6):
      assert np.array_equal(g, r, equal_nan=True), where
    else:
      scale = max(float(np.abs(x).max()), 1e-30)
      assert np.count_nonzero(np.abs(g.astype(np.float64) - r) > 1e-5 * np.abs(r) + 1e-6 * scale) == 0, (where, float(np.abs(g - r).max()))
