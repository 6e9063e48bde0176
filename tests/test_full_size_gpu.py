"""BASELINE.json configs[2..4] at their FULL per-item sizes on the GPU, checked through properties that do not
need the (brute-force, minutes-long) oracle at that size, plus an oracle spot check on a sample:
  configs[2]  PDBBind 24^3 x 8 channels: linearity of the voxeliser across channels
  configs[3]  MISATO 32^3 mixed static + dynamic: a density flow is the reducer of per-frame static grids
  configs[4]  128^3 x 200 000 atoms, window 100: a corner of the big grid equals a small grid placed there"""
import numpy as np
import pytest

import nearl_b200
from nearl_b200 import _lib, constants, features, featurizer, io, workloads
from oracle import host_ref

pytestmark = pytest.mark.gpu
nearl_b200.update_config(quiet=True)


def rel_bad(new, ref, rtol=1e-5, atol=0.0):
  ref = np.asarray(ref, dtype=np.float64); new = np.asarray(new, dtype=np.float64)
  scale = np.maximum(np.abs(ref), 1e-5 * max(np.abs(ref).max(), 1e-30))
  return int((np.abs(new - ref) > rtol * scale + atol).sum())


def test_pdbbind_channels_are_linear_in_the_weights():
  fz, trajs = workloads.pdbbind_static(n_complex=24, pin=False)
  fz.keep_results = True
  stats = fz.run()
  assert stats["native_tasks"] == 24 * 8
  R = {k: np.array(v, dtype=np.float64) for k, v in fz.RESULTS.items()}
  assert R["mass"].shape == (24, 24, 24, 24)
  # HeavyAtom = C + N + O + S indicator channels (the synthetic systems hold H, C, N, O, S only)
  heavy = R["type_6"] + R["type_7"] + R["type_8"] + R["type_16"]
  assert np.allclose(R["heavy"], heavy, rtol=2e-6, atol=1e-7 * heavy.max())
  mass = sum(constants.ATOMICMASS[z] * R[f"type_{z}"] for z in (1, 6, 7, 8, 16))
  assert np.allclose(R["mass"], mass, rtol=2e-6, atol=1e-7 * mass.max())
  assert (R["ligand"] <= R["heavy"] + R["type_1"] + 1e-6).all() and R["ligand"].sum() > 0
  # oracle spot check: the first two complexes, all 8 channels
  fz.traj = trajs[0]
  for i in range(2):
    fz.traj = trajs[i]
    fz.parse_focus()
    for f in fz.FEATURESPACE:
      f.cache(trajs[i])
    ref = host_ref.featurize(trajs[i].xyz, fz.FEATURESPACE[0].resids, workloads.describe(fz), fz.FOCALPOINTS, 1, buff_dim=_lib.buff_dim)
    for key, want in ref.items():
      assert rel_bad(R[key][i], want[0]) == 0, (i, key)


def test_misato_flow_is_the_reducer_of_static_frames():
  # configs[3] shape: 6 000 atoms x 100 frames, 32^3 over 20 A, window 10; cutoff 5.0 for the pair under test
  traj = io.synthetic_trajectory(6000, 100, seed=workloads.BASE_SEED + 4, pin=False)
  fz = featurizer.Featurizer({"dimensions": 32, "lengths": 20, "time_window": 10, "sigma": 1.5, "cutoff": 5.0})
  feats = [features.Mass(outkey=f"mass_{k}", frame_offset=k) for k in range(10)]
  feats += [features.DensityFlow(agg="mean", weight_type="mass", outkey="flow_mean"),
            features.DensityFlow(agg="max", weight_type="mass", outkey="flow_max"),
            features.MarchingObservers(obs="existence", agg="mean", weight_type="mass", outkey="exist"),
            features.MarchingObservers(obs="direct_count", agg="min", weight_type="mass", outkey="count_min")]
  fz.register_features(feats)
  fz.register_focus([":ARG"], "mask")
  fz.register_trajloader([traj])
  fz.keep_results = True
  fz.run()
  R = {k: np.array(v) for k, v in fz.RESULTS.items()}
  frames = np.stack([R[f"mass_{k}"] for k in range(10)], axis=1)           # [window, frame, x, y, z]
  s = np.zeros_like(frames[:, 0])
  for k in range(10):
    s = (s + frames[:, k]).astype(np.float32)                                # sum_device: sequential float32 adds
  want = s / np.float32(10)
  rel = np.abs(R["flow_mean"] - want) / np.maximum(np.abs(want), 1e-30)
  assert rel.max() <= 1e-6, rel.max()                                        # div.approx vs IEEE division: a few ulp
  assert np.array_equal(R["flow_max"], frames.max(axis=1))
  assert R["exist"].min() >= 0.0 and R["exist"].max() <= 1.0 and np.allclose(R["exist"] * 10, np.round(R["exist"] * 10), atol=1e-5)
  assert (R["count_min"] == np.round(R["count_min"])).all() and ((R["count_min"] > 0) <= (R["exist"] == 1.0)).all()


def test_stress_corner_of_the_big_grid_equals_a_small_grid():
  # configs[4]: 200 000 atoms, 128^3 at 0.5 A, window 100 (one window here), the reference's 9 999-atom cap lifted so
  # that both grids see the same atoms
  traj = io.synthetic_trajectory(200000, 100, seed=workloads.BASE_SEED + 5, pin=False)
  focal_big = np.array([[63.0, 63.0, 63.0]], dtype=np.float32)

  def run(dim, focal):
    fz = featurizer.Featurizer({"dimensions": dim, "lengths": dim * 0.5, "time_window": 100, "cutoff": 5.0, "sigma": 1.5})
    feats = [features.MarchingObservers(obs="existence", agg="mean", weight_type="mass", outkey="exist"),
             features.MarchingObservers(obs="direct_count", agg="max", weight_type="mass", outkey="count"),
             features.MarchingObservers(obs="distinct_count", agg="mean", weight_type="atomic_number", outkey="distinct"),
             features.MarchingObservers(obs="mean_distance", agg="standard_deviation", weight_type="mass", outkey="mean_distance"),
             features.DensityFlow(agg="mean", weight_type="mass", outkey="flow")]
    for f in feats:
      f.MAX_ALLOWED_ATOMS = 0
    fz.register_features(feats)
    fz.register_focus(focal, "absolute")
    fz.register_trajloader([traj])
    fz.keep_results = True
    stats = fz.run()
    assert stats["native_tasks"] == 5
    return {k: np.array(v)[0] for k, v in fz.RESULTS.items()}

  big = run(128, focal_big)
  assert big["exist"].shape == (128, 128, 128)
  small = run(32, focal_big - np.float32(24.0))          # (128 - 32) / 2 voxels * 0.5 A: the same lattice points
  corner = {k: v[:32, :32, :32] for k, v in big.items()}
  # translated coordinates differ in the last bit between the two boxes: over 100 frames a few atoms within an ulp
  # of some voxel's cutoff flip in or out (32 768 voxels x 100 frames x ~50 atoms in range each)
  for key in ("exist", "count", "distinct"):
    assert np.count_nonzero(corner[key] != small[key]) <= 32, key
  assert rel_bad(corner["mean_distance"], small["mean_distance"], rtol=1e-4, atol=1e-4 * float(small["mean_distance"].max())) <= 64
  # the voxeliser normalises over the BUFFERED lattice of its own grid: equal only away from the small grid's inner faces
  assert rel_bad(corner["flow"][:22, :22, :22], small["flow"][:22, :22, :22], rtol=2e-5, atol=1e-6 * float(small["flow"].max())) == 0
  assert big["exist"].min() >= 0 and big["exist"].max() <= 1 and big["count"].max() < 200
