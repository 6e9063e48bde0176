"""GPU parity of the six C-ABI entry points against the CPU oracle (bit-exact for the integer
observables, <= 1e-5 relative for float features) and against the reference's own kernels
(oracle/_ref).  Mirrors /root/reference/pytests/test_actions.py and test_features.py."""
import numpy as np
import pytest

from nearl_b200 import all_actions as nb
from nearl_b200 import _lib
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # north_star: <= 1e-5 relative error on float features


def rel_close(new, ref, rtol=RTOL, atol=0.0):
  """|new - ref| <= rtol * max(|ref|, 1e-5 * max|ref|) + atol  (SURVEY.md appendix B)."""
  ref = np.asarray(ref, dtype=np.float64)
  new = np.asarray(new, dtype=np.float64)
  scale = np.maximum(np.abs(ref), 1e-5 * max(np.abs(ref).max(), 1e-30))
  bad = np.abs(new - ref) > rtol * scale + atol
  return bad


def synth(rng, F, N, box, c):
  coords = rng.uniform(-c, box + c, size=(F, N, 3)).astype(np.float32)
  return coords


DIMS = np.array([32, 32, 32], dtype=np.int32)


def test_reference_known_answer_mass_conservation():
  # /root/reference/pytests/test_actions.py:18-31
  np.random.seed(0)
  coords = np.random.normal(size=(100, 3), loc=5, scale=1).astype(np.float32)
  weights = np.full((100,), 16.0, dtype=np.float32)
  ret = nb.frame_voxelize(coords, weights, DIMS, 0.5, 0.5, 1.0, 0)
  assert np.isclose(np.sum(ret), np.sum(weights))
  assert not np.isnan(ret).any()
  # /root/reference/pytests/test_benchmark.py:19-29
  np.random.seed(0)
  coords = np.random.normal(size=(300, 3), loc=5, scale=1).astype(np.float32)
  weights = np.full((300,), 1.5, dtype=np.float32)
  ret = nb.frame_voxelize(coords, weights, DIMS, 1.0, 2.5, 1.0, 0)
  assert np.isclose(np.sum(ret), np.sum(weights), rtol=1e-3)


@pytest.mark.parametrize("D,s,c,sig,N", [(32, 0.5, 5.0, 1.5, 1750), (24, 20 / 24, 2.5, 1.5, 1560),
                                         (32, 1.0, 2.5, 1.0, 300), (20, 0.625, 2.5, 1.5, 500), (9, 0.7, 3.1, 0.9, 77)])
def test_frame_voxelize_vs_oracle(D, s, c, sig, N):
  rng = np.random.default_rng(1)
  dims = np.array([D, D, D], dtype=np.int32)
  coords = synth(rng, 1, N, D * s, c)[0]
  weights = rng.uniform(0.5, 16, size=N).astype(np.float32)
  weights[::7] = 0.0
  coords[3] = 99999.0
  got = nb.frame_voxelize(coords, weights, dims, s, c, sig, 0)
  b = _lib.buff_dim(c, s)
  ref = orc.frame_voxelize(coords, weights, dims, s, c, sig, buff_dim=b)
  bad = rel_close(got, ref)
  assert bad.sum() == 0, (bad.sum(), np.abs(got - ref).max())


@pytest.mark.parametrize("obs", [1, 2, 3])
def test_integer_observers_bit_exact(obs):
  rng = np.random.default_rng(2)
  F, N, s, c = 4, 900, 0.5, 5.0
  coords = synth(rng, F, N, 16.0, c)
  coords[1, 5:9] = 99999.0
  weights = rng.integers(1, 9, size=F * N).astype(np.float32)
  weights[10] = 99999.0
  weights[11] = -0.0
  weights[12] = 0.0
  for f in range(F):
    got = nb.frame_observation(coords[f], weights[f * N:(f + 1) * N], DIMS, s, c, obs)
    ref = orc.frame_observation(coords[f], weights[f * N:(f + 1) * N], DIMS, s, c, obs)
    assert np.array_equal(got, ref), (obs, f, np.count_nonzero(got != ref))
  for agg in (1, 2, 3, 5, 6, 7, 8):
    got = nb.marching_observer(coords, weights, DIMS, s, c, obs, agg)
    ref = orc.marching_observer(coords, weights, DIMS, s, c, obs, agg)
    if agg in (3, 5, 6):
      assert np.array_equal(got, ref), (obs, agg)
    elif agg == 7:
      assert np.count_nonzero(~np.isclose(got, ref, atol=1e-3)) < 10
    else:
      assert np.allclose(got, ref, rtol=1e-5, atol=1e-6), (obs, agg, np.abs(got - ref).max())


@pytest.mark.parametrize("obs", [11, 12, 13, 15, 16])
def test_float_observers_vs_oracle(obs):
  rng = np.random.default_rng(3)
  F, N, s, c = 3, 700, 0.5, 4.0
  coords = synth(rng, F, N, 16.0, c)
  weights = rng.uniform(1.0, 16.0, size=F * N).astype(np.float32)
  # eccentricity is |COM - observer|, a difference of two positions of magnitude <= L + c: its
  # absolute error floor is a few ulp of the coordinates, whatever the size of the result
  atol = 8 * np.finfo(np.float32).eps * (16.0 + c) if obs == 15 else 0.0
  for f in range(F):
    got = nb.frame_observation(coords[f], weights[f * N:(f + 1) * N], DIMS, s, c, obs)
    ref = orc.frame_observation(coords[f], weights[f * N:(f + 1) * N], DIMS, s, c, obs)
    bad = rel_close(got, ref, atol=atol)
    assert bad.sum() == 0, (obs, f, bad.sum(), np.abs(got - ref).max())
  got = nb.marching_observer(coords, weights, DIMS, s, c, obs, 1)
  ref = orc.marching_observer(coords, weights, DIMS, s, c, obs, 1)
  assert rel_close(got, ref, atol=atol).sum() == 0


def test_dispersion_vs_oracle():
  rng = np.random.default_rng(4)
  dims = np.array([12, 12, 12], dtype=np.int32)
  N, s, c = 160, 1.0, 3.5
  coords = synth(rng, 1, N, 12.0, c)[0]
  weights = rng.uniform(1.0, 4.0, size=N).astype(np.float32)
  got = nb.frame_observation(coords, weights, dims, s, c, 14)
  ref = orc.frame_observation(coords, weights, dims, s, c, 14)
  assert rel_close(got, ref).sum() == 0


@pytest.mark.parametrize("agg", [1, 2, 4, 8])
def test_density_flow_vs_oracle(agg):
  rng = np.random.default_rng(5)
  F, N, s, c, sig = 5, 600, 0.5, 4.0, 1.5
  coords = synth(rng, F, N, 16.0, c)
  coords[2, 590:] = 99999.0
  weights = rng.uniform(1.0, 16.0, size=F * N).astype(np.float32)
  weights[2 * N + 590:3 * N] = 0.0
  got = nb.density_flow(coords, weights, DIMS, s, c, sig, agg)
  ref = orc.density_flow(coords, weights, DIMS, s, c, sig, agg, buff_dim=_lib.buff_dim(c, s))
  # std / variance / drift subtract nearly equal per-frame values: their error is relative to the
  # magnitude of the series (here: the largest per-frame density), not to the reduced value
  atol = 0.0
  if agg in (2, 4, 8):
    peak = float(orc.density_flow(coords, weights, DIMS, s, c, sig, 5, buff_dim=_lib.buff_dim(c, s)).max())
    atol = RTOL * (peak * peak if agg == 4 else peak)
  bad = rel_close(got, ref, atol=atol)
  assert bad.sum() == 0, (agg, bad.sum(), np.abs(got - ref).max())


def test_aggregate_known_answers():
  # /root/reference/pytests/test_features.py:86-126 (same seed, same shapes, same tolerances)
  np.random.seed(0)
  x = np.random.rand(50, 32 * 32 * 32)
  for agg in range(1, 9):
    got = nb.aggregate(x, agg)
    ref = orc.aggregate(x, agg)
    assert np.isclose(np.sum(ref), np.sum(got), rtol=1e-3, atol=1e-3)
    assert np.isclose(np.min(ref), np.min(got), rtol=1e-3, atol=1e-3)
    assert np.isclose(np.max(ref), np.max(got), rtol=1e-3, atol=1e-3)
    assert np.count_nonzero(~np.isclose(ref, got, atol=1e-3)) < 10
    if agg in (3, 5, 6):
      assert np.array_equal(got, ref)
    elif agg != 7:
      assert np.allclose(got, ref, rtol=1e-5, atol=1e-6), (agg, np.abs(got - ref).max())


def test_aggregate_ragged_and_unknown():
  rng = np.random.default_rng(6)
  x = rng.random((7, 1003)).astype(np.float32)      # G not a multiple of 4: non-bulk path
  for agg in (1, 3, 7):
    assert np.allclose(nb.aggregate(x, agg), orc.aggregate(x, agg), rtol=1e-5, atol=1e-6)
  assert np.array_equal(nb.aggregate(x, 99), np.zeros(1003, dtype=np.float32))   # gpuutils.cu:73-76
  x1 = rng.random((1, 64)).astype(np.float32)
  for agg in range(1, 9):
    assert np.allclose(nb.aggregate(x1, agg), orc.aggregate(x1, agg), rtol=1e-6, atol=0)


@pytest.mark.parametrize("F,G", [(65, 1000), (100, 4 * 6007), (257, 4 * 1001), (300, 2048), (512, 1028), (100, 1003)])
def test_aggregate_long_series_every_dispatch_branch(F, G):
  # > 64 frames: std / var / entropy go through the pipelined kernel (tiled TMA copies, partial last tile, two row
  # boxes beyond 256 frames, an odd row count), the median through the one-shot staged kernel, the recurrences
  # through the streaming kernel; G = 1003 is not a multiple of 4 and takes the non-bulk path
  rng = np.random.default_rng(F * 7 + G)
  x = (rng.random((F, G)) * 3.0).astype(np.float32)
  x[:, ::5] = 1.25                        # constant series: zero variance, entropy of a zero range
  x[F // 2, 3::11] = -7.5                 # one outlier frame
  for agg in range(1, 9):
    got = nb.aggregate(x, agg)
    ref = orc.aggregate(x, agg)
    if agg in (3, 5, 6):
      assert np.array_equal(got, ref), agg
    elif agg == 7:
      assert np.count_nonzero(~np.isclose(got, ref, atol=1e-3)) <= max(1, G // 3000), agg     # reference criterion (test_features.py:119-125)
    else:
      assert np.allclose(got, ref, rtol=1e-5, atol=2e-6), (agg, np.abs(got - ref).max())
  # the pipelined kernel gives the bits of the one-shot staged kernel
  for agg in (2, 4, 7):
    a = nb.aggregate(x, agg)
    _lib.set_option("agg_pipe_stages", -1)
    try:
      b = nb.aggregate(x, agg)
    finally:
      _lib.set_option("agg_pipe_stages", 0)
    assert np.array_equal(a.view(np.int32), b.view(np.int32)), agg


def test_summation():
  rng = np.random.default_rng(7)
  a = rng.random(100003).astype(np.float32)
  assert nb.summation(a) == orc.summation(a)


def test_errors_match_reference_binding():
  coords = np.zeros((3, 5, 3), dtype=np.float32)
  w = np.ones(15, dtype=np.float32)
  with pytest.raises(ValueError, match="observable type is not supported"):
    nb.marching_observer(coords, w, DIMS, 0.5, 2.0, 4, 1)
  with pytest.raises(ValueError, match="aggregation type is not supported"):
    nb.marching_observer(coords, w, DIMS, 0.5, 2.0, 1, 9)
  with pytest.raises(ValueError, match="aggregation type is not supported"):
    nb.density_flow(coords, w, DIMS, 0.5, 2.0, 1.0, 0)
  big = np.zeros((513, 1, 3), dtype=np.float32)
  with pytest.raises(ValueError, match="exceeds the maximum number of frames allowed 512"):
    nb.marching_observer(big, np.ones(513, dtype=np.float32), DIMS, 0.5, 2.0, 1, 1)
  assert nb.frame_voxelize(coords[0], np.ones(4, dtype=np.float32), DIMS, 0.5, 2.0, 1.0, 0).shape == (0,)


def test_empty_crop_semantics():
  # N = 0 -> observables 0, eccentricity = cutoff everywhere
  coords = np.zeros((2, 0, 3), dtype=np.float32)
  w = np.zeros(0, dtype=np.float32)
  dims = np.array([8, 8, 8], dtype=np.int32)
  assert np.all(nb.marching_observer(coords, w, dims, 1.0, 2.5, 2, 1) == 0)
  assert np.all(nb.marching_observer(coords, w, dims, 1.0, 2.5, 15, 1) == np.float32(2.5))


def test_noncubic_follows_reference_index_math():
  rng = np.random.default_rng(8)
  dims = np.array([10, 12, 14], dtype=np.int32)
  coords = rng.uniform(-2, 14, size=(200, 3)).astype(np.float32)
  w = rng.uniform(1, 3, size=200).astype(np.float32)
  assert np.array_equal(nb.frame_observation(coords, w, dims, 1.0, 2.5, 2), orc.frame_observation(coords, w, dims, 1.0, 2.5, 2))
  got = nb.frame_voxelize(coords, w, dims, 1.0, 2.5, 1.0, 0)
  ref = orc.frame_voxelize(coords, w, dims, 1.0, 2.5, 1.0, buff_dim=_lib.buff_dim(2.5, 1.0))
  assert rel_close(got, ref).sum() == 0


# ---- against the reference's own kernels on the same GPU --------------------------------------
def test_vs_reference_kernels(ref_actions):
  rng = np.random.default_rng(9)
  F, N, s, c, sig = 4, 800, 0.5, 5.0, 1.5
  coords = synth(rng, F, N, 16.0, c)
  weights = rng.integers(1, 9, size=F * N).astype(np.float32)
  ref = ref_actions.frame_voxelize(coords[0], weights[:N], DIMS, s, c, sig, 0)
  if not np.any(ref):
    pytest.skip("reference kernels returned zeros (no usable GPU for the reference build)")
  got = nb.frame_voxelize(coords[0], weights[:N], DIMS, s, c, sig, 0)
  assert rel_close(got, ref).sum() == 0, np.abs(got - ref).max()
  for obs in (1, 2, 3):
    for agg in (1, 2, 3, 5, 6, 8):
      r = ref_actions.marching_observer(coords, weights, DIMS, s, c, obs, agg)
      g = nb.marching_observer(coords, weights, DIMS, s, c, obs, agg)
      if agg in (3, 5, 6):
        assert np.array_equal(g, r), (obs, agg)
      else:
        assert np.allclose(g, r, rtol=1e-6, atol=1e-7), (obs, agg, np.abs(g - r).max())
  for obs in (11, 12, 13, 15, 16):
    r = ref_actions.marching_observer(coords, weights, DIMS, s, c, obs, 1)
    g = nb.marching_observer(coords, weights, DIMS, s, c, obs, 1)
    atol = 8 * np.finfo(np.float32).eps * (16.0 + c) if obs == 15 else 0.0
    assert rel_close(g, r, atol=atol).sum() == 0, (obs, np.abs(g - r).max())
  r = ref_actions.density_flow(coords, weights, DIMS, s, c, sig, 2)
  g = nb.density_flow(coords, weights, DIMS, s, c, sig, 2)
  peak = float(ref_actions.density_flow(coords, weights, DIMS, s, c, sig, 5).max())
  assert rel_close(g, r, atol=RTOL * peak).sum() == 0      # std cancels: error relative to the series magnitude
  r = ref_actions.density_flow(coords, weights, DIMS, s, c, sig, 1)
  g = nb.density_flow(coords, weights, DIMS, s, c, sig, 1)
  assert rel_close(g, r).sum() == 0
  x = rng.random((50, 32768)).astype(np.float32)
  for agg in range(1, 9):
    r = ref_actions.aggregate(x, agg)
    g = nb.aggregate(x, agg)
    assert np.count_nonzero(~np.isclose(g, r, rtol=1e-5, atol=1e-6)) == 0, agg
  # integer-valued series (counts, sums of integer weights) put many samples EXACTLY on a histogram bin edge, where the
  # reference's fast-math division decides the bin: the entropy must follow the reference's kernels there, short
  # series (register kernel), long series (pipelined kernel) and the one-shot staged kernel alike
  for F in (9, 50, 100, 130):
    xi = rng.integers(0, 9, size=(F, 8192)).astype(np.float32)
    xi[:, ::3] = rng.integers(0, 17, size=(F, xi[:, ::3].shape[1])).astype(np.float32)
    r = ref_actions.aggregate(xi, 7)
    g = nb.aggregate(xi, 7)
    assert np.count_nonzero(~np.isclose(g, r, rtol=1e-5, atol=1e-6)) == 0, (F, np.count_nonzero(~np.isclose(g, r, rtol=1e-5, atol=1e-6)))
    _lib.set_option("agg_pipe_stages", -1)
    try:
      g1 = nb.aggregate(xi, 7)
    finally:
      _lib.set_option("agg_pipe_stages", 0)
    assert np.array_equal(g.view(np.int32), g1.view(np.int32)), F


def test_buff_dim_probe(ref_actions):
  """Records how the device evaluates (cutoff+spacing)/spacing."""
  import json, os
  rows = []
  for c, s in [(5, 0.5), (2.5, 0.625), (2.5, 20 / 24), (5, 0.75), (2.5, 20 / 48), (3.5, 16 / 24), (2.55, 0.5), (6, 30 / 32)]:
    rows.append({"cutoff": c, "spacing": s, "device": _lib.buff_dim(c, s), "ieee": orc.buff_dim(c, s)})
  os.makedirs("gpurun_out", exist_ok=True)
  with open("gpurun_out/buff_dim_probe.json", "w") as fh:
    json.dump(rows, fh, indent=1)
  assert all(abs(r["device"] - r["ieee"]) <= 1 for r in rows)


# ---- overflow paths of the tile kernels ------------------------------------------------------------
def test_long_window_series_in_global_scratch():
  """40 frames x 512 voxels x 4 B = 80 KB per tile: the series leaves shared memory."""
  rng = np.random.default_rng(10)
  dims = np.array([16, 16, 16], dtype=np.int32)
  F, N, s, c = 40, 300, 0.5, 3.0
  coords = synth(rng, F, N, 8.0, c)
  weights = rng.integers(1, 5, size=F * N).astype(np.float32)
  for obs, agg in [(2, 1), (2, 3), (3, 5), (1, 7), (12, 2), (2, 8)]:
    got = nb.marching_observer(coords, weights, dims, s, c, obs, agg)
    ref = orc.marching_observer(coords, weights, dims, s, c, obs, agg)
    if agg in (3, 5):
      assert np.array_equal(got, ref), (obs, agg)
    elif agg == 7:
      assert np.count_nonzero(~np.isclose(got, ref, atol=1e-3)) < 3
    else:
      assert np.allclose(got, ref, rtol=1e-5, atol=1e-5), (obs, agg, np.abs(got - ref).max())
  got = nb.density_flow(coords, weights, dims, s, c, 1.0, 1)
  ref = orc.density_flow(coords, weights, dims, s, c, 1.0, 1, buff_dim=_lib.buff_dim(c, s))
  assert rel_close(got, ref).sum() == 0


def test_dense_system_candidates_in_global_scratch():
  """More than 512 atoms within reach of one tile: the candidate list leaves shared memory."""
  rng = np.random.default_rng(11)
  dims = np.array([12, 12, 12], dtype=np.int32)
  N, s, c = 3000, 0.5, 6.0
  coords = rng.uniform(-2.0, 8.0, size=(2, N, 3)).astype(np.float32)
  weights = rng.integers(1, 40, size=2 * N).astype(np.float32)
  for obs in (2, 3):
    got = nb.marching_observer(coords, weights, dims, s, c, obs, 5)
    ref = orc.marching_observer(coords, weights, dims, s, c, obs, 5)
    assert got.max() > 512 or obs == 3
    assert np.array_equal(got, ref), obs
  got = nb.frame_voxelize(coords[0], weights[:N], dims, s, c, 2.0, 0)
  ref = orc.frame_voxelize(coords[0], weights[:N], dims, s, c, 2.0, buff_dim=_lib.buff_dim(c, s))
  assert rel_close(got, ref).sum() == 0
  got = nb.frame_observation(coords[0], weights[:N], dims, s, c, 11)
  ref = orc.frame_observation(coords[0], weights[:N], dims, s, c, 11)
  assert rel_close(got, ref).sum() == 0


def test_max_frames_boundary():
  rng = np.random.default_rng(12)
  dims = np.array([8, 8, 8], dtype=np.int32)
  coords = rng.uniform(0, 8, size=(512, 20, 3)).astype(np.float32)
  w = np.ones(512 * 20, dtype=np.float32)
  got = nb.marching_observer(coords, w, dims, 1.0, 2.0, 2, 1)          # exactly MAX_FRAME_NUMBER frames is allowed
  ref = orc.marching_observer(coords, w, dims, 1.0, 2.0, 2, 1)
  assert np.allclose(got, ref, rtol=1e-6, atol=1e-6)
  big = np.concatenate([coords, coords[:8]], axis=0)                   # density_flow: frames beyond 512 are ignored
  wb = np.ones(520 * 20, dtype=np.float32)
  assert np.array_equal(nb.density_flow(big, wb, dims, 1.0, 2.0, 1.0, 5), nb.density_flow(coords, w, dims, 1.0, 2.0, 1.0, 5))


def test_single_frame_observables_and_reducers_bit_identical_to_the_reference(ref_actions):
  # On the index-ordered path (frame_observation: the scan kernel adds the atoms in the reference's order) the arithmetic is
  # the reference binary's op for op: all nine observables, and all eight reducers of `aggregate`, to the last bit
  # (scripts/bitwise_vs_reference.py prints the rates).  The voxeliser is excluded: the reference deposits with float
  # atomics, its own result varies from run to run.
  rng = np.random.default_rng(77)
  D, s, c = 20, 0.75, 3.5
  dims = np.array([D, D, D], dtype=np.int32)
  for N in (300, 2500):
    coords = rng.uniform(-c, D * s + c, size=(N, 3)).astype(np.float32)
    w = rng.uniform(0.5, 16.0, size=N).astype(np.float32)
    for obs in (1, 2, 3, 11, 12, 13, 14, 15, 16):
      g = nb.frame_observation(coords, w, dims, s, c, obs)
      r = ref_actions.frame_observation(coords, w, np.array(dims, dtype=int), s, c, obs)
      assert np.array_equal(g.view(np.int32), r.view(np.int32)), (N, obs, int(np.count_nonzero(g.view(np.int32) != r.view(np.int32))))
  coords = rng.uniform(-c, D * s + c, size=(10, 600, 3)).astype(np.float32)
  w = rng.uniform(0.5, 16.0, size=10 * 600).astype(np.float32)
  series = np.stack([ref_actions.frame_observation(coords[i], w[i * 600:(i + 1) * 600], np.array(dims, dtype=int), s, c, 11) for i in range(10)])
  for agg in range(1, 9):
    g = nb.aggregate(series, agg)
    r = ref_actions.aggregate(series, agg)
    assert np.array_equal(g.view(np.int32), r.view(np.int32)), agg


@pytest.mark.gpu
def test_dev_twins_on_different_streams_share_the_workspace_in_order():
  """The stream-ordered `_dev` twins return without a sync and use one workspace per device: a second call on ANOTHER
  stream must not repack atoms / tasks while the first call's kernels still read them.  A long call (64 frames, 48^3) on
  stream A is followed at once by short calls on streams B and C and by a host entry on the library's own stream; every
  result equals the one computed alone."""
  import ctypes
  import torch
  from nearl_b200 import _lib
  L = _lib.lib()
  rng = np.random.default_rng(77)
  void = ctypes.c_void_p

  def case(F, N, D, obs, agg):
    coords = (rng.random((F, N, 3)) * (D * 0.5 + 4) - 2).astype(np.float32)
    weights = rng.integers(1, 9, size=(F, N)).astype(np.float32)
    dims = np.array([D, D, D], dtype=np.int32)
    want = np.empty(D ** 3, dtype=np.float32)
    _lib.check(L.nrl_marching_observer(_lib.ptr(coords), _lib.ptr(weights), F, N, _lib.ptr(dims), 0.5, 4.0, obs, agg, _lib.ptr(want)))
    return coords, weights, dims, want, (F, N, obs, agg)

  cases = [case(64, 6000, 48, 11, 2), case(3, 400, 16, 2, 1), case(5, 700, 16, 12, 5), case(2, 300, 16, 3, 1)]
  streams = [torch.cuda.Stream() for _ in cases]
  dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(w).cuda(), torch.empty(len(want), device="cuda")) for c, w, _, want, _ in cases]
  torch.cuda.synchronize()
  for rounds in range(3):
    for (c, w, out), (_, _, dims, _, (F, N, obs, agg)), st in zip(dev, cases, streams):
      out.zero_()
    torch.cuda.synchronize()
    for (c, w, out), (_, _, dims, _, (F, N, obs, agg)), st in zip(dev, cases, streams):
      _lib.check(L.nrl_marching_observer_dev(void(c.data_ptr()), void(w.data_ptr()), F, N, _lib.ptr(dims), 0.5, 4.0, obs, agg,
                                             void(out.data_ptr()), void(st.cuda_stream)))
    # a host entry right behind them, on the library's own stream
    coords, weights, dims, want, (F, N, obs, agg) = cases[1]
    again = np.empty_like(want)
    _lib.check(L.nrl_marching_observer(_lib.ptr(coords), _lib.ptr(weights), F, N, _lib.ptr(dims), 0.5, 4.0, obs, agg, _lib.ptr(again)))
    assert np.array_equal(again, want)
    torch.cuda.synchronize()
    for (c, w, out), (_, _, _, want, key) in zip(dev, cases):
      assert np.array_equal(out.cpu().numpy(), want), (rounds, key)
