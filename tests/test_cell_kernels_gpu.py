"""The cell-list kernels (csrc/cell_kernels.cuh) against the full-frame scan kernels
(csrc/tile_kernels.cuh, forced with nrl_set_option("force_scan", 1)) and against the CPU oracle:
staged cells, direct reads when a frame overflows the staging buffer, several sweep rounds per
frame, NaN coordinates, atoms out of every voxel's reach, ragged grids."""
import numpy as np
import pytest

from nearl_b200 import all_actions as nb
from nearl_b200 import _lib
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def rel_bad(new, ref, rtol=1e-5, atol=0.0):
  ref = np.asarray(ref, dtype=np.float64)
  new = np.asarray(new, dtype=np.float64)
  scale = np.maximum(np.abs(ref), 1e-5 * max(np.abs(ref).max(), 1e-30))
  return int(np.count_nonzero(np.abs(new - ref) > rtol * scale + atol))


@pytest.fixture
def scan_switch():
  yield lambda on: _lib.set_option("force_scan", 1 if on else 0)
  _lib.set_option("force_scan", 0)


@pytest.mark.parametrize("D,s,c,N", [(32, 0.5, 5.0, 1800), (24, 20 / 24, 2.5, 1500), (13, 0.7, 3.1, 400), (40, 0.25, 1.0, 900)])
def test_cell_and_scan_kernels_agree(scan_switch, D, s, c, N):
  rng = np.random.default_rng(21)
  dims = np.array([D, D, D], dtype=np.int32)
  F = 3
  coords = rng.uniform(-2 * c, D * s + 2 * c, size=(F, N, 3)).astype(np.float32)      # some atoms out of every voxel's reach
  coords[0, :5] = 99999.0
  weights = rng.uniform(0.5, 16.0, size=F * N).astype(np.float32)
  weights[::11] = 0.0
  launches = _lib.kernel_launches()
  for obs in (1, 2, 11, 12, 13, 15):
    scan_switch(False)
    got = nb.marching_observer(coords, weights, dims, s, c, obs, 1)
    scan_switch(True)
    ref = nb.marching_observer(coords, weights, dims, s, c, obs, 1)
    if obs in (1, 2):
      assert np.array_equal(got, ref), obs
    else:
      atol = 8 * np.finfo(np.float32).eps * (D * s + c) if obs == 15 else 0.0
      assert rel_bad(got, ref, rtol=2e-5, atol=atol) == 0, (obs, np.abs(got - ref).max())
  for agg in (0, 2):
    scan_switch(False)
    got = nb.density_flow(coords, weights, dims, s, c, 1.5, agg) if agg else nb.frame_voxelize(coords[1], weights[N:2 * N], dims, s, c, 1.5, 0)
    scan_switch(True)
    ref = nb.density_flow(coords, weights, dims, s, c, 1.5, agg) if agg else nb.frame_voxelize(coords[1], weights[N:2 * N], dims, s, c, 1.5, 0)
    assert rel_bad(got, ref, rtol=2e-5, atol=1e-6 * float(np.abs(ref).max())) == 0, (agg, np.abs(got - ref).max())
  assert _lib.kernel_launches() > launches


def test_nan_and_infinite_coordinates():
  """A NaN coordinate makes d2 NaN: `d2 > c2` is false, the observers count the atom for EVERY voxel
  (marching_observers.cu:40); the voxeliser's `d2 < c2` never sees it.  Infinite coordinates are out of range."""
  rng = np.random.default_rng(22)
  dims = np.array([16, 16, 16], dtype=np.int32)
  N, s, c = 300, 0.5, 3.0
  coords = rng.uniform(-c, 8.0 + c, size=(N, 3)).astype(np.float32)
  coords[7, 1] = np.nan
  coords[8] = np.nan
  coords[9, 0] = np.inf
  coords[10, 2] = -np.inf
  weights = rng.integers(1, 6, size=N).astype(np.float32)
  for obs in (1, 2, 12):
    got = nb.frame_observation(coords, weights, dims, s, c, obs)
    ref = orc.frame_observation(coords, weights, dims, s, c, obs)
    assert np.array_equal(got, ref), obs
  assert nb.frame_observation(coords, weights, dims, s, c, 2).min() >= 2.0
  got = nb.frame_voxelize(coords, weights, dims, s, c, 1.0, 0)
  ref = orc.frame_voxelize(coords, weights, dims, s, c, 1.0, buff_dim=_lib.buff_dim(c, s))
  assert np.isfinite(got).all() and rel_bad(got, ref) == 0


def test_crowded_cells_and_many_rounds():
  """5 000 atoms in a 6 A cube: every frame overflows the staging buffer (direct reads), every warp
  sweeps its list in several rounds, and single cells hold more atoms than the insertion sort takes."""
  rng = np.random.default_rng(23)
  dims = np.array([12, 12, 12], dtype=np.int32)
  N, s, c = 5000, 0.5, 4.0
  coords = rng.uniform(0.0, 6.0, size=(2, N, 3)).astype(np.float32)
  coords[0, :200] = coords[0, 0]                                 # 200 atoms on one point
  weights = rng.integers(1, 5, size=2 * N).astype(np.float32)
  for obs in (1, 2):
    got = nb.marching_observer(coords, weights, dims, s, c, obs, 5)
    ref = orc.marching_observer(coords, weights, dims, s, c, obs, 5)
    assert np.array_equal(got, ref), obs
  got = nb.frame_observation(coords[0], weights[:N], dims, s, c, 11)
  ref = orc.frame_observation(coords[0], weights[:N], dims, s, c, 11)
  assert rel_bad(got, ref) == 0
  got = nb.frame_voxelize(coords[0], weights[:N], dims, s, c, 2.0, 0)
  ref = orc.frame_voxelize(coords[0], weights[:N], dims, s, c, 2.0, buff_dim=_lib.buff_dim(c, s))
  assert rel_bad(got, ref) == 0
  # deterministic: the same call twice gives the same bits
  again = nb.frame_voxelize(coords[0], weights[:N], dims, s, c, 2.0, 0)
  assert np.array_equal(got, again)


def test_large_grid_falls_back_to_scan_when_cells_do_not_fit():
  """A grid whose cell list would need more counters than the binning kernel holds keeps the scan kernels."""
  rng = np.random.default_rng(24)
  dims = np.array([200, 200, 200], dtype=np.int32)
  N, s, c = 50, 0.5, 0.6
  coords = rng.uniform(0.0, 100.0, size=(N, 3)).astype(np.float32)
  weights = np.ones(N, dtype=np.float32)
  got = nb.frame_observation(coords, weights, dims, s, c, 2)
  assert got.size == 200 ** 3
  # every atom is seen by the lattice points within 0.6 A of it: 7 on average, at least 1
  assert 1 * N <= got.sum() <= 33 * N
