"""ctypes front-end of the CPU oracle (oracle/nearl_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of nearl_oracle.c.  Nothing under
``nearl_b200/`` imports this module; it is used by ``tests/``, by
``__graft_entry__.smoke()`` and by the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py``.

The functions mirror the six ``all_actions`` entry points of the reference
(/root/reference/src/actions_py.cpp:297-351): numpy in, flat float32 numpy out.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnearl_oracle.so")
_lib = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
  """Compile the C restatement with the Makefile next to it (gcc only)."""
  if force or not os.path.exists(_LIB_PATH) or \
      os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "nearl_oracle.c")):
    subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
  return _LIB_PATH


def lib():
  global _lib
  if _lib is not None:
    return _lib
  build()
  try:
    L = ctypes.CDLL(_LIB_PATH)
  except OSError:
    build(force=True)
    L = ctypes.CDLL(_LIB_PATH)
  c_int, c_long, c_float = ctypes.c_int, ctypes.c_long, ctypes.c_float
  L.orc_buff_dim.restype = c_int
  L.orc_buff_dim.argtypes = [c_float, c_float]
  L.orc_frame_voxelize.restype = None
  L.orc_frame_voxelize.argtypes = [_f32p, _f32p, c_int, _i32p, c_float, c_float, c_float, c_int, _f32p]
  L.orc_observe_frame.restype = None
  L.orc_observe_frame.argtypes = [_f32p, _f32p, c_int, _i32p, c_float, c_float, c_int, _f32p]
  L.orc_marching_observer.restype = None
  L.orc_marching_observer.argtypes = [_f32p, _f32p, c_int, c_int, _i32p, c_float, c_float, c_int, c_int, _f32p]
  L.orc_marching_observer_frames.restype = None
  L.orc_marching_observer_frames.argtypes = [_f32p, _f32p, c_int, c_int, _i32p, c_float, c_float, c_int, _f32p]
  L.orc_density_flow.restype = None
  L.orc_density_flow.argtypes = [_f32p, _f32p, c_int, c_int, _i32p, c_float, c_float, c_float, c_int, c_int, _f32p]
  L.orc_aggregate.restype = None
  L.orc_aggregate.argtypes = [_f32p, c_int, c_long, c_int, _f32p]
  L.orc_summation.restype = c_float
  L.orc_summation.argtypes = [_f32p, c_long]
  _lib = L
  return L


def _f32(a):
  return np.ascontiguousarray(a, dtype=np.float32)


def _dims(d):
  d = np.ascontiguousarray(np.asarray(d).reshape(-1)[:3], dtype=np.int32)
  assert d.size == 3
  return d


def buff_dim(cutoff, spacing):
  return int(lib().orc_buff_dim(float(cutoff), float(spacing)))


def frame_voxelize(coords, weights, grid_dims, spacing, cutoff, sigma, buff_dim=-1):
  coords, weights, dims = _f32(coords).reshape(-1, 3), _f32(weights).reshape(-1), _dims(grid_dims)
  assert coords.shape[0] == weights.shape[0]
  out = np.zeros(int(np.prod(dims)), dtype=np.float32)
  lib().orc_frame_voxelize(coords, weights, coords.shape[0], dims, spacing, cutoff, sigma, int(buff_dim), out)
  return out


def frame_observation(coords, weights, grid_dims, spacing, cutoff, type_obs):
  coords, weights, dims = _f32(coords).reshape(-1, 3), _f32(weights).reshape(-1), _dims(grid_dims)
  out = np.zeros(int(np.prod(dims)), dtype=np.float32)
  lib().orc_observe_frame(coords, weights, coords.shape[0], dims, spacing, cutoff, int(type_obs), out)
  return out


def marching_observer(coords, weights, grid_dims, spacing, cutoff, type_obs, type_agg):
  coords, weights, dims = _f32(coords), _f32(weights).reshape(-1), _dims(grid_dims)
  assert coords.ndim == 3
  F, N = coords.shape[0], coords.shape[1]
  assert weights.size >= F * N
  out = np.zeros(int(np.prod(dims)), dtype=np.float32)
  lib().orc_marching_observer(coords, weights, F, N, dims, spacing, cutoff, int(type_obs), int(type_agg), out)
  return out


def marching_observer_frames(coords, weights, grid_dims, spacing, cutoff, type_obs):
  """Per-frame [F, G] observations before aggregation."""
  coords, weights, dims = _f32(coords), _f32(weights).reshape(-1), _dims(grid_dims)
  F, N = coords.shape[0], coords.shape[1]
  out = np.zeros((F, int(np.prod(dims))), dtype=np.float32)
  lib().orc_marching_observer_frames(coords, weights, F, N, dims, spacing, cutoff, int(type_obs), out)
  return out


def density_flow(traj, weights, grid_dims, spacing, cutoff, sigma, type_agg, buff_dim=-1):
  traj, weights, dims = _f32(traj), _f32(weights).reshape(-1), _dims(grid_dims)
  assert traj.ndim == 3
  F, N = traj.shape[0], traj.shape[1]
  out = np.zeros(int(np.prod(dims)), dtype=np.float32)
  lib().orc_density_flow(traj, weights, F, N, dims, spacing, cutoff, sigma, int(type_agg), int(buff_dim), out)
  return out


def aggregate(arr, type_agg):
  arr = _f32(arr)
  assert arr.ndim == 2
  out = np.zeros(arr.shape[1], dtype=np.float32)
  lib().orc_aggregate(arr, arr.shape[0], arr.shape[1], int(type_agg), out)
  return out


def summation(arr):
  arr = _f32(arr).reshape(-1)
  return float(lib().orc_summation(arr, arr.size))
