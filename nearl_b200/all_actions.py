"""Drop-in for the reference's ``nearl.all_actions`` pybind11 module
(/root/reference/src/actions_py.cpp:297-351): the same six functions, the same argument
names and order, numpy in -> flat float32 numpy out, ``ValueError`` where the reference
raises ``py::value_error`` -- computed by the sm_100a kernels behind the C ABI of
``include/nearl_b200.h``.  No CPU path exists: without a CUDA device every call raises.
"""
import sys

import numpy as np

from . import _lib
from ._lib import check, dims3, f32c, ptr

__all__ = ["frame_voxelize", "frame_observation", "marching_observer", "density_flow", "aggregate", "summation"]


def frame_voxelize(coords, weights, grid_dims, spacing, cutoff, sigma, auto_translate):
  """actions_py.cpp:45-89.  ``auto_translate`` recentres ``coords`` in place when it is a
  C-contiguous float32 array (the reference mutates the caller's buffer)."""
  c = f32c(coords)
  w = f32c(weights)
  dims = dims3(grid_dims)
  if c.shape[0] != w.shape[0]:
    # actions_py.cpp:58-61: message on stderr and an empty array, not an exception
    print("Input arrays must have the same length", file=sys.stderr)
    return np.zeros((0,), dtype=np.float32)
  out = np.empty(int(dims[0]) * int(dims[1]) * int(dims[2]), dtype=np.float32)
  check(_lib.lib().nrl_frame_voxelize(ptr(c), ptr(w), c.shape[0], ptr(dims), spacing, cutoff, sigma,
                                      int(auto_translate), ptr(out)))
  return out


def frame_observation(coords, weights, dims, spacing, cutoff, type_obs):
  """actions_py.cpp:268-294."""
  c = f32c(coords)
  w = f32c(weights)
  d = dims3(dims)
  out = np.empty(int(d[0]) * int(d[1]) * int(d[2]), dtype=np.float32)
  check(_lib.lib().nrl_frame_observation(ptr(c), ptr(w), c.shape[0], ptr(d), spacing, cutoff, int(type_obs), ptr(out)))
  return out


def marching_observer(coords, weights, dims, spacing, cutoff, type_obs, type_agg):
  """actions_py.cpp:112-166."""
  c = f32c(coords)
  w = f32c(weights)
  d = dims3(dims)
  if c.ndim != 3:
    raise ValueError("Error: The input array must have 3 dimensions: (frame_nr, atom_nr, 3)")
  if w.size < c.shape[0] * c.shape[1]:
    raise ValueError("weights must hold frame_nr * atom_nr values")   # the reference reads out of bounds here
  out = np.empty(int(d[0]) * int(d[1]) * int(d[2]), dtype=np.float32)
  check(_lib.lib().nrl_marching_observer(ptr(c), ptr(w), c.shape[0], c.shape[1], ptr(d), spacing, cutoff,
                                         int(type_obs), int(type_agg), ptr(out)))
  return out


def density_flow(traj, weights, grid_dims, spacing, cutoff, sigma, type_agg):
  """actions_py.cpp:188-229."""
  c = f32c(traj)
  w = f32c(weights)
  d = dims3(grid_dims)
  if c.ndim != 3:
    raise ValueError("Error: The input array must have 3 dimensions: (frame_nr, atom_nr, 3)")
  if w.size < c.shape[0] * c.shape[1]:
    raise ValueError("weights must hold frame_nr * atom_nr values")
  out = np.empty(int(d[0]) * int(d[1]) * int(d[2]), dtype=np.float32)
  check(_lib.lib().nrl_density_flow(ptr(c), ptr(w), c.shape[0], c.shape[1], ptr(d), spacing, cutoff, sigma,
                                    int(type_agg), ptr(out)))
  return out


def aggregate(arr, type_agg):
  """actions_py.cpp:232-252: [frame_nr, gridpoint_nr] -> [gridpoint_nr]."""
  a = f32c(arr)
  if a.ndim != 2:
    raise ValueError("arr must be (frame_nr, gridpoint_nr)")
  out = np.empty(a.shape[1], dtype=np.float32)
  check(_lib.lib().nrl_aggregate(ptr(a), a.shape[0], a.shape[1], int(type_agg), ptr(out)))
  return out


def summation(arr):
  """actions_py.cpp:254-266."""
  a = f32c(arr).reshape(-1)
  out = np.zeros(1, dtype=np.float32)
  check(_lib.lib().nrl_summation(ptr(a), a.size, ptr(out)))
  return float(out[0])
