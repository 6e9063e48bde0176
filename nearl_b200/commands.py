"""``nearl.commands`` twin (/root/reference/nearl/commands.py:23-190): dtype coercion, int
dims, reshape of the flat native result to ``grid_dims``.  Same names and argument order;
``frame_observation`` keeps the reference's Python signature (with the unused ``sigma``) but
calls the native function with the arity it really has (the reference passes one argument
too many, commands.py:105 vs actions_py.cpp:268-275)."""
import numpy as np

from . import all_actions, printit

__all__ = ["frame_voxelize", "frame_observation", "density_flow", "marching_observer"]


def _f32(a):
  a = np.asarray(a)
  return a if a.dtype == np.float32 else a.astype(np.float32)


def frame_voxelize(coords, weights, grid_dims, spacing, cutoff, sigma):
  grid_dims = np.array(grid_dims, dtype=int)
  ret = all_actions.frame_voxelize(_f32(coords), _f32(weights), grid_dims, float(spacing), float(cutoff), float(sigma), 0)
  return ret.reshape(grid_dims)


def frame_observation(coords, weights, grid_dims, spacing, cutoff, sigma, type_obs):
  grid_dims = np.array(grid_dims, dtype=int)
  ret = all_actions.frame_observation(_f32(coords), _f32(weights), grid_dims, float(spacing), float(cutoff), int(type_obs))
  return ret.reshape(grid_dims)


def marching_observer(coords, weights, grid_dims, spacing, cutoff, type_obs, type_agg):
  grid_dims = np.asarray(grid_dims, dtype=int)
  ret = all_actions.marching_observer(_f32(coords), _f32(weights), grid_dims, spacing, cutoff, type_obs, type_agg)
  return ret.reshape(grid_dims)


def density_flow(traj, weights, grid_dims, spacing, cutoff, sigma, type_agg):
  grid_dims = np.array(grid_dims, dtype=int)
  ret = all_actions.density_flow(_f32(traj), _f32(weights), grid_dims, float(spacing), float(cutoff), float(sigma), int(type_agg))
  if np.isnan(ret).any():
    printit(f"Warning: Found nan in the return: {np.count_nonzero(np.isnan(ret))}")
  return ret.reshape(grid_dims)
