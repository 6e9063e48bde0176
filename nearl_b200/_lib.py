"""ctypes binding of libnearl_b200.so (include/nearl_b200.h).

There is deliberately no CPU fallback: if the library is missing it is built with nvcc,
and if no CUDA device is usable every compute call raises RuntimeError.
"""
import ctypes
import os

import numpy as np

from . import _build

_lib = None

c_int, c_float, c_double, c_ll, c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_longlong, ctypes.c_void_p

# every symbol include/nearl_b200.h declares (tests check the library exports all of them)
EXPORTED = [
  "nrl_frame_voxelize", "nrl_frame_observation", "nrl_marching_observer", "nrl_density_flow", "nrl_aggregate",
  "nrl_summation", "nrl_frame_voxelize_dev", "nrl_frame_observation_dev", "nrl_marching_observer_dev",
  "nrl_density_flow_dev", "nrl_aggregate_dev", "nrl_featurize_batch", "nrl_buff_dim", "nrl_device_count",
  "nrl_set_device", "nrl_synchronize", "nrl_lane_synchronize", "nrl_lane_fork", "nrl_lane_join", "nrl_focal_means", "nrl_pdb_parse", "nrl_set_stream", "nrl_set_option", "nrl_cell_list_fits", "nrl_superpose", "nrl_profile", "nrl_profile_read", "nrl_profile_name", "nrl_host_register", "nrl_host_unregister", "nrl_kernel_launches",
  "nrl_last_error", "nrl_version",
]

NRL_OK = 0
ERR_OBSERVABLE, ERR_AGGREGATION, ERR_FRAMES, ERR_SHAPE, ERR_ARGUMENT, ERR_NO_DEVICE, ERR_PARSE, ERR_CUDA = 1, 2, 3, 4, 5, 6, 7, 100


class CropStruct(ctypes.Structure):
  _fields_ = [
    ("frame_start", c_int), ("n_frames", c_int),
    ("focal", c_float * 3), ("center", c_float * 3),
    ("half_spacing_f32", c_float), ("half_spacing_f64", c_double),
    ("lo", c_double * 3), ("hi", c_double * 3),
    ("byres", c_int), ("selection", c_int), ("max_atoms", c_int), ("sort_table", c_int),
    ("atom_begin", c_int), ("atom_end", c_int), ("table_shift", c_int),
    ("focal_f64", c_double * 3),
  ]


class TaskStruct(ctypes.Structure):
  _fields_ = [
    ("crop", c_int), ("mode", c_int), ("agg", c_int), ("weight_table", c_int),
    ("dims", c_int * 3), ("spacing", c_float), ("cutoff", c_float), ("sigma", c_float),
    ("rank_table", c_int),
    ("x_begin", c_int), ("x_count", c_int),
  ]


class BatchStruct(ctypes.Structure):
  _fields_ = [
    ("xyz", c_void_p), ("xyz_is_f64", c_int), ("xyz_on_device", c_int), ("n_frames_total", c_int), ("n_atoms", c_int),
    ("weight_tables", c_void_p), ("n_weight_tables", c_int),
    ("selections", c_void_p), ("n_selections", c_int),
    ("residue_index", c_void_p), ("n_residues", c_int),
    ("value_orders", c_void_p),
    ("crops", ctypes.POINTER(CropStruct)), ("n_crops", c_int),
    ("tasks", ctypes.POINTER(TaskStruct)), ("n_tasks", c_int),
    ("out", c_void_p), ("out_on_device", c_int),
    ("atom_counts", c_void_p),
    ("flags", c_int),
    ("lane", c_int),
  ]


_load_lock = __import__("threading").Lock()


def lib():
  """Load (building first if needed) the native library."""
  if _lib is not None:
    return _lib
  with _load_lock:                  # loader threads (io.TrajectoryLoader.prefetched) may be the first to ask
    return _load()


def _load():
  global _lib
  if _lib is not None:
    return _lib
  path = _build.LIB_PATH
  if _build.needs_build():
    try:
      _build.build()
    except Exception as exc:  # a stale prebuilt .so is still better than nothing, a missing one is fatal
      if not os.path.exists(path):
        raise RuntimeError(f"nearl_b200: native library missing and could not be built: {exc}") from exc
  L = ctypes.CDLL(path)
  p = c_void_p
  L.nrl_last_error.restype = ctypes.c_char_p
  L.nrl_version.restype = ctypes.c_char_p
  L.nrl_kernel_launches.restype = c_ll
  L.nrl_frame_voxelize.argtypes = [p, p, c_int, p, c_float, c_float, c_float, c_int, p]
  L.nrl_frame_observation.argtypes = [p, p, c_int, p, c_float, c_float, c_int, p]
  L.nrl_marching_observer.argtypes = [p, p, c_int, c_int, p, c_float, c_float, c_int, c_int, p]
  L.nrl_density_flow.argtypes = [p, p, c_int, c_int, p, c_float, c_float, c_float, c_int, p]
  L.nrl_aggregate.argtypes = [p, c_int, c_ll, c_int, p]
  L.nrl_summation.argtypes = [p, c_ll, p]
  L.nrl_frame_voxelize_dev.argtypes = [p, p, c_int, p, c_float, c_float, c_float, p, p]
  L.nrl_frame_observation_dev.argtypes = [p, p, c_int, p, c_float, c_float, c_int, p, p]
  L.nrl_marching_observer_dev.argtypes = [p, p, c_int, c_int, p, c_float, c_float, c_int, c_int, p, p]
  L.nrl_density_flow_dev.argtypes = [p, p, c_int, c_int, p, c_float, c_float, c_float, c_int, p, p]
  L.nrl_aggregate_dev.argtypes = [p, c_int, c_ll, c_int, p, p]
  L.nrl_featurize_batch.argtypes = [ctypes.POINTER(BatchStruct)]
  L.nrl_buff_dim.argtypes = [c_float, c_float, p]
  L.nrl_device_count.argtypes = [p]
  L.nrl_set_device.argtypes = [c_int]
  L.nrl_host_register.argtypes = [p, c_ll]
  L.nrl_host_unregister.argtypes = [p]
  L.nrl_set_stream.argtypes = [p, c_int]
  L.nrl_lane_synchronize.argtypes = [c_int]
  L.nrl_lane_fork.argtypes = [c_int, p]
  L.nrl_lane_join.argtypes = [c_int, p]
  L.nrl_focal_means.argtypes = [p, c_int, c_int, c_int, p, c_int, p, c_int, p]
  L.nrl_pdb_parse.argtypes = [ctypes.c_char_p, ctypes.c_longlong, p, p, p, p, p]
  L.nrl_profile.argtypes = [c_int]
  L.nrl_set_option.argtypes = [ctypes.c_char_p, c_int]
  L.nrl_cell_list_fits.argtypes = [p, c_float, c_float, p]
  L.nrl_superpose.argtypes = [p, c_int, c_int, c_int, p, c_int, c_int, c_int, p, p]
  L.nrl_profile_read.argtypes = [p, p, c_int]
  L.nrl_profile_name.argtypes = [c_int]
  L.nrl_profile_name.restype = ctypes.c_char_p
  for name in EXPORTED:
    fn = getattr(L, name)
    if fn.restype is c_int and name not in ("nrl_last_error", "nrl_version", "nrl_kernel_launches", "nrl_profile_name"):
      fn.restype = c_int
  _lib = L
  return L


def last_error():
  return lib().nrl_last_error().decode("utf-8", "replace")


def check(rc):
  """Map a status code to the exception the reference's pybind layer would raise
  (/root/reference/src/actions_py.cpp:132-153,207-214) or RuntimeError for CUDA trouble."""
  if rc == NRL_OK:
    return
  msg = last_error()
  if rc in (ERR_OBSERVABLE, ERR_AGGREGATION, ERR_FRAMES, ERR_SHAPE, ERR_ARGUMENT):
    raise ValueError(msg)
  raise RuntimeError(f"nearl_b200 native error {rc}: {msg}")


def ptr(a):
  return a.ctypes.data_as(c_void_p)


def addr(a):
  """Address of an array the CALLER keeps alive for the length of the native call (several times cheaper than `ptr`, whose
  ctypes object also holds a reference to the array -- needed for temporaries only)."""
  return c_void_p(a.__array_interface__["data"][0])


def device_count():
  n = c_int(0)
  check(lib().nrl_device_count(ctypes.byref(n)))
  return n.value


def kernel_launches():
  return int(lib().nrl_kernel_launches())


PROFILE_CATEGORIES = 19


def set_stream(stream_ptr, enable=True):
  check(lib().nrl_set_stream(c_void_p(stream_ptr), 1 if enable else 0))


def cell_list_fits(dims, spacing, cutoff):
  d = (c_int * 3)(*[int(v) for v in dims])
  out = c_int(0)
  check(lib().nrl_cell_list_fits(d, float(spacing), float(cutoff), ctypes.byref(out)))
  return bool(out.value)


OPTION_EPOCH = 0    # bumped by set_option: memoised answers of the library (cell_list_fits) depend on the options


def set_option(name, value):
  global OPTION_EPOCH
  check(lib().nrl_set_option(name.encode(), int(value)))
  OPTION_EPOCH += 1


def profile(enable=True):
  check(lib().nrl_profile(1 if enable else 0))


def profile_read(reset=True):
  """{category name: (total ms, launches)} of the kernels timed since the last reset."""
  ms = (c_double * PROFILE_CATEGORIES)()
  cnt = (c_ll * PROFILE_CATEGORIES)()
  check(lib().nrl_profile_read(ms, cnt, 1 if reset else 0))
  L = lib()
  return {L.nrl_profile_name(i).decode(): (ms[i], int(cnt[i])) for i in range(PROFILE_CATEGORIES) if cnt[i]}


def buff_dim(cutoff, spacing):
  n = c_int(0)
  check(lib().nrl_buff_dim(float(cutoff), float(spacing), ctypes.byref(n)))
  return n.value


def f32c(a):
  return np.ascontiguousarray(a, dtype=np.float32)


def dims3(d):
  d = np.ascontiguousarray(np.asarray(d).reshape(-1), dtype=np.int32)
  if d.size != 3:
    raise ValueError("grid_dims must have 3 entries")
  return d
