"""HDF5 helpers that define the output layout (/root/reference/nearl/utils.py:491-570):
``append_hdf_data`` (one extensible dataset per feature key, rows appended along axis 0) and
``dump_dict`` (a group of scalar / 1-D datasets, recreated on every call).

Back-end: ``h5py`` when it is importable, otherwise the self-contained writer/reader in
``nearl_b200.hdf5_lite`` (this image has neither h5py nor libhdf5); both produce files that
stock h5py opens with the layout of the reference.  While a ``batched_output()`` context is
active (the featuriser opens one per trajectory) appends are buffered per (file, key) and
written with one open/close per file, instead of the reference's open/close per result.
"""
import contextlib

import numpy as np

try:
  import h5py as _h5py
except Exception:          # pragma: no cover - depends on the image
  _h5py = None

from . import hdf5_lite

__all__ = ["append_hdf_data", "dump_dict", "batched_output", "open_hdf", "hdf_backend"]

_pending = None     # {(file, key): {"rows": [...], "dtype":, "maxshape":, "kwargs":}} while batching


def hdf_backend():
  return "h5py" if _h5py is not None else "hdf5_lite"


def open_hdf(path, mode="r"):
  """File object with the subset of the h5py API this package uses."""
  if _h5py is not None:
    return _h5py.File(path, mode)
  return hdf5_lite.File(path, mode)


def _append_now(hdf, key, data, dtype, maxshape, **kwargs):
  if key in hdf:
    dset = hdf[key]
    old = dset.shape[0]
    dset.resize((old + data.shape[0], *dset.shape[1:]))
    dset[old:old + data.shape[0]] = data
  else:
    dset = hdf.create_dataset(key, data.shape, dtype=dtype, maxshape=maxshape, **kwargs)
    dset[:] = data


def append_hdf_data(hdffile, key, data, dtype, maxshape, **kwargs):
  """Append `data` (leading axis = rows) to dataset `key`, creating it on first use with the
  given maxshape / chunks / compression (utils.py:491-521)."""
  data = np.asarray(data)
  if _pending is not None:
    slot = _pending.setdefault((hdffile, key), {"rows": [], "dtype": dtype, "maxshape": maxshape, "kwargs": kwargs})
    slot["rows"].append(data)
    return
  with open_hdf(hdffile, "a") as hdf:
    _append_now(hdf, key, data, dtype, maxshape, **kwargs)


@contextlib.contextmanager
def batched_output():
  """Buffer append_hdf_data calls and write them file by file on exit (row order preserved)."""
  global _pending
  if _pending is not None:      # nested: the outer context flushes
    yield
    return
  _pending = {}
  try:
    yield
  finally:
    pending, _pending = _pending, None
    by_file = {}
    for (path, key), slot in pending.items():
      by_file.setdefault(path, []).append((key, slot))
    for path, items in by_file.items():
      with open_hdf(path, "a") as hdf:
        for key, slot in items:
          _append_now(hdf, key, np.concatenate(slot["rows"], axis=0), slot["dtype"], slot["maxshape"], **slot["kwargs"])


def dump_dict(outfile, groupname, dic):
  """(Re)create group `groupname` holding one dataset per entry of `dic` with h5py's native
  type mapping: float -> f64, int -> i64, str -> variable-length UTF-8, sequences -> arrays
  (utils.py:562-570)."""
  with open_hdf(outfile, "a") as hdf:
    if groupname in hdf:
      del hdf[groupname]
    grp = hdf.create_group(groupname)
    for key, val in dic.items():
      grp[key] = val
