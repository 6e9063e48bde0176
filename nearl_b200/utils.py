"""HDF5 helpers that define the output layout (/root/reference/nearl/utils.py:491-570):
``append_hdf_data`` (one extensible dataset per feature key, rows appended along axis 0) and
``dump_dict`` (a group of scalar / 1-D datasets, recreated on every call).

Back-end: ``h5py`` when it is importable, otherwise the self-contained writer/reader in
``nearl_b200.hdf5_lite`` (this image has neither h5py nor libhdf5); both produce files with the layout of
the reference.  The reference opens and closes the file once per result; here

  * inside ``hdf_session()`` (the featuriser holds one for the length of ``run()`` / ``loop_by_residue()``) a file
    opened for appending stays open until the session ends -- every ``open_hdf(path, "a")`` hands out the same
    handle -- so its metadata is read once per run, not once per trajectory;
  * inside ``batched_output()`` (one per trajectory) appends are buffered per (file, key) and written with one
    resize per dataset; the file's metadata is brought up to date when the context exits, so that what is on
    disk after every trajectory is a complete, readable file.
"""
import contextlib
import sys

import numpy as np

try:
  import h5py as _h5py
except Exception:          # pragma: no cover - depends on the image
  _h5py = None

from . import constants, hdf5_lite

__all__ = ["append_hdf_data", "update_hdf_data", "shared_reader", "dump_dict", "batched_output", "hdf_session", "open_hdf", "hdf_backend", "find_block_single",
           "find_block_dual", "dual_blocks_by_key"]

_pending = None     # {(file, key): {"rows": [...], "dtype":, "maxshape":, "kwargs":}} while batching
_session = None     # {path: open file} while an hdf_session is active


def hdf_backend():
  return "h5py" if _h5py is not None else "hdf5_lite"


class _Borrowed:
  """A file of the active session, handed out by open_hdf: leaving its `with` block (or close()) only flushes."""

  def __init__(self, hdf):
    self._hdf = hdf

  def __enter__(self):
    return self._hdf

  def __exit__(self, *exc):
    self._hdf.flush()
    return False

  def close(self):
    self._hdf.flush()

  def __getattr__(self, name):
    return getattr(self._hdf, name)

  def __contains__(self, key):
    return key in self._hdf

  def __getitem__(self, key):
    return self._hdf[key]

  def __setitem__(self, key, value):
    self._hdf[key] = value

  def __delitem__(self, key):
    del self._hdf[key]


def _open(path, mode):
  return _h5py.File(path, mode) if _h5py is not None else hdf5_lite.File(path, mode)


def open_hdf(path, mode="r"):
  """File object with the subset of the h5py API this package uses."""
  if _session is not None:
    import os
    key = os.path.abspath(path)
    if mode in ("a", "r+"):
      if key not in _session:
        _session[key] = _open(path, mode)
      return _Borrowed(_session[key])
    if key in _session:
      if mode == "r":
        _session[key].flush()                  # a reader sees what the session has written so far
      else:
        _session.pop(key).close()              # "w": the file is about to be truncated
  return _open(path, mode)


_readers = {}       # (path, mtime, size) -> open read-only file, for inputs that are opened once per trajectory
_readers_lock = __import__("threading").Lock()


class _SharedReader:
  """A read-only file kept open between uses (`shared_reader`): leaving its `with` block does not close it."""

  def __init__(self, hdf):
    self._hdf = hdf

  def __enter__(self):
    return self._hdf

  def __exit__(self, *exc):
    return False

  def close(self):
    pass

  def __getattr__(self, name):
    return getattr(self._hdf, name)

  def __contains__(self, key):
    return key in self._hdf

  def __getitem__(self, key):
    return self._hdf[key]


def shared_reader(path):
  """`open_hdf(path, "r")` for a file that is read again and again (MISATO's MD.hdf5: one open per trajectory in the
  reference, /root/reference/nearl/io/traj.py:375): the handle -- and what it has learnt about the file, e.g. the names in
  the root group -- is kept until the file changes on disk.  At most four files are kept."""
  import os
  st = os.stat(path)
  key = (os.path.abspath(path), st.st_mtime_ns, st.st_size)
  with _readers_lock:                     # loader threads share the handle (hdf5_lite reads by position)
    hdf = _readers.get(key)
    if hdf is None:
      for old in [k for k in _readers if k[0] == key[0]] + (list(_readers)[:1] if len(_readers) >= 4 else []):
        stale = _readers.pop(old, None)
        if stale is not None:
          try:
            stale.close()
          except Exception:
            pass
      hdf = _readers[key] = _open(path, "r")
  return _SharedReader(hdf)


@contextlib.contextmanager
def hdf_session():
  """Keep every file opened for appending open until the context exits (nested sessions join the outer one)."""
  global _session
  if _session is not None:
    yield
    return
  _session = {}
  try:
    yield
  finally:
    session, _session = _session, None
    for hdf in session.values():
      hdf.close()


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


def update_hdf_data(hdffile, dataset_name, data, hdf_slice, **kwargs):
  """Write `data` into the rows `hdf_slice` of a dataset, growing it along axis 0 when the slice ends beyond it; the
  dataset is created from `data` (with `kwargs`: dtype, maxshape ...) when it does not exist yet (utils.py:524-559).
  `hdffile`: a path or an open file."""
  if isinstance(hdffile, (str, bytes)) or hasattr(hdffile, "__fspath__"):
    with open_hdf(hdffile, "a") as hdf:
      return update_hdf_data(hdf, dataset_name, data, hdf_slice, **kwargs)
  data = np.asarray(data)
  if dataset_name not in hdffile:
    hdffile.create_dataset(dataset_name, data=data, **kwargs)
    return
  dset = hdffile[dataset_name]
  if hdf_slice.stop is not None and hdf_slice.stop > dset.shape[0]:
    dset.resize(hdf_slice.stop, axis=0)
  dset[hdf_slice] = data


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


# ---- residue blocks of Featurizer.loop_by_residue ---------------------------------------------------
def get_pdbcode(pdbcode):
  """Lower-case PDB code, superseded entries replaced (utils.py:196-214)."""
  pdbcode = pdbcode.lower()
  return constants.PDBCODE_SUPERCEDES.get(pdbcode, pdbcode)


def _pair_distances(a, b):
  """[len(a), len(b)] Euclidean distances in the input dtype: sqrt((dx^2 + dy^2) + dz^2), the quantity (and the order of
  additions) of np.linalg.norm(a[i] - b, axis=1) and of scipy.spatial.distance_matrix, which the reference calls at
  features.py:1727 and utils.py:118-136; only row-wise argmins are taken from it.  Column by column: a reduction over an
  axis of length 3 is the slow way to say this in numpy."""
  at, bt = np.ascontiguousarray(np.asarray(a).T), np.ascontiguousarray(np.asarray(b).T)     # contiguous coordinate columns
  dx = at[0][:, None] - bt[0][None, :]
  dy = at[1][:, None] - bt[1][None, :]
  dz = at[2][:, None] - bt[2][None, :]
  dx *= dx
  dy *= dy
  dz *= dz
  dx += dy
  dx += dz
  return np.sqrt(dx, out=dx)


def pair_track(xyz, first, second):
  """[len(first), n_frames] distance of atom first[i] to atom second[i] in every frame, in the dtype of `xyz`:
  sqrt((dx^2 + dy^2) + dz^2), the arithmetic of np.sqrt(np.sum((a - b) ** 2, axis=1)) at utils.py:146-147."""
  d = np.take(xyz, first, axis=1) - np.take(xyz, second, axis=1)
  dx, dy, dz = d[:, :, 0], d[:, :, 1], d[:, :, 2]
  return np.sqrt((dx * dx + dy * dy) + dz * dz).T


def compute_pcdt(traj, mask1, mask2, use_mean=False, ref=None, return_info=False):
  """Pairwise Closest DisTance (utils.py:70-166): every atom of `mask1` is paired with the atom of `mask2`
  closest to it in the reference frame (a frame index, a Frame, the mean structure, or frame 0); returns the
  distance of each pair in every frame, ``[len(mask1), n_frames]`` float64."""
  from . import io
  is_number = isinstance(ref, (int, float, np.int32, np.int64, np.float32, np.float64)) and not isinstance(ref, bool)
  if ref is not None and is_number:
    traj.top.set_reference(traj[int(ref)])
  elif ref is not None and isinstance(ref, io.Frame):
    traj.top.set_reference(ref)
  selmask1 = traj.top.select(mask1) if isinstance(mask1, str) else np.asarray(mask1)
  selmask2 = traj.top.select(mask2) if isinstance(mask2, str) else np.asarray(mask2)
  if len(selmask1) == 0:
    print("No target atom selected, please check the mask")
    return None, None
  if len(selmask2) == 0:
    print("No counter part atom selected, please check the mask")
    return None, None
  if use_mean is True:
    base = np.mean(traj.xyz, axis=0)
  elif isinstance(ref, io.Frame):
    base = ref.xyz
  elif ref is not None and is_number:
    base = traj.xyz[int(ref)]
  else:
    base = traj.xyz[0]
  ref_frame = _pair_distances(base[selmask1], base[selmask2])
  selclosest = selmask2[np.argmin(ref_frame, axis=1)]
  distarr = np.zeros((len(selmask1), traj.n_frames))
  distarr[:, :] = pair_track(traj.xyz, selmask1, selclosest)
  if return_info:
    atoms = traj.atoms if hasattr(traj, "atoms") else list(traj.top.atoms)
    info = {"atom_name_group1": [atoms[i].name for i in selmask1], "indices_group1": [atoms[i].index for i in selmask1],
            "atom_name_group2": [atoms[i].name for i in selclosest], "indices_group2": [atoms[i].index for i in selclosest]}
    return distarr, info
  return distarr


def rmsd_nofit(xyz, index, ref_xyz):
  """Per-frame RMSD of the atoms `index` against `ref_xyz [N, 3]` without fitting and without mass weighting
  (what pytraj.rmsd_nofit returns for `mask=`, `ref=`; features.py:1664,1672)."""
  d = np.asarray(xyz, dtype=np.float64)[:, index, :] - np.asarray(ref_xyz, dtype=np.float64)[index][None]
  return np.sqrt(np.mean(np.sum(d * d, axis=2), axis=1))


def _known_residue(name):
  return name in constants.RES or name in constants.RES_PATCH


def find_block_single(traj, key):
  """Atom slices of every residue named `key`, in topology order (utils.py:573-596)."""
  return [np.s_[res.first_atom_index:res.last_atom_index] for res in traj.top.residues if res.name == key]


def dual_blocks_by_key(traj):
  """{name1 + name2: [atom slices]} of every bonded pair of consecutive residues, in topology order: the lists
  find_block_dual returns for each key, from one pass over the residues (memoised on the topology + first frame)."""
  memo = traj.top.__dict__.get("_dual_blocks_memo")
  if memo is not None and memo[0] is traj.xyz:
    return memo[1]
  threshold = 1.32 * 1.25
  coord = np.asarray(traj.xyz[0])
  out = {}
  previous = None
  names = np.asarray(traj.top.names, dtype=object) if hasattr(traj.top, "names") else None
  for this in traj.top.residues:
    if previous is not None and _known_residue(previous.name) and _known_residue(this.name):
      b1, e1, b2, e2 = previous.first_atom_index, previous.last_atom_index, this.first_atom_index, this.last_atom_index
      if names is not None:
        atomc, atomn = np.flatnonzero(names[b1:e1] == "C"), np.flatnonzero(names[b2:e2] == "N")
      else:
        atomc, atomn = traj.top[b1:e1].select("@C"), traj.top[b2:e2].select("@N")
      if len(atomc) == 1 and len(atomn) == 1 and np.linalg.norm(coord[b1 + atomc[0]] - coord[b2 + atomn[0]]) <= threshold:
        out.setdefault(f"{previous.name}{this.name}", []).append(np.s_[b1:e2])
    previous = this
  traj.top.__dict__["_dual_blocks_memo"] = (traj.xyz, out)
  return out


def find_block_dual(traj, key):
  """Atom slices of every pair of consecutive residues whose names concatenate to `key` and whose
  peptide bond (C of the first, N of the second, first frame of the trajectory) is at most
  1.32 * 1.25 A long (utils.py:599-659)."""
  threshold = 1.32 * 1.25
  coord = np.asarray(traj.xyz[0])
  results = []
  previous = None
  for this in traj.top.residues:
    if previous is not None and f"{previous.name}{this.name}" == key:
      top1 = traj.top[previous.first_atom_index:previous.last_atom_index]
      top2 = traj.top[this.first_atom_index:this.last_atom_index]
      atomc, atomn = top1.select("@C"), top2.select("@N")
      if not (_known_residue(previous.name) and _known_residue(this.name)):
        print(f"Residue pair {previous.name}-{this.name} is not supported yet", file=sys.stderr)
      elif len(atomc) == 0 or len(atomn) == 0:
        pass
      elif len(atomc) > 1 or len(atomn) > 1:
        print("Warning: More than one C or N atom found in the residue, meaning there are some problem with the input structure.")
      else:
        crd1 = coord[previous.first_atom_index:previous.last_atom_index]
        crd2 = coord[this.first_atom_index:this.last_atom_index]
        if np.linalg.norm(crd1[atomc[0]] - crd2[atomn[0]]) <= threshold:
          results.append(np.s_[previous.first_atom_index:this.last_atom_index])
    previous = this
  return results
