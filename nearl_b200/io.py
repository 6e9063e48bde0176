"""Trajectory / topology duck types and the synthetic systems of the benchmark.

The featuriser only needs, from a trajectory, what ``Featurizer.run`` touches in the
reference (/root/reference/nearl/featurizer.py:229-241,345-397,399-469 and
/root/reference/nearl/features.py:397-423,846,1133-1230): ``n_frames``, ``xyz`` as an
``[F, N, 3]`` array, ``identity``, and a topology with ``n_atoms``, ``atoms`` (items with
``index``, ``name``, ``resid``, ``resname``, ``atomic_number``) and ``select(mask)``.
pytraj objects satisfy this; so do the light classes here, which exist because pytraj (the
reference's trajectory back-end, /root/reference/nearl/io/traj.py:17) is not part of this
path and not installed.  Any sequence of such objects is a valid "trajloader"
(featurizer.py:295,405 only use ``len`` and ``[]``).
"""
import os
import re

import numpy as np

from . import constants

__all__ = ["Atom", "Residue", "Topology", "Trajectory", "TrajectoryLoader", "MisatoTraj", "load_amber_topology", "load_pdb",
           "write_pdb", "load_coordinates", "load_topology", "superpose",
           "synthetic_trajectory", "synthetic_complex", "synthetic_peptide", "pinned_empty"]


def pinned_empty(shape, dtype):
  """Page-locked host array (torch is used for the allocation only) so that host<->device copies
  inside the C ABI run as DMA; falls back to ordinary memory without a CUDA device."""
  try:
    import torch
    if torch.cuda.is_available():
      tdtype = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8}[np.dtype(dtype)]
      return torch.empty(tuple(int(s) for s in shape), dtype=tdtype, pin_memory=True).numpy()
  except Exception:
    pass
  return np.empty(shape, dtype=dtype)


class Atom:
  __slots__ = ("index", "name", "resid", "resname", "atomic_number")

  def __init__(self, index, name, resid, resname, atomic_number):
    self.index = int(index)
    self.name = name
    self.resid = int(resid)
    self.resname = resname
    self.atomic_number = int(atomic_number)


class Residue:
  """The attributes of a pytraj Residue the reference reads (utils.py:590-594,617-652)."""
  __slots__ = ("index", "name", "first_atom_index", "last_atom_index")

  def __init__(self, index, name, first_atom_index, last_atom_index):
    self.index = int(index)
    self.name = name
    self.first_atom_index = int(first_atom_index)
    self.last_atom_index = int(last_atom_index)

  def __repr__(self):
    return f"<Residue {self.name}{self.index} atoms {self.first_atom_index}:{self.last_atom_index}>"


class Topology:
  """Column-store topology with an Amber-mask subset (``select``)."""

  def __init__(self, names, resids, resnames, atomic_numbers, charges=None):
    self.names = np.asarray(names, dtype=object)
    self.resids = np.asarray(resids, dtype=np.int64)
    self.resnames = np.asarray(resnames, dtype=object)
    self.atomic_numbers = np.asarray(atomic_numbers, dtype=np.int64)
    n = len(self.names)
    assert len(self.resids) == n and len(self.resnames) == n and len(self.atomic_numbers) == n
    # partial charges in units of the electron charge (pytraj's Topology.charge; zero when the source has none)
    self.charge = np.zeros(n, dtype=np.float64) if charges is None else np.asarray(charges, dtype=np.float64)
    assert len(self.charge) == n
    self._atoms = None

  @property
  def n_atoms(self):
    return len(self.names)

  @property
  def atoms(self):
    if self._atoms is None:
      self._atoms = [Atom(i, self.names[i], self.resids[i], self.resnames[i], self.atomic_numbers[i])
                     for i in range(self.n_atoms)]
    return self._atoms

  def residue_blocks(self):
    """(name, first_atom_index, last_atom_index) of every residue = maximal run of atoms with one residue id;
    `last_atom_index` is one past the residue's last atom, as in pytraj / cpptraj."""
    memo = self.__dict__.get("_residue_blocks")
    if memo is None:
      n = self.n_atoms
      starts = np.flatnonzero(np.r_[True, self.resids[1:] != self.resids[:-1]]) if n else np.zeros(0, dtype=int)
      ends = np.r_[starts[1:], n] if n else starts
      memo = self._residue_blocks = [(str(self.resnames[b]), int(b), int(e)) for b, e in zip(starts, ends)]
    return memo

  @property
  def residues(self):
    """Iterator over the residues (pytraj's Topology.residues is one, and
    /root/reference/nearl/utils.py:617-623 drives it with next())."""
    return iter(Residue(i, name, b, e) for i, (name, b, e) in enumerate(self.residue_blocks()))

  def __getitem__(self, key):
    idx = np.arange(self.n_atoms)[key]
    return Topology(self.names[idx], self.resids[idx], self.resnames[idx], self.atomic_numbers[idx], self.charge[idx])

  def set_reference(self, frame):
    """pytraj's Topology.set_reference (the frame distance-based masks refer to; /root/reference/nearl/utils.py:97-100
    sets it before selecting).  Distance-based masks are outside the supported subset: the frame is only kept."""
    self._reference = frame

  def strip(self, mask):
    """Remove the atoms selected by `mask`, in place (pytraj's Topology.strip; MisatoTraj drops water and
    ions this way, /root/reference/nearl/io/traj.py:362-369)."""
    keep = np.ones(self.n_atoms, dtype=bool)
    keep[self.select(mask)] = False
    self.names, self.resids = self.names[keep], self.resids[keep]
    self.resnames, self.atomic_numbers = self.resnames[keep], self.atomic_numbers[keep]
    self.charge = self.charge[keep]
    self._atoms = None
    for key in ("_select_cache", "_residue_blocks", "_dual_blocks_memo"):
      self.__dict__.pop(key, None)

  # ---- Amber mask subset: ':RES', ':1-10', ':A,B', ':*', '@NAME', '@1-5', ':RES@NAME', '@/El', '!', '&', '|', '()',
  # ---- name wildcards '*' / '=' (any run of characters) and '?' (one character), e.g. '!@H='
  def select(self, mask):
    cache = self.__dict__.setdefault("_select_cache", {})
    if mask not in cache:
      cache[mask] = self._select(mask)
    return cache[mask].copy()

  def _select(self, mask):
    mask = mask.strip()
    mask = re.sub(r",\s*(?=[:@])", "|", mask)       # '@C, @N, @CA' (features.py:1725): a list of masks is their union
    tokens = re.findall(r"[()!&|]|[:@][^()!&|:@\s]+", mask)
    if "".join(tokens) != re.sub(r"\s+", "", mask) or any(
        t[0] in ":@" and not re.fullmatch(r"[:@][A-Za-z0-9*=?,\-+'/_]+", t) for t in tokens):
      raise ValueError(f"unsupported Amber mask: {mask!r}")
    pos = [0]

    def peek():
      return tokens[pos[0]] if pos[0] < len(tokens) else None

    def take():
      pos[0] += 1
      return tokens[pos[0] - 1]

    def primary():
      tok = peek()
      if tok == "(":
        take()
        val = expr()
        if take() != ")":
          raise ValueError(f"unbalanced parentheses in mask {mask!r}")
        return val
      if tok == "!":
        take()
        return ~primary()
      if tok is None or tok in ")&|":
        raise ValueError(f"malformed mask {mask!r}")
      val = self._spec(take())
      while peek() is not None and peek()[0] in ":@":   # ':RES@CA' is an implicit AND
        val = val & self._spec(take())
      return val

    def conj():
      val = primary()
      while peek() == "&":
        take()
        val = val & primary()
      return val

    def expr():
      val = conj()
      while peek() == "|":
        take()
        val = val | conj()
      return val

    result = expr()
    if pos[0] != len(tokens):
      raise ValueError(f"malformed mask {mask!r}")
    return np.flatnonzero(result)

  def _spec(self, tok):
    kind, body = tok[0], tok[1:]
    out = np.zeros(self.n_atoms, dtype=bool)
    if kind == ":":
      _, first, inverse = np.unique(self.resids, return_index=True, return_inverse=True)
      resnum = np.argsort(np.argsort(first))[inverse] + 1     # 1-based residue number in order of appearance
      for part in body.split(","):
        if part == "*":
          out[:] = True
        elif re.fullmatch(r"\d+(-\d+)?", part):
          lo, _, hi = part.partition("-")
          out |= (resnum >= int(lo)) & (resnum <= int(hi or lo))
        else:
          out |= self._name_match(self.resnames, part)
    else:
      if body.startswith("/"):
        for sym in body[1:].split(","):
          out |= self.atomic_numbers == constants.SYMBOL2ATOMICNR.get(sym, -1)
        return out
      for part in body.split(","):
        if part == "*":
          out[:] = True
        elif re.fullmatch(r"\d+(-\d+)?", part):
          lo, _, hi = part.partition("-")
          idx = np.arange(1, self.n_atoms + 1)
          out |= (idx >= int(lo)) & (idx <= int(hi or lo))
        else:
          out |= self._name_match(self.names, part)
    return out

  @staticmethod
  def _name_match(names, pattern):
    if not any(ch in pattern for ch in "*=?"):
      return names == pattern
    rx = re.compile("".join(".*" if ch in "*=" else "." if ch == "?" else re.escape(ch) for ch in pattern))
    uniq, inverse = np.unique(names, return_inverse=True)
    return np.array([rx.fullmatch(str(u)) is not None for u in uniq], dtype=bool)[inverse]


class Frame:
  """One frame of a trajectory (pytraj.Frame as far as the label features read it: ``xyz [N, 3]``)."""

  def __init__(self, xyz):
    self.xyz = np.asarray(xyz)

  @property
  def n_atoms(self):
    return self.xyz.shape[0]


class Trajectory:
  """Coordinates ``xyz [F, N, 3]`` (float32 or float64) + topology + identity
  (/root/reference/nearl/io/traj.py:17-155, which wraps ``pytraj.load``).  Three ways in, as in the reference:

  * arrays -- ``Trajectory(xyz, top[, identity])`` (what this package and its tests build);
  * file names -- ``Trajectory("traj.nc", "top.parm7")``, ``Trajectory("traj.nc", "top.pdb")``,
    ``Trajectory("complex.pdb")`` (a PDB file is its own topology; one frame per MODEL); read here without pytraj:
    PDB by columns, Amber NetCDF trajectories / restarts through ``scipy.io.netcdf_file``, parm7 as `load_amber_topology`;
  * another trajectory -- ``Trajectory(traj, mask=":LIG")``.

  Keyword arguments of the reference: ``stride`` (wins over) ``frame_indices``, ``mask``, ``identity`` (default: the
  trajectory file name), ``superpose`` (fit onto the first frame over ``@CA`` on the device); ``dtype`` (default
  float64, what pytraj hands out; ``np.float32`` keeps the precision of a NetCDF trajectory and halves the upload)."""

  @staticmethod
  def read_raw(traj_src=None, pdb_src=None, identity=None, **kwarg):
    """The file half of the constructor for a trajectory given by file names (None otherwise): coordinates and topology as
    read, in the requested dtype.  `TrajectoryLoader.prefetched` runs it on loader threads and hands the result to the
    constructor as `_raw`; nothing here touches the device."""
    import os
    if not isinstance(traj_src, (str, os.PathLike)):
      return None
    traj_src = os.fspath(traj_src)
    dtype = np.dtype(kwarg.get("dtype", np.float64))
    if pdb_src is None:
      xyz, top = load_pdb(traj_src)
    else:
      pdb_src = os.fspath(pdb_src)
      top = load_topology(pdb_src)
      xyz = load_coordinates(traj_src)
    return {"xyz": np.ascontiguousarray(xyz, dtype=dtype), "top": top, "top_filename": pdb_src, "traj_filename": traj_src}

  def __init__(self, traj_src=None, pdb_src=None, identity=None, **kwarg):
    stride, frame_indices, mask = kwarg.get("stride"), kwarg.get("frame_indices"), kwarg.get("mask")
    self.top_filename = self.traj_filename = None
    if isinstance(traj_src, (str, os.PathLike)):
      raw = kwarg.pop("_raw", None) or self.read_raw(traj_src, pdb_src, **kwarg)
      xyz, top = raw["xyz"], raw["top"]
      self.top_filename, self.traj_filename = raw["top_filename"], raw["traj_filename"]
    elif isinstance(traj_src, Trajectory):
      xyz, top = traj_src.xyz, traj_src.top
      self.top_filename, self.traj_filename = traj_src.top_filename, traj_src.traj_filename
      identity = identity if identity is not None else traj_src.identity_
    elif traj_src is None and pdb_src is None:
      raise ValueError("Trajectory needs coordinates: arrays, file names or another trajectory")
    else:
      xyz, top = np.asarray(traj_src), pdb_src
    if xyz.ndim == 2:
      xyz = xyz[None]
    if not (xyz.ndim == 3 and xyz.shape[2] == 3 and xyz.shape[1] == top.n_atoms):
      raise ValueError(f"The number of atoms in the topology and the coordinates should be the same, rather than {top.n_atoms} and {xyz.shape}")
    if stride is not None:
      xyz = np.ascontiguousarray(xyz[::int(stride)])
    elif frame_indices is not None:
      xyz = np.ascontiguousarray(xyz[np.asarray(frame_indices, dtype=int)])
    if mask is not None:
      sel = top.select(mask)
      xyz, top = np.ascontiguousarray(xyz[:, sel, :]), top[sel]
    self.xyz = xyz
    self.top = top
    self.mask = mask
    self.identity_ = identity
    if kwarg.get("superpose", False):
      superpose(self, mask="@CA")

  @property
  def identity(self):
    """What meta-data is looked up under (traj.py:160-174): the given identity, else the trajectory file name."""
    if self.identity_ is not None:
      return self.identity_
    return self.traj_filename if self.traj_filename is not None else "traj"

  @identity.setter
  def identity(self, value):
    self.identity_ = value

  @property
  def n_frames(self):
    return self.xyz.shape[0]

  @property
  def n_atoms(self):
    return self.xyz.shape[1]

  @property
  def atoms(self):
    return np.array(self.top.atoms, dtype=object)            # traj.py:192-193 (make_index)

  @property
  def residues(self):
    return np.array(list(self.top.residues), dtype=object)

  def __len__(self):
    return self.n_frames

  def __getitem__(self, index):
    """`traj[i]` -> Frame (copy), `traj[slice]` -> Trajectory of those frames, `traj[":LIG"]` -> Trajectory of those
    atoms (pytraj's indexing, as far as the label features and scripts use it)."""
    if isinstance(index, (int, np.integer)):
      return Frame(np.array(self.xyz[int(index)]))
    if isinstance(index, str):
      return Trajectory(self, mask=index)
    out = Trajectory(np.ascontiguousarray(self.xyz[index]), self.top, identity=self.identity_)
    out.top_filename, out.traj_filename = self.top_filename, self.traj_filename
    return out

  def copy_traj(self):
    out = Trajectory(self.xyz.copy(), self.top[:], identity=self.identity_)
    out.top_filename, out.traj_filename = self.top_filename, self.traj_filename
    return out

  def write_frame(self, frames, outfile="", mask="", **_ignored):
    """Save frames (an index, a list, a slice) of the atoms of `mask` as a PDB file, one MODEL per frame
    (traj.py:195-236)."""
    if isinstance(frames, (int, np.integer)):
      coords = self.xyz[[int(frames)]]
    else:
      coords = self.xyz[np.arange(self.n_frames)[frames] if isinstance(frames, (slice, np.ndarray)) else np.asarray(frames, dtype=int)]
    top = self.top
    if len(mask) > 0:
      sel = self.top.select(mask)
      coords, top = coords[:, sel, :], self.top[sel]
    if len(outfile) > 0:
      write_pdb(outfile, coords, top)


# ---- file formats (what pt.load reads for the reference: traj.py:97-106) -------------------------------------------
def _guess_atomic_number(name_field, resname):
  """Element of a PDB atom without an element column, from the atom-name columns 13-16: two-letter symbols are
  right-justified into columns 13-14 ('FE  ', 'CL  ', 'CA  ' = calcium), everything else starts in column 14
  (' CA ' = C-alpha)."""
  two = name_field[:2].strip().capitalize()
  if name_field[0].isalpha() and len(two) == 2 and two in constants.SYMBOL2ATOMICNR and (two.upper() == resname.strip().upper()[:2] or two not in ("Ca", "Cd", "Ce", "Cf", "Cm", "Co", "Cr", "Cs", "Cu", "Hf", "Hg", "Ho", "Hs", "Nb", "Nd", "Ne", "Ni", "No", "Np", "Os", "Sb", "Sc", "Se", "Sg", "Si", "Sm", "Sn", "Sr")):
    return constants.SYMBOL2ATOMICNR[two]
  core = name_field.strip().lstrip("0123456789")
  return constants.SYMBOL2ATOMICNR.get(core[:1].upper(), 0) if core else 0


def _load_pdb_native(data):
  """`load_pdb` on the column parser of the native library (`nrl_pdb_parse`: host code, one pass over the text for the
  coordinates and the fixed-width columns of the first model), names / residues / elements resolved once per DISTINCT
  value.  None when the library is not there or finds the text not plain enough: the caller reads it line by line, which
  also words the errors."""
  import ctypes
  try:
    from . import _lib
    lib = _lib.lib()
  except Exception:
    return None
  counts = (ctypes.c_longlong * 3)()
  ref = lambda k: ctypes.byref(counts, 8 * k)
  if lib.nrl_pdb_parse(data, len(data), ref(0), ref(1), ref(2), None, None) != 0:
    return None
  n_records, n, n_frames = int(counts[0]), int(counts[1]), int(counts[2])
  xyz = np.empty((n_frames, n, 3), dtype=np.float64)
  cols = np.empty((n, 16), dtype=np.uint8)
  if n_records != n_frames * n or lib.nrl_pdb_parse(data, len(data), ref(0), ref(1), ref(2), xyz.ctypes.data_as(ctypes.c_void_p),
                                                     cols.ctypes.data_as(ctypes.c_void_p)) != 0:
    return None

  # every column group as an integer per record: sorting integers is what np.unique is fast at
  def codes(lo, hi):
    width = hi - lo
    raw = np.zeros((n, 8), dtype=np.uint8)
    raw[:, :width] = cols[:, lo:hi]
    values, inverse = np.unique(raw.view("<u8").ravel(), return_inverse=True)
    texts = [int(v).to_bytes(8, "little")[:width].decode() for v in values.tolist()]
    return texts, inverse
  name_texts, name_code = codes(0, 4)
  res_texts, res_raw_code = codes(4, 8)
  elem_texts, elem_code = codes(14, 16)
  names = np.array([t.strip() for t in name_texts], dtype=object)[name_code]
  res_stripped = [t.strip() for t in res_texts]
  resnames = np.array(res_stripped, dtype=object)[res_raw_code]
  # element: from columns 77-78, else guessed from the name columns and the residue name -- once per distinct combination
  combo = (elem_code.astype(np.int64) * len(name_texts) + name_code) * len(res_texts) + res_raw_code
  combos, combo_inverse = np.unique(combo, return_inverse=True)
  z_of = np.empty(len(combos), dtype=np.int64)
  for k, c in enumerate(combos.tolist()):
    c, r = divmod(c, len(res_texts))
    e, a = divmod(c, len(name_texts))
    sym = elem_texts[e].strip().capitalize()
    z_of[k] = constants.SYMBOL2ATOMICNR[sym] if sym in constants.SYMBOL2ATOMICNR else _guess_atomic_number(name_texts[a], res_stripped[r])
  zs = z_of[combo_inverse]
  # residues: maximal runs of one (chain, number + insertion code, stripped name)
  stripped_id = {t: k for k, t in enumerate(dict.fromkeys(res_stripped))}
  res_code = np.array([stripped_id[t] for t in res_stripped], dtype=np.int64)[res_raw_code]
  change = np.ones(n, dtype=bool)
  change[1:] = np.any(cols[1:, 8:14] != cols[:-1, 8:14], axis=1) | (res_code[1:] != res_code[:-1])
  return xyz, Topology(names, np.cumsum(change) - 1, resnames, zs)


def load_pdb(path):
  """(xyz [frames, atoms, 3] float64, Topology) of a PDB file: ATOM / HETATM records by column (name 13-16, residue name
  18-21, chain 22, residue number 23-26, insertion code 27, x / y / z 31-54, element 77-78), one frame per MODEL (every
  model must list the same atoms; the topology is the first model's).  Residues are maximal runs of one (chain, number,
  insertion code, name); `resids` count them from 0 like pytraj's ``atom.resid``.  Alternate locations are all kept.
  The text goes through the native library's column parser (`nrl_pdb_parse`) when it is plain; this loop is the definition
  and the fallback."""
  import gzip
  with open(path, "rb") as fh:
    data = fh.read()
  if str(path).endswith(".gz"):
    data = gzip.decompress(data)
  fast = _load_pdb_native(data)
  if fast is not None:
    return fast
  return _load_pdb_lines(path)


def _load_pdb_lines(path):
  import gzip
  opener = gzip.open if str(path).endswith(".gz") else open
  names, resnames, reskeys, zs = [], [], [], []
  frames, coords = [], []
  first = True
  with opener(path, "rt") as fh:
    for line in fh:
      rec = line[:6]
      if rec in ("ATOM  ", "HETATM"):
        coords.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
        if first:
          name_field = line[12:16]
          resname = line[17:21].strip()
          names.append(name_field.strip())
          resnames.append(resname)
          reskeys.append((line[21], line[22:27], resname))
          sym = line[76:78].strip().capitalize() if len(line) >= 78 else ""
          zs.append(constants.SYMBOL2ATOMICNR[sym] if sym in constants.SYMBOL2ATOMICNR else _guess_atomic_number(name_field, resname))
      elif rec == "ENDMDL" or (rec.startswith("END") and coords):
        if coords:
          frames.append(coords)
          coords, first = [], False
  if coords:
    frames.append(coords)
  if not frames:
    raise ValueError(f"{path}: no ATOM / HETATM record")
  n = len(names)
  if any(len(f) != n for f in frames):
    raise ValueError(f"{path}: the models do not list the same number of atoms")
  change = np.ones(n, dtype=bool)
  change[1:] = [reskeys[i] != reskeys[i - 1] for i in range(1, n)]
  resids = np.cumsum(change) - 1
  return np.array(frames, dtype=np.float64), Topology(names, resids, resnames, zs)


def write_pdb(path, xyz, top):
  """Minimal PDB writer (one MODEL per frame) for `Trajectory.write_frame`."""
  xyz = np.asarray(xyz)
  if xyz.ndim == 2:
    xyz = xyz[None]
  with open(path, "w") as fh:
    for f, frame in enumerate(xyz):
      if len(xyz) > 1:
        fh.write(f"MODEL     {f + 1:4d}\n")
      for i in range(top.n_atoms):
        name = str(top.names[i])
        sym = constants.ATOMICNR2SYMBOL.get(int(top.atomic_numbers[i]), "")
        field = name.ljust(4) if (len(name) >= 4 or len(sym) == 2) else (" " + name).ljust(4)
        fh.write("ATOM  %5d %4s %-4s%1s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f          %2s\n" % (
          (i + 1) % 100000, field[:4], str(top.resnames[i])[:4], "A", (int(top.resids[i]) + 1) % 10000,
          frame[i, 0], frame[i, 1], frame[i, 2], 1.0, 0.0, sym.upper().rjust(2)))
      fh.write("ENDMDL\n" if len(xyz) > 1 else "END\n")


def load_coordinates(path):
  """Coordinates [frames, atoms, 3] of a trajectory file: Amber NetCDF trajectory (.nc / .ncdf / .netcdf: variable
  ``coordinates`` [frame, atom, spatial]) or restart (.ncrst / .rst7 in NetCDF form: [atom, spatial]) -- NetCDF-3, read
  through scipy --, or a PDB file (.pdb)."""
  lower = str(path).lower()
  if lower.endswith((".pdb", ".pdb.gz")):
    return load_pdb(path)[0]
  with open(path, "rb") as fh:
    magic = fh.read(4)
  if magic[:3] != b"CDF":
    raise ValueError(f"{path}: only NetCDF-3 (Amber .nc / .ncrst) and PDB trajectories are read without pytraj")
  from scipy.io import netcdf_file
  with netcdf_file(path, "r", mmap=False) as nc:
    if "coordinates" not in nc.variables:
      raise ValueError(f"{path}: no `coordinates` variable (not an Amber NetCDF file)")
    var = nc.variables["coordinates"]
    xyz = np.array(var[:])
    scale = getattr(var, "scale_factor", None)
    if scale is not None:
      xyz = xyz * scale
  if xyz.ndim == 2:
    xyz = xyz[None]
  return xyz


def load_topology(path):
  """Topology of a topology file: Amber parm7 (.parm7 / .prmtop / .top, optionally .gz) or PDB."""
  lower = str(path).lower()
  if lower.endswith((".pdb", ".pdb.gz")):
    return load_pdb(path)[1]
  return load_amber_topology(path)


_ELEMENTS = np.array([1, 6, 7, 8, 16])
_ELEMENT_P = np.array([0.50, 0.32, 0.085, 0.09, 0.005])


def _synthetic_topology(rng, n_atoms, ligand_atoms=0):
  """Consecutive 16-atom residues named cyclically over the 20 amino acids; the first four atoms
  of a residue are N, CA, C, O.  Optional trailing residue LIG."""
  n_prot = n_atoms - ligand_atoms
  z = rng.choice(_ELEMENTS, size=n_atoms, p=_ELEMENT_P)
  resid = np.arange(n_prot) // 16
  within = np.arange(n_prot) % 16
  for slot, (_, zz) in enumerate((("N", 7), ("CA", 6), ("C", 6), ("O", 8))):
    z[:n_prot][within == slot] = zz
  sym_lut = np.array([constants.ATOMICNR2SYMBOL.get(i, "X") for i in range(0, 119)], dtype=object)
  suffix = np.concatenate([within, np.arange(1, ligand_atoms + 1)]).astype(str).astype(object)
  names = sym_lut[z] + suffix
  for slot, nm in enumerate(("N", "CA", "C", "O")):
    names[:n_prot][within == slot] = nm
  resnames = np.empty(n_atoms, dtype=object)
  resnames[:n_prot] = np.array(constants.RES, dtype=object)[resid % 20]
  resids = np.empty(n_atoms, dtype=np.int64)
  resids[:n_prot] = resid
  if ligand_atoms:
    resnames[n_prot:] = "LIG"
    resids[n_prot:] = (resid[-1] + 1) if n_prot else 0
  return Topology(names, resids, resnames, z)


def synthetic_trajectory(n_atoms, n_frames, seed, dtype=np.float32, identity=None, pin=True, step=0.3, ligand_atoms=0):
  """Uniform-density (0.1 atoms/A^3) cube, random-walk dynamics of `step` A per frame; optionally the last
  `ligand_atoms` atoms are a ligand (residue LIG) at the cube centre."""
  rng = np.random.default_rng(seed)
  top = _synthetic_topology(rng, n_atoms, ligand_atoms=ligand_atoms)
  side = (n_atoms / 0.1) ** (1.0 / 3.0)
  xyz = pinned_empty((n_frames, n_atoms, 3), dtype) if pin else np.empty((n_frames, n_atoms, 3), dtype=dtype)
  frame = rng.uniform(0.0, side, size=(n_atoms, 3))
  if ligand_atoms:
    frame[n_atoms - ligand_atoms:] = side / 2 + rng.normal(0.0, 2.0, size=(ligand_atoms, 3))
  for t in range(n_frames):
    if t:
      frame = frame + rng.normal(0.0, step, size=(n_atoms, 3))
    xyz[t] = frame.astype(dtype)
  return Trajectory(xyz, top, identity or f"synthetic_{n_atoms}x{n_frames}_{seed}")


def synthetic_complex(n_atoms, seed, ligand_atoms=30, dtype=np.float32, identity=None, pin=False):
  """Single-frame protein + 30-atom ligand (residue LIG) at the cube centre (PDBBind-style)."""
  rng = np.random.default_rng(seed)
  top = _synthetic_topology(rng, n_atoms, ligand_atoms=ligand_atoms)
  side = (n_atoms / 0.1) ** (1.0 / 3.0)
  frame = rng.uniform(0.0, side, size=(n_atoms, 3))
  frame[n_atoms - ligand_atoms:] = side / 2 + rng.normal(0.0, 2.0, size=(ligand_atoms, 3))
  xyz = pinned_empty((1, n_atoms, 3), dtype) if pin else np.empty((1, n_atoms, 3), dtype=dtype)
  xyz[0] = frame.astype(dtype)
  return Trajectory(xyz, top, identity or f"complex_{n_atoms}_{seed}")


def synthetic_peptide(n_residues, n_frames, seed, dtype=np.float32, identity=None, pin=False, break_every=0):
  """A chain of residues of DIFFERENT lengths with real peptide geometry between neighbours (C of residue i
  1.32 A from N of residue i+1), for Featurizer.loop_by_residue: residue names are drawn from the 20 amino
  acids plus the HID / CYX variants, a residue of a given name always has the same number of atoms and some
  names share a length.  `break_every` > 0 opens the chain after every that many residues."""
  rng = np.random.default_rng(seed)
  pool = list(constants.RES) + ["HID", "CYX"]
  size_of = {name: 8 + (3 * i) % 11 for i, name in enumerate(pool)}
  res_names = [pool[i] for i in rng.integers(0, len(pool), size=n_residues)]
  centres = np.cumsum(rng.normal(0.0, 1.0, size=(n_residues + 1, 3)) * 2.2 + np.array([3.0, 0.0, 0.0]), axis=0)
  names, resids, resnames, z, xyz0 = [], [], [], [], []
  for r, rname in enumerate(res_names):
    n = size_of[rname]
    t_prev = centres[r] - centres[r - 1] if r > 0 else centres[1] - centres[0]
    t_next = centres[r + 1] - centres[r]
    t_prev, t_next = t_prev / np.linalg.norm(t_prev), t_next / np.linalg.norm(t_next)
    pos = centres[r] + rng.normal(0.0, 1.3, size=(n, 3))
    elem = rng.choice(_ELEMENTS, size=n, p=_ELEMENT_P)
    nm = [f"X{j}" for j in range(n)]
    # backbone: N next to the previous residue's C, CA at the centre, C towards the next residue, O beside it
    gap = 3.0 if (break_every and r % break_every == 0) else 0.0
    pos[0] = (centres[r - 1] + centres[r]) / 2 + (0.66 + gap) * t_prev if r > 0 else centres[r] - 1.4 * t_next
    pos[1] = centres[r]
    pos[2] = (centres[r] + centres[r + 1]) / 2 - 0.66 * t_next
    pos[3] = pos[2] + np.array([0.0, 1.2, 0.0])
    for slot, (an, zz) in enumerate((("N", 7), ("CA", 6), ("C", 6), ("O", 8))):
      nm[slot], elem[slot] = an, zz
    names += nm; resids += [r] * n; resnames += [rname] * n; z += list(elem); xyz0.append(pos)
  top = Topology(names, resids, resnames, z)
  frame = np.concatenate(xyz0, axis=0)
  n_atoms = len(frame)
  xyz = pinned_empty((n_frames, n_atoms, 3), dtype) if pin else np.empty((n_frames, n_atoms, 3), dtype=dtype)
  for t in range(n_frames):
    if t:
      frame = frame + rng.normal(0.0, 0.15, size=(n_atoms, 3))
    xyz[t] = frame.astype(dtype)
  return Trajectory(xyz, top, identity or f"peptide_{n_residues}x{n_frames}_{seed}")


# ---- trajectory ingestion without pytraj ------------------------------------------------------------
def _parm7_sections(text):
  """{flag: (format, [data lines])} of an Amber parm7 / prmtop file."""
  sections, flag = {}, None
  for line in text.splitlines():
    if line.startswith("%FLAG"):
      flag = line.split()[1]
      sections[flag] = [None, []]
    elif line.startswith("%FORMAT") and flag is not None:
      sections[flag][0] = line[line.index("(") + 1:line.rindex(")")]
    elif line.startswith("%") or flag is None:
      continue
    else:
      sections[flag][1].append(line)
  return sections


def _parm7_values(section):
  """Fixed-width fields of one section: Fortran formats like 20a4, 10I8, 5E16.8."""
  fmt, lines = section
  m = re.fullmatch(r"(\d+)([aAiIeEfF])(\d+)(?:\.\d+)?", fmt.strip())
  if not m:
    raise ValueError(f"unsupported parm7 format {fmt!r}")
  width, kind = int(m.group(3)), m.group(2).lower()
  out = []
  for line in lines:
    line = line.rstrip("\n")
    for i in range(0, len(line), width):
      field = line[i:i + width]
      if not field.strip() and kind != "a":
        continue
      out.append(field.strip() if kind == "a" else (int(field) if kind == "i" else float(field)))
  return out


_PARM7_REQUIRED = ("POINTERS", "ATOM_NAME", "RESIDUE_LABEL", "RESIDUE_POINTER")
_PARM7_WANTED = _PARM7_REQUIRED + ("CHARGE", "MASS", "ATOMIC_NUMBER")


def _parm7_head(path):
  """Bytes of a parm7 file up to the end of the last section the topology needs.  The per-atom and per-residue tables
  stand at the head of the file (tleap / parmed order); the bonded terms and exclusion lists that follow are four fifths
  of a solvated system's file, so a gzip member is inflated only as far as needed (the whole file when a wanted
  section has not shown up by then -- files that predate ATOMIC_NUMBER, unusual orders)."""
  import zlib
  with open(path, "rb") as fh:
    if not str(path).endswith(".gz"):
      return fh.read()
    inflate = zlib.decompressobj(16 + zlib.MAX_WBITS)
    buf = bytearray()
    whole, scan = set(), 0          # sections with a later %FLAG behind them; where the search for the next marker resumes
    last = None
    while True:
      chunk = fh.read(1 << 16)
      if not chunk:
        break
      buf += inflate.decompress(chunk)
      if inflate.eof:
        if not inflate.unused_data.strip(b"\0"):
          break
        raise ValueError("multi-member gzip")        # rare: the caller falls back to the gzip module
      while True:
        at = buf.find(b"%FLAG", scan)
        nl = buf.find(b"\n", at) if at >= 0 else -1
        if nl < 0:
          scan = at if at >= 0 else max(len(buf) - 4, scan)
          break
        if last is not None:
          whole.add(last)
        words = bytes(buf[at + 5:nl]).split()
        last = words[0].decode("ascii", "replace") if words else ""
        scan = nl
      if whole.issuperset(_PARM7_WANTED):
        break
    return bytes(buf)


def _parm7_sections_bytes(buf, wanted):
  """{flag: [format, body bytes]} of the wanted sections that are whole in `buf`."""
  marks, at = [], buf.find(b"%FLAG")
  while at >= 0:
    marks.append(at)
    at = buf.find(b"%FLAG", at + 5)
  marks.append(len(buf))
  sections = {}
  for at, end in zip(marks[:-1], marks[1:]):
    p = buf.find(b"\n", at, end) + 1
    if p <= 0:
      continue
    words = buf[at + 5:p].split()
    flag = words[0].decode() if words else ""
    if flag not in wanted or flag in sections:
      continue
    fmt = None
    while p < end and buf[p:p + 1] == b"%":                # %FORMAT(...), %COMMENT lines
      e = buf.find(b"\n", p, end)
      e = end if e < 0 else e
      line = buf[p:e]
      if line.startswith(b"%FORMAT"):
        fmt = line[line.index(b"(") + 1:line.rindex(b")")].decode()
      p = e + 1
    if fmt is not None:
      sections[flag] = [fmt, buf[p:end]]
  return sections


def _parm7_array(section, count=None, take=None):
  """The fields of one section as an array, parsed in one piece: every line but the last holds its full count of
  fixed-width fields (else None: the caller takes the line-by-line parser).  `count` / `take`: only the first `count`
  fields, and of those only the ones a boolean mask picks, are converted (the per-atom tables of a solvated system are
  nine tenths water that the caller drops)."""
  fmt, body = section
  m = re.fullmatch(r"(\d+)([aAiIeEfF])(\d+)(?:\.\d+)?", fmt.strip())
  if not m:
    return None
  per, kind, width = int(m.group(1)), m.group(2).lower(), int(m.group(3))
  body = body.replace(b"\r", b"")
  if body.endswith(b"\n"):
    body = body[:-1]
  full = per * width
  breaks = np.flatnonzero(np.frombuffer(body, dtype=np.uint8) == 10)
  if len(breaks) and not np.array_equal(breaks, (np.arange(len(breaks)) + 1) * (full + 1) - 1):
    return None
  flat = body.replace(b"\n", b"")
  if kind != "a":
    flat = flat.rstrip()
  flat += b" " * (-len(flat) % width)
  fields = np.frombuffer(flat, dtype=f"S{width}")
  if count is not None:
    fields = fields[:count]
  if take is not None:
    if len(take) != len(fields):
      return None
    fields = fields[take]
  try:
    if kind == "a":
      # names and labels repeat: strip and decode the distinct ones only
      distinct, inverse = np.unique(fields, return_inverse=True)
      return np.array([d.decode().strip() for d in distinct.tolist()], dtype=object)[inverse]
    return fields.astype(np.int64 if kind == "i" else np.float64)
  except ValueError:
    return None


def load_amber_topology(path, drop_residues=()):
  """Topology from an Amber parm7 / prmtop file, optionally gzip-compressed (what pt.load_topology gives
  MisatoTraj for production.top.gz, /root/reference/nearl/io/traj.py:361): atom names, residue labels and
  pointers, atomic numbers (from the masses when the file predates the ATOMIC_NUMBER section), charges.
  `drop_residues`: residue names whose atoms are left out -- the topology `strip(":NAME")` of each would leave, without
  converting the fields of the dropped atoms."""
  sec = None
  try:
    sec = _parm7_sections_bytes(_parm7_head(path), _PARM7_WANTED)
  except (OSError, ValueError, EOFError, __import__("zlib").error):
    pass
  if sec is None or any(need not in sec for need in _PARM7_REQUIRED):
    import gzip
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rt") as fh:
      sec = _parm7_sections(fh.read())
  for need in _PARM7_REQUIRED:
    if need not in sec:
      raise ValueError(f"{path}: parm7 section {need} is missing")

  def values(flag, count, take=None):
    arr = _parm7_array(sec[flag], count, take) if isinstance(sec[flag][1], bytes) else None
    if arr is None:
      fmt, body = sec[flag]
      arr = np.array(_parm7_values([fmt, body.decode().splitlines() if isinstance(body, bytes) else body])[:count],
                     dtype=object if "a" in fmt.lower() else None)
      if take is not None:
        arr = arr[take]
    return arr

  pointers = values("POINTERS", None)
  n_atoms, n_res = int(pointers[0]), int(pointers[11])
  labels = values("RESIDUE_LABEL", n_res)
  first = np.asarray(values("RESIDUE_POINTER", n_res), dtype=np.int64) - 1
  sizes = np.diff(np.r_[first, n_atoms])
  resids = np.repeat(np.arange(n_res), sizes)
  take = None
  if len(drop_residues):
    keep_res = ~np.isin(np.asarray(labels, dtype=object).astype(str), [str(r) for r in drop_residues])
    if not keep_res.all():
      take = np.repeat(keep_res, sizes)
      resids = resids[take]
  names = values("ATOM_NAME", n_atoms, take)
  resnames = np.asarray(labels, dtype=object)[resids]
  if "ATOMIC_NUMBER" in sec:
    z = np.asarray(values("ATOMIC_NUMBER", n_atoms, take), dtype=np.int64)
  else:
    masses = np.asarray(values("MASS", n_atoms, take), dtype=float)
    table = np.array([constants.ATOMICMASS.get(i, 1e9) for i in range(1, 119)])
    z = np.abs(masses[:, None] - table[None, :]).argmin(axis=1) + 1
  charges = None
  if "CHARGE" in sec:
    # parm7 stores charges multiplied by 18.2223 (the square root of the electrostatic constant in kcal/mol units)
    charges = np.asarray(values("CHARGE", n_atoms, take), dtype=np.float64) / 18.2223
  return Topology(names, resids, resnames, z, charges)


def superpose(traj, mask="@CA", ref=0):
  """Fit every frame of `traj` onto frame `ref` over the atoms of `mask`, in place, on the device
  (pt.superpose(traj, mask="@CA"), /root/reference/nearl/io/traj.py:398-404).  Returns the per-frame RMSD."""
  from . import _lib
  idx = np.ascontiguousarray(traj.top.select(mask), dtype=np.int32)
  if len(idx) == 0:
    raise ValueError(f"superpose: no atom matches {mask!r}")
  xyz = traj.xyz
  if not (isinstance(xyz, np.ndarray) and xyz.flags.c_contiguous and xyz.flags.writeable and xyz.dtype in (np.float32, np.float64)):
    raise ValueError("superpose needs a writable C-contiguous float32 / float64 coordinate array")
  rmsd = np.zeros(xyz.shape[0], dtype=np.float32)
  _lib.check(_lib.lib().nrl_superpose(_lib.ptr(xyz), int(xyz.dtype == np.float64), xyz.shape[0], xyz.shape[1], _lib.ptr(idx), len(idx),
                                      int(ref), 0, _lib.ptr(rmsd), None))
  return rmsd


class MisatoTraj(Trajectory):
  """A trajectory of the MISATO data set (/root/reference/nearl/io/traj.py:287-423) without pytraj:
  topology from ``{misatodir}/parameter_restart_files_MD/{pdbcode}/production.top.gz`` with water and ions
  stripped, coordinates from ``{misatodir}/MD.hdf5:/{PDBCODE}/trajectory_coordinates``.  Keyword arguments as
  in the reference: ``stride`` or ``frame_indices``, ``mask``, ``superpose`` (fit onto the first frame over
  ``@CA``, done on the device)."""

  def __init__(self, pdbcode, misatodir, _raw=None, **kwarg):
    raw = _raw if _raw is not None else self.read_raw(pdbcode, misatodir, **kwarg)
    self.topfile, self.trajfile, self.pdbcode = raw["topfile"], raw["trajfile"], pdbcode
    coord, top = raw["coord"], raw["top"]
    xyz = pinned_empty(coord.shape, coord.dtype if coord.dtype in (np.float32, np.float64) else np.float32)
    xyz[...] = coord
    from . import utils
    super().__init__(xyz, top, identity=utils.get_pdbcode(str(pdbcode)))       # traj.py:411-421: lower case, superseded codes replaced
    assert self.xyz.ndim == 3, f"Fatal: The shape of the trajectory {self.identity} is {self.xyz.shape}!!! It should be (n_frames, n_atoms, 3). "
    if kwarg.get("superpose", False):
      superpose(self, mask="@CA")

  @staticmethod
  def read_raw(pdbcode, misatodir, **kwarg):
    """The file half of the constructor -- topology parsed and stripped, coordinates read, strided and masked -- with no
    use of the device or of the native library: `TrajectoryLoader.prefetched` runs it on loader threads (inflating the
    files is most of it and releases the interpreter lock) and hands the result to the constructor as `_raw`."""
    import os
    from . import utils
    topfile = f"{misatodir}/parameter_restart_files_MD/{pdbcode.lower()}/production.top.gz"
    if not os.path.exists(topfile):
      raise FileNotFoundError(f"The topology file of PDB {pdbcode} is not found ({topfile})")
    trajfile = os.path.join(misatodir, "MD.hdf5")
    if not os.path.exists(trajfile):
      raise FileNotFoundError(f"The trajectory file is not found ({trajfile})")
    top = load_amber_topology(topfile, drop_residues=("WAT", "Cl-", "Na+"))   # keep the coordinates aligned with the topology (traj.py:362-369)
    with utils.shared_reader(trajfile) as hdf:          # kept open between trajectories
      if pdbcode.upper() not in hdf:
        raise ValueError(f"Not found the key for PDB code {pdbcode.upper()} in the HDF5 trajectory file.")
      coord = hdf[f"{pdbcode.upper()}/trajectory_coordinates"][:]
    if kwarg.get("stride", None) is not None:
      coord = coord[::int(kwarg["stride"])]
    elif kwarg.get("frame_indices", None) is not None:
      coord = coord[kwarg["frame_indices"]]
    if kwarg.get("mask", None) is not None:
      sel = top.select(kwarg["mask"])
      coord, top = coord[:, sel, :], top[sel]
    return {"coord": coord, "top": top, "topfile": topfile, "trajfile": trajfile}


class TrajectoryLoader:
  """Lazy list of trajectories (/root/reference/nearl/io/trajloader.py:12-247): every entry is a tuple of constructor
  arguments -- or an object that already is a trajectory -- and the entry's trajectory type builds it on access with the
  loading options (``stride``, ``frame_indices``, ``mask``, ``superpose``; per-trajectory ``strides`` / ``masks`` win
  over the common ones) and its identity (``trajids``).  ``loader[i]``, ``loader[[i, j]]``, ``loader[a:b]``, iteration,
  ``len`` and ``append`` as in the reference.  (The reference reads ``self.strides`` / ``self.masks`` before it assigns
  them when ``strides=`` / ``masks=`` are given, trajloader.py:92-104; here the lengths are checked on the arguments.)"""

  def __init__(self, trajs=None, trajtype=None, trajids=None, **kwarg):
    self.trajs = []
    if trajs is not None:
      if not hasattr(trajs, "__iter__") or isinstance(trajs, (str, bytes)):
        raise ValueError(f"The input should be a list or tuple of trajectory arguments rather than {type(trajs)}")
      for item in trajs:
        self.trajs.append(item if (hasattr(item, "xyz") and hasattr(item, "top")) else (tuple(item) if isinstance(item, (list, tuple)) else (item,)))
    n = len(self.trajs)
    self.OUTPUT_TYPE = [trajtype or Trajectory] * n
    if trajids is not None and len(trajids) != n:
      raise ValueError(f"The number of trajectory ids {len(trajids)} does not match the number of trajectories {n}")
    self.trajids = list(trajids) if trajids is not None else [None] * n
    self.i_ = 0
    self._loading_options = {"stride": None, "frame_indices": None, "mask": None, "superpose": False}
    strides, masks = kwarg.pop("strides", None), kwarg.pop("masks", None)
    self.loading_options = kwarg
    if strides is not None and len(strides) != n:
      raise ValueError(f"The manually defined number of strides {len(strides)} does not match the number of trajectories {n}")
    if masks is not None and len(masks) != n:
      raise ValueError(f"The manually defined number of masks {len(masks)} does not match the number of trajectories {n}")
    self.strides = [int(v) if v is not None else None for v in strides] if strides is not None else [None] * n
    self.masks = list(masks) if masks is not None else [None] * n

  # kept for callers of the first version of this class
  @property
  def trajtype(self):
    return self.OUTPUT_TYPE[0] if self.OUTPUT_TYPE else Trajectory

  @property
  def kwargs(self):
    return {k: v for k, v in self._loading_options.items() if v is not None and v is not False}

  @property
  def loading_options(self):
    """The common loading options that are set, plus the identity of the trajectory being processed."""
    options = {k: v for k, v in self._loading_options.items() if v is not None}
    if self.trajids and self.i_ < len(self.trajids):
      options["identity"] = self.trajids[self.i_]
    return options

  @loading_options.setter
  def loading_options(self, kwargs):
    self._loading_options.update(kwargs)

  def get_loading_options(self, index):
    """Constructor keywords of trajectory `index`: the common options, its own stride / mask, its identity."""
    options = {k: v for k, v in self._loading_options.items() if v is not None and v is not False}
    if self.strides[index] is not None:
      options["stride"] = self.strides[index]
    if self.masks[index] is not None:
      options["mask"] = self.masks[index]
    if self.trajids[index] is not None:
      options["identity"] = self.trajids[index]
    return options

  def _load(self, index):
    index = int(index)
    self.i_ = index
    item = self.trajs[index]
    if hasattr(item, "xyz") and hasattr(item, "top"):
      return item
    return self.OUTPUT_TYPE[index](*item, **self.get_loading_options(index))

  def prefetched(self, indices, workers=4, ahead=None):
    """(index, trajectory) for every index of `indices`, in order, with the FILE half of the next trajectories read ahead
    on `workers` threads: for trajectory types that split their constructor (`read_raw(*args, **options)` -> what the
    constructor takes as `_raw`; `MisatoTraj` does) up to `ahead` (default 2 x workers) reads are in flight while the
    caller works on the current trajectory.  Everything that touches the device -- page-locking, superposition -- stays on
    the calling thread, inside the constructor.  Other entries are built on access, as `loader[i]` does; an exception of a
    read is raised when its trajectory's turn comes, as the unthreaded loop would."""
    indices = [int(i) for i in indices]
    # (a type takes part when it defines `read_raw` ITSELF: a subclass with a constructor of its own does not inherit it)
    can = lambda i: (workers > 0 and not (hasattr(self.trajs[i], "xyz") and hasattr(self.trajs[i], "top"))
                     and "read_raw" in vars(self.OUTPUT_TYPE[i]))
    if not any(can(i) for i in indices):
      for i in indices:
        yield i, self._load(i)
      return
    from concurrent.futures import ThreadPoolExecutor
    ahead = max(1, int(ahead) if ahead is not None else 2 * int(workers))
    with ThreadPoolExecutor(max_workers=int(workers), thread_name_prefix="nearl-loader") as pool:
      futures, submitted = {}, 0
      try:
        for k, i in enumerate(indices):
          while submitted < len(indices) and submitted <= k + ahead:
            j = indices[submitted]
            if can(j) and submitted not in futures:
              options = self.get_loading_options(j)
              futures[submitted] = pool.submit(self.OUTPUT_TYPE[j].read_raw, *self.trajs[j],
                                               **{o: v for o, v in options.items() if o != "identity"})
            submitted += 1
          fut = futures.pop(k, None)
          if fut is None:
            yield i, self._load(i)
            continue
          raw = fut.result()
          if raw is None:                       # not file-backed after all
            yield i, self._load(i)
            continue
          self.i_ = i
          yield i, self.OUTPUT_TYPE[i](*self.trajs[i], _raw=raw, **self.get_loading_options(i))
      finally:
        for fut in futures.values():
          fut.cancel()

  def __len__(self):
    return len(self.trajs)

  def __getitem__(self, index):
    if isinstance(index, (int, np.integer)):
      return self._load(index + len(self) if index < 0 else index)
    if isinstance(index, (list, tuple, np.ndarray)):
      return [self._load(i) for i in np.array(index, dtype=int)]
    if isinstance(index, slice):
      return [self._load(i) for i in np.arange(len(self))[index]]
    raise IndexError("Index must be either an integer or a slice")

  def __iter__(self):
    return (self._load(i) for i in range(len(self)))

  def __str__(self):
    return "\n".join(str(t) for t in self).strip("\n")

  def append(self, trajs=None, trajtype=None, **kwarg):
    """Append trajectories (a list of constructor-argument tuples) of type `trajtype`, with optional ``trajids``,
    ``strides`` and ``masks`` lists of the same length (trajloader.py:213-247)."""
    if not isinstance(trajs, (list, tuple)):
      raise ValueError(f"The input should be a list or tuple of trajectory arguments rather than {type(trajs)}")
    if len(trajs) == 0:
      raise ValueError("No trajectory detected in the input")
    for idx, item in enumerate(trajs):
      self.trajs.append(item if (hasattr(item, "xyz") and hasattr(item, "top")) else (tuple(item) if isinstance(item, (list, tuple)) else (item,)))
      self.OUTPUT_TYPE.append(trajtype or Trajectory)
      self.trajids.append(kwarg["trajids"][idx] if kwarg.get("trajids") is not None else None)
      self.strides.append(kwarg["strides"][idx] if kwarg.get("strides") is not None else None)
      self.masks.append(kwarg["masks"][idx] if kwarg.get("masks") is not None else None)

  def push(self, item):
    self.append([item])
