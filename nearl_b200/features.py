"""Feature classes of the hot path, with the reference's plug-in protocol.

Every feature is ``hook(featurizer)`` -> ``cache(trajectory)`` -> ``query(topology, frames,
focal)`` -> ``run(*queried)`` -> ``dump(result)`` exactly as in the reference
(/root/reference/nearl/features.py:144-186), so user subclasses written against Nearl keep
working through the per-task host protocol.  The built-in classes additionally describe
themselves to the featuriser (``_native_spec``) so that their focus-box crop, frame slicing,
voxelisation / observation and time aggregation run batched on the device
(``nrl_featurize_batch``) instead of one host round trip per task.

In scope: ``Feature``, ``AtomicNumber``, ``Mass``, ``HeavyAtom``,
``Selection``, ``Backbone``, ``AtomType``, ``Electronegativity``, ``Hydrophobicity``,
``DynamicFeature``, ``DensityFlow``, ``MarchingObservers``, ``crop``, ``cache_properties``.
The chemistry-perception features (OpenBabel / RDKit / ChargeFW2) and label features are not
part of this path.
"""
import logging

import numpy as np

from . import _lib, commands, constants, utils
from . import printit, config
from .io import Trajectory

logger = logging.getLogger(__name__)

__all__ = [
  "Feature", "AtomicNumber", "Mass", "HeavyAtom", "Selection", "Backbone", "AtomType", "Electronegativity",
  "Hydrophobicity", "PartialCharge", "DynamicFeature", "DensityFlow", "MarchingObservers", "cache_properties", "crop",
  "LabelIdentity", "LabelAffinity", "LabelStepping", "LabelRMSD", "LabelPCDT",
  "SUPPORTED_FEATURES", "SUPPORTED_AGGREGATION", "SUPPORTED_OBSERVATION",
]

# ids are API: they are the integers the native layer dispatches on
# (features.py:56-89, src/constants.h:15-48)
SUPPORTED_FEATURES = {
  "atomic_id": 1, "residue_id": 2, "atomic_number": 3, "hybridization": 4,
  "mass": 11, "radius": 12, "electronegativity": 13, "hydrophobicity": 14, "partial_charge": 15, "uniformed": 16,
  "heavy_atom": 21, "aromaticity": 22, "ring": 23, "hbond_donor": 24, "hbond_acceptor": 25,
  "backboneness": 26, "sidechainness": 27, "atom_type": 28,
}
SUPPORTED_AGGREGATION = {
  "mean": 1, "standard_deviation": 2, "median": 3, "variance": 4, "max": 5, "min": 6,
  "information_entropy": 7, "drift": 8,
}
SUPPORTED_OBSERVATION = {
  "existence": 1, "direct_count": 2, "distinct_count": 3, "mean_distance": 11, "cumulative_weight": 12,
  "density": 13, "dispersion": 14, "eccentricity": 15, "radius_of_gyration": 16,
}

_NUMBER_TYPES = (int, float, np.integer, np.floating)


_CELL_FIT_MEMO = {}
_BUILTIN_MEMO = {}


def _lut(table):
  arr = np.full(max(table) + 1, np.nan, dtype=np.float64)
  for z, v in table.items():
    arr[z] = v
  return arr


_MASS_LUT = _lut(constants.ATOMICMASS)
_ELECNEG_LUT = _lut(constants.ELECNEG)


def _lookup(lut, table, z):
  """`np.array([table[i] for i in z], float32)` without the Python loop (KeyError preserved)."""
  z = np.asarray(z, dtype=int)
  if len(z) and (z.min() < 0 or z.max() >= len(lut) or np.isnan(lut[z]).any()):
    return np.array([table[i] for i in z], dtype=np.float32)     # raises the reference's KeyError
  return lut[z].astype(np.float32)


def atom_table(top):
  """(index, resid, atomic_number, name) columns of a topology, memoised on the topology object.
  Uses column arrays when the topology has them (nearl_b200.io.Topology), else iterates
  ``top.atoms`` like the reference (features.py:405-407,1133-1134)."""
  cached = getattr(top, "_nearl_atom_table", None)
  if cached is not None:
    return cached
  if all(hasattr(top, a) for a in ("resids", "atomic_numbers", "names")):
    table = {"index": np.arange(top.n_atoms), "resid": np.asarray(top.resids, dtype=int),
             "atomic_number": np.asarray(top.atomic_numbers, dtype=int), "name": np.asarray(top.names, dtype=object)}
  else:
    atoms = list(top.atoms)
    table = {"index": np.array([a.index for a in atoms], dtype=int), "resid": np.array([a.resid for a in atoms], dtype=int),
             "atomic_number": np.array([a.atomic_number for a in atoms], dtype=int),
             "name": np.array([a.name for a in atoms], dtype=object)}
  table["is_backbone"] = np.isin(table["name"], constants.BACKBONE_ATOM_NAMES)
  try:
    top._nearl_atom_table = table
  except Exception:
    pass
  return table


def crop(points, upperbound, padding, spacing):
  """Boolean mask of the points strictly inside the padded box, shifted by half a voxel
  (features.py:92-123): ``-padding - spacing/2 < p < upperbound + padding - spacing/2``."""
  inside = np.ones(len(points), dtype=bool)
  for axis in range(3):
    inside &= points[:, axis] < upperbound[axis] + padding - spacing / 2
    inside &= points[:, axis] > 0 - padding - spacing / 2
  return inside


class Feature:
  """Base class: a per-atom weight voxelised with a Gaussian around a focal point."""

  _native_kind = "static"

  def __init__(self, dims=None, spacing=None, outfile=None, outkey=None, cutoff=None, sigma=None, padding=None,
               byres=None, outshape=None, force_recache=None, selection=None, translate="origin", frame_offset=None,
               **kwargs):
    self._dims = None
    self._spacing = None
    self._center = None
    self._lengths = None
    self._padding = None
    self._frame_offset = frame_offset
    if dims is not None:
      self.dims = dims
    if spacing is not None:
      self.spacing = spacing
    if padding is not None:
      self.padding = padding
    self.cutoff = cutoff
    self.sigma = sigma
    self.outfile = outfile
    self.outkey = outkey
    self.outshape = outshape
    self.byres = byres
    self.selection = selection
    self.selected = None
    self.cached_array = None
    self.resids = None
    self.atomic_numbers = None
    self.classname = self.__class__.__name__
    self.force_recache = force_recache is not None
    self.translate = translate       # accepted and reported only, as in the reference (features.py:229-235)
    if translate not in ("origin", "center"):
      logger.warning(f"{self.classname}: The translate parameter is not recognized. ")
    self.hdf_compress_level = kwargs.get("hdf_compress_level", 0)
    self.hdf_dump_opts = {}
    self.PARAMSPACE = {
      "dims": self.dims, "spacing": self.spacing, "cutoff": self.cutoff, "sigma": self.sigma, "outfile": self.outfile,
      "padding": self.padding, "byres": self.byres, "selection": self.selection, "translate": self.translate,
      "force_recache": self.force_recache, "outshape": self.outshape, "HDF_OPTS": self.hdf_dump_opts,
    }

  def __str__(self):
    return (f"{self.classname} <dims:{self.dims}|lengths:{self._lengths}|spacing:{self.spacing}|cutoff:{self.cutoff}"
            f"|sigma:{self.sigma}|selection:{self.selection}>")

  # ---- geometry -------------------------------------------------------------------------------
  @property
  def params(self):
    return self.PARAMSPACE

  @property
  def dims(self):
    return self._dims

  @dims.setter
  def dims(self, value):
    if isinstance(value, _NUMBER_TYPES):
      self._dims = np.array([int(value)] * 3)
    elif isinstance(value, (list, tuple, np.ndarray)) and np.ndim(value) == 1:
      self._dims = np.array([int(i) for i in value][:3])
    elif isinstance(value, np.ndarray) and np.ndim(value) == 0:
      self._dims = np.array([int(value)] * 3)      # Featurizer({"dimensions": 32}) hands over a 0-d array
    else:
      logger.warning(f"{self.classname} The dimension should be a number, list, tuple or a numpy array, not {type(value)}")
      self._dims = None
    self._refresh_geometry()

  @property
  def spacing(self):
    return self._spacing

  @spacing.setter
  def spacing(self, value):
    if isinstance(value, _NUMBER_TYPES):
      self._spacing = float(value)
    else:
      logger.warning(f"{self.classname}: The spacing is not a valid number, setting to None")
      self._spacing = None
    self._refresh_geometry()

  def _refresh_geometry(self):
    if self._dims is not None and self._spacing is not None:
      self._center = np.array(self._dims * self._spacing, dtype=np.float32) / 2      # features.py:312
      self._lengths = self._dims * self._spacing                                       # float64

  @property
  def center(self):
    return self._center

  @property
  def lengths(self):
    return self._lengths

  @property
  def padding(self):
    return self._padding

  @padding.setter
  def padding(self, value):
    self._padding = float(value)

  @property
  def frame_offset(self):
    return self._frame_offset

  @frame_offset.setter
  def frame_offset(self, value):
    self._frame_offset = int(value)

  # ---- protocol -------------------------------------------------------------------------------
  def hook(self, featurizer):
    """Inherit unset parameters from the featuriser (features.py:349-394)."""
    if config.verbose():
      printit(f"{self}: Hooking feature ({self}) to the featurizer. ")
    if self.dims is None and featurizer.dims is not None:
      self.dims = featurizer.dims
    if self.spacing is None and featurizer.spacing is not None:
      self.spacing = featurizer.spacing
    for key in constants.COMMON_FEATURE_PARMS:
      if getattr(self, key, None) is None:
        inherited = featurizer.FEATURE_PARMS.get(key, None)
        if inherited is not None:
          setattr(self, key, inherited)
          self.PARAMSPACE[key] = inherited
    if self.padding is None and self.cutoff is not None:
      self.padding = self.cutoff
    if self.frame_offset is None:
      self.frame_offset = 0
    if getattr(self, "hdf_compress_level", 0):
      self.hdf_dump_opts["compression_opts"] = self.hdf_compress_level
      self.hdf_dump_opts["compression"] = "gzip"
      self.PARAMSPACE["HDF_OPTS"] = dict(self.hdf_dump_opts)

  def cache(self, trajectory):
    """Per-trajectory atom tables and the selection mask (features.py:397-423)."""
    table = atom_table(trajectory.top)
    self.resids = table["resid"]
    self.atomic_numbers = table["atomic_number"]
    n = len(self.resids)
    if self.selection is None:
      self.selected = np.full(n, True, dtype=bool)
    elif isinstance(self.selection, str):
      self.selected = np.full(n, False, dtype=bool)
      self.selected[trajectory.top.select(self.selection)] = True
    elif isinstance(self.selection, (list, tuple, np.ndarray)):
      if len(self.selection) != n:
        logger.warning(f"{self.classname}: The length of selection array does not match the number of atoms {n} in the topology. ")
      self.selected = np.full(n, False, dtype=bool)
      self.selected[np.array(self.selection)] = True

  def query(self, topology, frame_coords, focal_point):
    """Host protocol: translate to the focus box, crop, apply by-residue and selection
    (features.py:426-483).  Returns (mask over all atoms, coordinates of the kept atoms)."""
    focal_point = np.asarray(focal_point)
    assert focal_point.ndim == 1, f"The focal point should be a 1D array with length 3, rather than the given {focal_point.shape}"
    if self.resids is None or len(self.resids) != topology.n_atoms or self.force_recache:
      coords = np.asarray(frame_coords)
      self.cache(Trajectory(coords if coords.ndim == 3 else coords[None], topology))
    if self.center is None or self.lengths is None or self.padding is None:
      logger.warning(f"{self} Skipping the coordinates cropping due to the missing center, lengths or padding information")
      return np.full(topology.n_atoms, True, dtype=bool), frame_coords
    frame_coords = frame_coords - focal_point + self.center - self.spacing / 2
    mask = crop(frame_coords, self.lengths, self.padding, self.spacing)
    if self.byres:
      final_mask = np.isin(self.resids, np.unique(self.resids[mask]))
    else:
      final_mask = mask
    if self.selection is not None:
      final_mask = final_mask * self.selected
    return final_mask, np.ascontiguousarray(frame_coords[final_mask])

  def run(self, coords, weights):
    """Host protocol: Gaussian voxelisation of the queried atoms (features.py:485-508)."""
    if len(coords) == 0:
      logger.warning(f"{self.classname}: The coordinates from query method are empty, returning an all-zero array")
      return np.zeros(self.dims, dtype=np.float32)
    return commands.frame_voxelize(coords, weights, self.dims, self.spacing, self.cutoff, self.sigma)

  def dump(self, result):
    """NaN -> 0, then append one row to dataset ``outkey`` of ``outfile`` (features.py:510-534)."""
    if self.outfile is None or self.outkey is None or len(self.outfile) == 0:
      logger.warning(f"{self.classname}: The outfile and/or outkey are not set, the result is not dumped into file. ")
      return
    result = np.asarray(result)
    if np.isnan(result).any():
      logger.warning(f"{self.classname}: Found {np.isnan(result).sum()} NaN values in the result array")
      result = np.nan_to_num(result).astype(result.dtype)
    row = np.array([result], dtype=np.float32)
    if self.outshape is not None:
      utils.append_hdf_data(self.outfile, self.outkey, row, dtype=np.float32, maxshape=self.outshape, chunks=True,
                            compression="gzip", compression_opts=self.hdf_compress_level)
    elif len(self.dims) == 3:
      opts = dict(self.hdf_dump_opts) if self.hdf_compress_level else {}
      utils.append_hdf_data(self.outfile, self.outkey, row, dtype=np.float32, maxshape=(None, *self.dims),
                            chunks=(1, *self.dims), **opts)

  # ---- device fast path -------------------------------------------------------------------------
  # per-topology memo of everything cache() derives (a loader cycling over a few topologies pays once each)
  def _remember(self, top):
    memo = self.__dict__.setdefault("_top_memo", {})
    if len(memo) >= 64:
      memo.clear()
    memo[id(top)] = (top, self.resids, self.atomic_numbers, self.selected, self.cached_array,
                     getattr(self, "_value_order_memo", None))

  def _recall(self, top):
    if self.force_recache:
      return False
    hit = self.__dict__.get("_top_memo", {}).get(id(top))
    if hit is None or hit[0] is not top:
      return False
    _, self.resids, self.atomic_numbers, self.selected, self.cached_array, order = hit
    if order is not None:
      self._value_order_memo = order
    return True

  def value_order(self):
    """Stable argsort of the cached weights (equal values adjacent, NaN last, -0.0 == +0.0): the
    emission order of the crop kernel for distinct_count.  Memoised per weight array."""
    memo = getattr(self, "_value_order_memo", None)
    if memo is None or memo[0] is not self.cached_array:
      order = np.argsort(np.asarray(self.cached_array, dtype=np.float32) + np.float32(0.0), kind="stable").astype(np.int32)
      self._value_order_memo = memo = (self.cached_array, order)
      hit = self.__dict__.get("_top_memo", {})
      for k, v in list(hit.items()):
        if v[4] is self.cached_array:
          hit[k] = v[:5] + (memo,)
    return memo[1]

  def value_ranks(self):
    """Rank of every atom's weight among the distinct weight values, as float32 (-1 for the 99999.0
    placeholder, which distinct_count never counts: marching_observers.cu:118-138), or None when there are
    more than 32 distinct values.  -0.0 == +0.0; every NaN is a value of its own.  Memoised per weight array."""
    memos = self.__dict__.setdefault("_value_ranks_memos", {})
    memo = memos.get(id(self.cached_array))
    if memo is None or memo[0] is not self.cached_array:
      if len(memos) >= 64:
        memos.clear()
      w = np.asarray(self.cached_array, dtype=np.float32) + np.float32(0.0)
      order = self.value_order()
      sv = w[order]
      fresh = np.ones(len(sv), dtype=bool)
      fresh[1:] = ~(sv[1:] == sv[:-1])
      rank_sorted = np.cumsum(fresh) - 1
      ranks = None
      if len(sv) == 0 or rank_sorted[-1] < 32:
        ranks = np.empty(len(sv), dtype=np.float32)
        ranks[order] = rank_sorted
        ranks[w == np.float32(99999.0)] = -1.0
      memos[id(self.cached_array)] = memo = (self.cached_array, ranks)
    return memo[1]

  def _slice_invariant(self):
    """True when the cached per-atom values of an atom do not depend on which slice of the topology the
    feature was cached on (Featurizer.loop_by_residue caches on residue slices): table lookups by element /
    atom name are, positions inside the slice (atomic_id, residue_id) and number-based masks are not."""
    if isinstance(self.selection, str) and any(ch.isdigit() for ch in self.selection):
      return False
    if self.selection is not None and not isinstance(self.selection, str):
      return False
    return True

  def _is_builtin(self, *methods):
    """True when none of `methods` was overridden outside this module (a user subclass that
    customises query/run must go through the host protocol)."""
    key = (type(self), methods)
    hit = _BUILTIN_MEMO.get(key)
    if hit is None:
      hit = _BUILTIN_MEMO[key] = all(getattr(type(self), m).__module__ == __name__ for m in methods)
    return hit

  def _native_spec(self):
    """Description of this feature for nrl_featurize_batch, or None for the host protocol."""
    if self._native_kind != "static" or not self._is_builtin("query", "run"):
      return None
    if self.cached_array is None or self.center is None or self.padding is None or self.cutoff is None or self.sigma is None:
      return None
    weights = np.asarray(self.cached_array, dtype=np.float32)
    if self.selection is not None:
      # For the voxeliser an unselected atom (dropped by query, features.py:479-480) and a zero-weight atom
      # (skipped by the kernel, voxelize.cu:63) are the same thing: fold the selection into the weights so
      # that the feature shares its focus box -- and its sweep -- with the other static channels.
      memo = getattr(self, "_masked_weights", None)
      if memo is None or memo[0] is not self.cached_array or memo[1] is not self.selected:
        self._masked_weights = memo = (self.cached_array, self.selected, np.where(self.selected, weights, np.float32(0.0)).astype(np.float32))
      weights = memo[2]
    return {
      "kind": "static", "mode": 0, "agg": 0, "weights": weights, "selection_folded": self.selection is not None,
      "frame_offset": int(self.frame_offset or 0), "max_atoms": 0, "sigma": float(self.sigma),
    }


class _PerAtomWeight(Feature):
  """Static features: one cached weight per atom, first frame (``frame_offset``) of the window."""

  def _weights(self, trajectory):
    raise NotImplementedError

  def cache(self, trajectory):
    if self._recall(trajectory.top):
      return                                   # same topology object: the per-atom weights cannot have changed
    super().cache(trajectory)
    self.cached_array = np.asarray(self._weights(trajectory), dtype=np.float32)
    self._remember(trajectory.top)

  def query(self, topology, frame_coords, focal_point):
    frame_coords = np.asarray(frame_coords)
    if frame_coords.ndim == 3:
      frame_coords = frame_coords[self.frame_offset]
    mask, coords = super().query(topology, frame_coords, focal_point)
    return coords, self.cached_array[mask]


class AtomicNumber(_PerAtomWeight):
  """features.py:538-560"""
  def _weights(self, trajectory):
    return self.atomic_numbers


class Mass(_PerAtomWeight):
  """features.py:563-605"""
  def _weights(self, trajectory):
    return _lookup(_MASS_LUT, constants.ATOMICMASS, self.atomic_numbers)


class HeavyAtom(_PerAtomWeight):
  """features.py:608-637"""
  def __init__(self, default_weight=1, **kwargs):
    super().__init__(**kwargs)
    self.default_weight = default_weight

  def _weights(self, trajectory):
    return np.where(self.atomic_numbers > 1, self.default_weight, 0)


class Selection(_PerAtomWeight):
  """features.py:724-754: `default_value` on the selected atoms; unselected atoms are dropped."""
  def __init__(self, default_value=1, **kwargs):
    if "selection" not in kwargs:
      raise ValueError("Selection: The selection parameter should be set")
    super().__init__(**kwargs)
    self.default_value = default_value

  def _weights(self, trajectory):
    return np.where(self.selected, self.default_value, 0)


class Backbone(_PerAtomWeight):
  """features.py:826-857"""
  def __init__(self, reverse=False, **kwargs):
    super().__init__(**kwargs)
    self.reverse = reverse

  def _weights(self, trajectory):
    bb = atom_table(trajectory.top)["is_backbone"]
    return (~bb if self.reverse else bb).astype(np.float32)


class AtomType(_PerAtomWeight):
  """features.py:860-892"""
  def __init__(self, focus_element, **kwargs):
    super().__init__(**kwargs)
    self.focus_element = int(focus_element)

  def __str__(self):
    return f"{self.classname} <focus:{constants.ATOMICNR2SYMBOL.get(self.focus_element, self.focus_element)}>"

  def _weights(self, trajectory):
    return (self.atomic_numbers == self.focus_element).astype(np.float32)


class Electronegativity(_PerAtomWeight):
  """features.py:1033-1046"""
  def _weights(self, trajectory):
    return _lookup(_ELECNEG_LUT, constants.ELECNEG, self.atomic_numbers)


class Hydrophobicity(_PerAtomWeight):
  """features.py:1049-1067: |EN - EN(carbon)| evaluated in float32"""
  def _weights(self, trajectory):
    en = _lookup(_ELECNEG_LUT, constants.ELECNEG, self.atomic_numbers)
    return np.abs(en - constants.ELECNEG[6])


class PartialCharge(_PerAtomWeight):
  """Partial charge of every atom (features.py:895-1031): ``charge_type`` "topology" (the charges of the topology, an error
  when they are all zero), "manual" (``charge_parm``: {trajectory identity: charges}) or "external" (``charge_parm``: a
  function of the trajectory); ``keep_sign`` "positive" / "p" keeps max(q, 0), "negative" / "n" min(q, 0), "both" / "b"
  everything.  "chargefw" needs ChargeFW2, which is outside this path.  Negative weights deposit negative density; atoms
  of charge exactly 0 are skipped by the voxeliser like any zero weight (voxelize.cu:63)."""

  def __init__(self, charge_type="topology", charge_parm=None, keep_sign="both", **kwargs):
    super().__init__(**kwargs)
    if charge_type not in ("topology", "manual", "chargefw", "external"):
      raise ValueError(f"{self.classname}: The charge type should be either 'topology', 'manual', 'chargefw' or 'external' rather than {charge_type}")
    if keep_sign not in ("positive", "negative", "both", "p", "n", "b"):
      raise ValueError(f"{self.classname}: The keep_sign parameter should be either positive/p, negative/n or both/b rather than {keep_sign}")
    self.charge_type, self.charge_parm = charge_type, charge_parm
    self.keep_sign = {"p": "positive", "n": "negative", "b": "both"}.get(keep_sign, keep_sign)

  def __str__(self):
    return f"{self.classname} <type:{self.charge_type}|sign:{self.keep_sign}>"

  def cache(self, trajectory):
    if self.charge_type != "topology":
      self.force_recache = True            # the charges belong to the trajectory (its identity, a user function), not to the topology object
    super().cache(trajectory)

  def _weights(self, trajectory):
    if self.charge_type == "topology":
      q = np.asarray(getattr(trajectory.top, "charge", np.zeros(trajectory.top.n_atoms)), dtype=np.float64)
      if np.allclose(q, 0):
        raise ValueError(f"{self.classname}: The charge values in the topology are all zero meaning the topology of the trajectory does not contain necessary charge information. ")
    elif self.charge_type == "manual":
      q = np.array(self.charge_parm[trajectory.identity], dtype=np.float32)
    elif self.charge_type == "external":
      q = np.asarray(self.charge_parm(trajectory))
    else:
      raise NotImplementedError(f"{self.classname}: charge_type 'chargefw' needs ChargeFW2 (chargefw2_python), which is outside this path")
    if len(q) != trajectory.top.n_atoms:
      raise AssertionError(f"{self.classname}: The number of charge values does not match the number of atoms in the trajectory. ")
    if self.keep_sign == "positive":
      q = np.maximum(q, 0)
    elif self.keep_sign == "negative":
      q = np.minimum(q, 0)
    return q


def cache_properties(trajectory, property_type, **kwargs):
  """Per-atom weights of the dynamic features by id (features.py:1073-1230).  The ids whose
  reference implementation depends on OpenBabel / RDKit / ChargeFW2 or is broken in the
  reference snapshot (4, 12, 15, 22-25) raise NotImplementedError."""
  table = atom_table(trajectory.top)
  z = table["atomic_number"]
  if property_type == 1:
    arr = table["index"]
  elif property_type == 2:
    arr = table["resid"]
  elif property_type == 3:
    arr = z
  elif property_type == 11:
    arr = _lookup(_MASS_LUT, constants.ATOMICMASS, z)
  elif property_type == 13:
    arr = _lookup(_ELECNEG_LUT, constants.ELECNEG, z)
  elif property_type == 14:
    arr = np.abs(_lookup(_ELECNEG_LUT, constants.ELECNEG, z) - constants.ELECNEG[6])
  elif property_type == 16:
    arr = np.full(len(z), kwargs.get("manual_weight", 1), dtype=np.float32)
  elif property_type == 21:
    arr = (z > 1)
  elif property_type in (26, 27):
    # 26 "backboneness" yields the side-chain indicator in the reference too (features.py:1212-1215)
    arr = ~table["is_backbone"]
  elif property_type == 28:
    if "element_type" not in kwargs:
      raise ValueError("The focus element should be set for the atom type")
    arr = (z == kwargs["element_type"])
  elif property_type in (4, 12, 15, 22, 23, 24, 25):
    raise NotImplementedError(f"weight type {property_type} needs chemistry perception that is outside this path")
  else:
    raise ValueError(f"The weight type is not recognized {property_type}")
  return np.array(arr, dtype=np.float32)


class DynamicFeature(Feature):
  """Base of the frame-window features (features.py:1233-1374)."""

  _native_kind = None

  def __init__(self, agg="mean", weight_type="mass", **kwargs):
    super().__init__(**kwargs)
    self._agg_type = agg
    self._weight_type = weight_type
    self.feature_args = kwargs
    self.MAX_ALLOWED_ATOMS = 9999
    self.DEFAULT_COORD = 99999.0

  def __str__(self):
    return f"{self.classname} <agg:{self._agg_type}|weight:{self._weight_type}|cutoff:{self.cutoff}>"

  def _slice_invariant(self):
    # atomic_id / residue_id number atoms and residues inside the topology slice the feature was cached on
    return super()._slice_invariant() and self.weight_type not in (1, 2)

  @property
  def agg(self):
    if self._agg_type in SUPPORTED_AGGREGATION:
      return SUPPORTED_AGGREGATION[self._agg_type]
    raise ValueError("The aggregation type is not recognized")

  @agg.setter
  def agg(self, value):
    assert isinstance(value, str), "The aggregation type should be a string"
    self._agg_type = value

  @property
  def weight_type(self):
    if self._weight_type in SUPPORTED_FEATURES:
      return SUPPORTED_FEATURES[self._weight_type]
    raise ValueError(f"The weight type is not recognized {self._weight_type}; Available types are {SUPPORTED_FEATURES.keys()}")

  @weight_type.setter
  def weight_type(self, value):
    assert isinstance(value, str), "The weight type should be a string"
    self._weight_type = value

  def cache(self, trajectory):
    if self._recall(trajectory.top):
      return
    super().cache(trajectory)
    self.cached_array = cache_properties(trajectory, self.weight_type, **self.feature_args)
    self._remember(trajectory.top)

  def query(self, topology, frame_coords, focal_point):
    """Host protocol: per-frame crop packed into [F, max_n, 3] padded with 99999 and weights
    padded with 0, truncated to the first MAX_ALLOWED_ATOMS atoms (features.py:1329-1368)."""
    frame_coords = np.asarray(frame_coords)
    assert frame_coords.ndim == 3, f"{self.classname}: The coordinates should follow the convention (frames, atoms, 3); "
    F = len(frame_coords)
    coords = np.full((F, self.MAX_ALLOWED_ATOMS, 3), self.DEFAULT_COORD, dtype=np.float32)
    weights = np.full((F, self.MAX_ALLOWED_ATOMS), 0.0, dtype=np.float32)
    max_n = 0
    for idx, frame in enumerate(frame_coords):
      mask, inbox = Feature.query(self, topology, frame, focal_point)
      n = np.count_nonzero(mask)
      if n > self.MAX_ALLOWED_ATOMS:
        logger.warning(f"{self.classname}: The maximum allowed atom slice is {self.MAX_ALLOWED_ATOMS} but the maximum atom number is {n}")
      n = min(n, self.MAX_ALLOWED_ATOMS)
      coords[idx, :n] = inbox[:n]
      weights[idx, :n] = self.cached_array[mask][:n]
      max_n = max(max_n, n)
    return (np.ascontiguousarray(coords[:, :max_n], dtype=np.float32),
            np.ascontiguousarray(weights[:, :max_n].flatten(), dtype=np.float32))

  def run(self, frames, weights):
    raise NotImplementedError("The run function should be implemented in the child class")

  def _dynamic_spec(self, mode, sigma):
    if not self._is_builtin("query", "run"):
      return None
    if self.cached_array is None or self.center is None or self.padding is None or self.cutoff is None:
      return None
    weights = np.asarray(self.cached_array, dtype=np.float32)
    folded = False
    if int(mode) == 0 and self.selection is not None and (self.MAX_ALLOWED_ATOMS == 0 or len(weights) <= self.MAX_ALLOWED_ATOMS):
      # A density flow deposits weight x Gaussian: an unselected atom and a zero-weight atom (skipped by the kernel,
      # voxelize.cu:63) are the same thing, as for the static channels -- fold the selection into the weights so that
      # the flows of several selections share ONE focus box and ONE sweep.  Only while the atom cap of the dynamic
      # query (features.py:1353-1359) cannot bite: beyond it the cap counts selected atoms only.
      memo = getattr(self, "_masked_weights", None)
      if memo is None or memo[0] is not self.cached_array or memo[1] is not self.selected:
        self._masked_weights = memo = (self.cached_array, self.selected, np.where(self.selected, weights, np.float32(0.0)).astype(np.float32))
      weights, folded = memo[2], True
    spec = {
      "kind": "dynamic", "mode": int(mode), "agg": int(self.agg), "weights": weights, "selection_folded": folded,
      "frame_offset": 0, "max_atoms": int(self.MAX_ALLOWED_ATOMS), "sigma": float(sigma),
    }
    if self._is_builtin("cache") and self.selection is None:
      # features with the same signature cache the same per-atom weights: they share one weight table (and one sweep)
      spec["weight_sig"] = ("dynamic", self.weight_type, repr(self.feature_args.get("manual_weight", 1)),
                            repr(self.feature_args.get("element_type", None)))       # what cache_properties reads
    if int(mode) == 3:
      # distinct_count over few distinct values (atomic numbers, atom types ...) runs on the cell-list kernel with
      # per-atom value ranks; otherwise the crop is emitted in value order for the scan kernel
      gkey = (tuple(int(v) for v in self.dims), float(self.spacing), float(self.cutoff), _lib.OPTION_EPOCH)
      if gkey not in _CELL_FIT_MEMO:
        _CELL_FIT_MEMO[gkey] = _lib.cell_list_fits(*gkey[:3])
      spec["ranks"] = self.value_ranks() if _CELL_FIT_MEMO[gkey] else None
    return spec


class DensityFlow(DynamicFeature):
  """Per-frame Gaussian voxelisation reduced over the window (features.py:1377-1421)."""

  def __str__(self):
    return f"{self.classname} <agg:{self._agg_type}|weight:{self._weight_type}|cutoff:{self.cutoff}|sigma:{self.sigma}>"

  def run(self, frames, weights):
    ret = commands.density_flow(frames, weights, self.dims, self.spacing, self.cutoff, self.sigma, self.agg)
    return ret.reshape(self.dims)

  def _native_spec(self):
    if self.sigma is None:
      return None
    return self._dynamic_spec(0, self.sigma)


class MarchingObservers(DynamicFeature):
  """Every voxel observes the atoms within `cutoff` per frame, reduced over the window
  (features.py:1424-1523)."""

  def __init__(self, obs="existence", **kwargs):
    super().__init__(**kwargs)
    self._obs_type = obs

  def __str__(self):
    return f"{self.classname} <obs:{self._obs_type}|type:{self._weight_type}|agg:{self._agg_type}|cutoff:{self.cutoff}>"

  @property
  def obs(self):
    if self._obs_type in SUPPORTED_OBSERVATION:
      return SUPPORTED_OBSERVATION[self._obs_type]
    raise ValueError(f"The observation type is not recognized {self._obs_type}; Available types are {SUPPORTED_OBSERVATION.keys()}")

  @obs.setter
  def obs(self, value):
    assert isinstance(value, str), "The observation type should be a string"
    self._obs_type = value

  def run(self, coords, weights):
    ret = commands.marching_observer(coords, weights, self.dims, self.spacing, self.cutoff, self.obs, self.agg)
    return ret.reshape(self.dims)

  def _native_spec(self):
    return self._dynamic_spec(self.obs, 0.0)


# ---- label features: per-task scalars written next to the grids (features.py:1529-1775) --------------------------
class _Label(Feature):
  """A feature whose task result is a scalar, not a grid: always the host protocol, dumped as rows of a 1-D
  dataset (``outshape=(None,)``)."""
  _native_kind = "label"

  def _native_spec(self):
    return None


class LabelIdentity(_Label):
  """The identity of the trajectory as meta-data (features.py:1529-1556)."""

  def cache(self, trajectory):
    super().cache(trajectory)
    self.identity = trajectory.identity

  def query(self, *args):
    return (self.identity, )

  def run(self, identity):
    return identity


class LabelAffinity(_Label):
  """The affinity of the complex (trajectory identity -> PDB code -> column `colname` of a PDBBind-style csv
  table with a `pdbcode` column), features.py:1559-1634."""

  def __init__(self, baseline_map, colname="pK", **kwargs):
    kwargs.setdefault("outshape", (None,))
    super().__init__(**kwargs)
    import pandas as pd
    self.baseline_table = pd.read_csv(baseline_map, header=0, delimiter=",")
    self.base_value = None
    self.colname = colname
    if self.colname not in self.baseline_table.columns:
      raise ValueError(f"The column name {self.colname} is not found in the baseline table")
    self._lookup = None

  def search_baseline(self, pdbcode):
    pdbcode = utils.get_pdbcode(pdbcode)
    if self._lookup is None:       # first match wins, as row[...].values[0] does
      codes, values = self.baseline_table["pdbcode"].values, self.baseline_table[self.colname].values
      self._lookup = {}
      for code, value in zip(codes, values):
        self._lookup.setdefault(code, value)
    if pdbcode in self._lookup:
      return self._lookup[pdbcode]
    raise ValueError(f"Cannot find the baseline value for {pdbcode}")

  def cache(self, trajectory):
    self.base_value = self.search_baseline(trajectory.identity)
    if config.verbose():
      printit(f"{self.classname}: Affinity value of {trajectory.identity} is {self.base_value}")

  def query(self, *args):
    return (self.base_value, )

  def run(self, affinity_val):
    return affinity_val


class LabelStepping(LabelAffinity):
  """Affinity as a class: floor(pK / 2) capped at 5 (features.py:1637-1645)."""

  def query(self, *args):
    assert self.base_value is not None, "The base value should be set before the query"
    return (min(self.base_value // 2, 5), )


class LabelRMSD(_Label):
  """base_value scaled down by the mean no-fit RMSD of the window against the first frame of the trajectory
  (features.py:1648-1676).  As in the reference the class needs a `selection`."""

  def __init__(self, base_value=0, outshape=(None,), **kwargs):
    if "selection" not in kwargs:
      raise ValueError("The selection should be set for the RMSD label feature")
    super().__init__(outshape=outshape, **kwargs)
    self.base_value = float(base_value)

  def cache(self, trajectory):
    super().cache(trajectory)
    self.refframe = trajectory[0]
    self._index = np.flatnonzero(self.selected)
    self.cached_array = np.minimum(utils.rmsd_nofit(trajectory.xyz, self._index, self.refframe.xyz), 8)

  def query(self, topology, frames, focus):
    rmsd = np.mean(utils.rmsd_nofit(frames, self._index, self.refframe.xyz))
    return (self.base_value * (1 - 0.1 * (rmsd / 8.0)), )

  def run(self, value):
    return value


class LabelPCDT(LabelAffinity):
  """Affinity with a conformation penalty: the z-score of the window's pairwise-closest-distance profile against the
  whole trajectory's (features.py:1679-1775)."""

  def __init__(self, selection=None, search_cutoff=None, **kwargs):
    super().__init__(**kwargs)
    self.selection_prototype = selection
    self.selected = None
    self.selected_counterpart = None
    self.refframe = None
    self.cached_array = None
    self.search_cutoff = search_cutoff

  def cache(self, trajectory):
    super().cache(trajectory)
    if isinstance(self.selection_prototype, str):
      selected = trajectory.top.select(self.selection_prototype)
    elif isinstance(self.selection_prototype, (list, tuple, np.ndarray)):
      selected = np.array([int(i) for i in self.selection_prototype])
    else:
      raise ValueError("The selection should be either a string or a list of atom indices")
    self.refframe = trajectory[0]
    self.selected = selected
    backbone_cb = trajectory.top.select("@C, @N, @CA, @O, @CB")
    ref = self.refframe.xyz
    # counterpart of every selected atom: the closest backbone / CB atom in the reference frame (one distance matrix
    # instead of the reference's loop over atoms; the same differences, squares, sums and square roots)
    self.selected_counterpart = backbone_cb[np.argmin(utils._pair_distances(ref[selected], ref[backbone_cb]), axis=1)]
    self.cached_array = utils.compute_pcdt(trajectory, self.selected, self.selected_counterpart, ref=self.refframe)
    # compute_pcdt pairs atom i with the member of the counterpart LIST closest to it in the reference frame: the pairs do
    # not depend on the window, so query() re-uses them
    self._closest = self.selected_counterpart[np.argmin(utils._pair_distances(ref[self.selected], ref[self.selected_counterpart]), axis=1)]
    if self.search_cutoff is not None:
      self.cached_array = np.minimum(self.cached_array, self.search_cutoff)
    self._xyz = trajectory.xyz
    self._mean_0 = np.mean(self.cached_array, axis=1)
    std_0 = np.std(self.cached_array, axis=1)
    if np.any(std_0 == 0):
      std_0[std_0 == 0] = self.search_cutoff if self.search_cutoff is not None else 6.0
    self._std_0 = std_0

  def _window_start(self, frames):
    """Index of the first frame when `frames` is a view of consecutive frames of the cached trajectory, else None."""
    xyz = getattr(self, "_xyz", None)
    if not (isinstance(frames, np.ndarray) and isinstance(xyz, np.ndarray)) or frames.ndim != 3 or frames.dtype != xyz.dtype:
      return None
    if frames.shape[1:] != xyz.shape[1:] or frames.strides != xyz.strides or len(frames) == 0:
      return None
    offset = frames.__array_interface__["data"][0] - xyz.__array_interface__["data"][0]
    if offset < 0 or offset % xyz.strides[0]:
      return None
    first = offset // xyz.strides[0]
    return int(first) if first + len(frames) <= len(xyz) else None

  def query(self, topology, frames, focus):
    first = self._window_start(frames)
    if first is not None:
      # the window is a run of frames of the cached trajectory: its pair distances are columns of the cached profile
      # (same pairs, same arithmetic, cutoff already applied)
      pdist_arr = self.cached_array[:, first:first + len(frames)]
    else:
      pdist_arr = utils.pair_track(np.asarray(frames), self.selected, self._closest).astype(np.float64)    # compute_pcdt on the window
      if self.search_cutoff is not None:
        pdist_arr = np.minimum(pdist_arr, self.search_cutoff)
    mean_1 = np.mean(pdist_arr, axis=1)
    penalty = abs(np.mean((self._mean_0 - mean_1) / self._std_0))
    return (self.base_value * (1 - 0.1 * penalty), )
