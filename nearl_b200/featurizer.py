"""``Featurizer``: trajectory -> frame window -> focal point -> feature loop of the reference
(/root/reference/nearl/featurizer.py:29-469) with the same registration API, attributes and
output order, where the inner three loops of every trajectory are lowered to ONE batched
native call (``nrl_featurize_batch``): focus-box transform, crop and frame slicing happen on
the device and no per-task host round trip remains.  Features the batch cannot describe
(user subclasses overriding ``query``/``run``) still run through the host protocol, task by
task, interleaved in the reference's order.
"""
import ctypes
import json
import logging
import time

import numpy as np

from . import _lib, constants, utils
from . import printit, config
from .io import pinned_empty

__all__ = ["Featurizer"]

logger = logging.getLogger(__name__)

_NUMBER_TYPES = (int, float, np.integer, np.floating)


def _as3(value, dtype):
  if isinstance(value, _NUMBER_TYPES) or np.ndim(value) == 0:
    return np.array([value] * 3, dtype=dtype)
  arr = np.array(value, dtype=dtype).reshape(-1)
  assert len(arr) == 3, "length should be 3"
  return arr


class Featurizer:
  """Automates the featurisation of many features over a batch of trajectories.

  Parameters (featurizer.py:92-162): ``dimensions`` (int or 3 ints), ``lengths`` or ``spacing``,
  ``time_window`` (default 1) and the parameters features inherit
  (``constants.COMMON_FEATURE_PARMS``).  With an ``outfile`` the parameter set is written to
  the group ``featurizer_parms`` immediately.
  """

  def __init__(self, parms={}, **kwargs):
    self._dims = None
    self._lengths = None
    self._spacing = None
    if parms.get("dimensions", None) is not None:
      self._dims = _as3(parms["dimensions"], int)
    if "lengths" in parms:
      self._lengths = _as3(parms.get("lengths", 16), float)
      if self._dims is not None:
        self._spacing = float(np.mean(self._lengths / self._dims))
    elif "spacing" in parms:
      self._spacing = float(parms.get("spacing", 1.0))
      if self._dims is not None:
        self._lengths = self._dims * self._spacing
    self.time_window = int(parms.get("time_window", 1))

    self.FEATURE_PARMS = {}
    for key in constants.COMMON_FEATURE_PARMS:
      self.FEATURE_PARMS[key] = parms[key] if key in parms else kwargs.get(key, None)
    self.OTHER_PARMS = {k: v for k, v in parms.items() if k not in constants.COMMON_FEATURE_PARMS}
    self.OTHER_PARMS.update({k: v for k, v in kwargs.items() if k not in constants.COMMON_FEATURE_PARMS})

    self._traj = None
    self.frame_slice = None
    self.FRAMENUMBER = 0
    self.SLICENUMBER = 0
    self.FRAMESLICENUMBER = 0
    self.FRAMESLICES = []
    self.FEATURESPACE = []
    self.FEATURENUMBER = 0
    self.FOCALPOINTS = []
    self.FOCALPOINTS_PROTOTYPE = None
    self.FOCALPOINTS_TYPE = None
    self.FOCALNUMBER = 0
    self.TRAJLOADER = None
    self.TRAJECTORYNUMBER = 0
    self.classname = self.__class__.__name__
    # nearl_b200: set False to force every feature through the per-task host protocol
    self.use_native_batch = True
    # nearl_b200: collect results in self.RESULTS[outkey] (list of arrays) besides / instead of dumping
    self.keep_results = False
    self.RESULTS = {}
    # nearl_b200: callable(outkey, trajectory_index, window, focal_index, grid) invoked for every result in
    # the reference's order instead of (or, with an outfile, besides) Feature.dump; grids are views into a
    # re-used page-locked buffer and must be consumed inside the call
    self.result_sink = None
    self.last_run_stats = {}
    # nearl_b200: a task-sharded run (run(task_shard=...), sharding.run_task_sharded) leaves its parts here
    self._task_shard = self._shard_device_out = self._shard_device_xyz = None
    self.SHARD_RESULTS = {}
    # nearl_b200: trajectories in flight in run(), one per lane of the native library (the uploads of the next ones and the
    # crop / binning of the one after run under the sweeps of the current one); 1 = no overlap.  Measured on B200, headline
    # workload, frames/s end to end: 3 -> 178 k, 4 -> 192 k, 5 -> 205 k, 7 / 8 -> 207 k (the host-to-device link's limit)
    self.pipeline_depth = 8
    # nearl_b200: coordinates in flight, in bytes: trajectories larger than pipeline_bytes / pipeline_depth use fewer lanes
    self.pipeline_bytes = 8 << 30
    # nearl_b200: single-frame systems featurised per native call when every feature is static (time_window 1);
    # 1 = one system per call
    self.static_group = 32
    # nearl_b200: loader threads that read the files of the next trajectories ahead (io.TrajectoryLoader.prefetched; 0: none;
    # None: half of this process's share of the host cores, between 2 and 8)
    self.prefetch_workers = None
    if config.verbose():
      printit(f"{self.classname}: Featurizer is initialized successfully with dimensions: {self.dims} and lengths: {self.lengths}")
    if "outfile" in parms:
      printit(f"{self.classname}: Dumping the parameters to {parms['outfile']} : {self.parms}")
      utils.dump_dict(parms["outfile"], "featurizer_parms", self.parms)

  def __str__(self):
    return f"Feature Number: {self.FEATURENUMBER}; \n" + "".join(f"Feature: {feat}\n" for feat in self.FEATURESPACE)

  # ---- parameters ---------------------------------------------------------------------------
  @property
  def parms(self):
    """All parameters with None replaced by 0 (featurizer.py:170-175)."""
    return {k: (v if v is not None else 0) for k, v in {**self.FEATURE_PARMS, **self.OTHER_PARMS}.items()}

  @property
  def dims(self):
    return np.array(self._dims) if self._dims is not None else None

  @dims.setter
  def dims(self, newdims):
    self._dims = _as3(newdims, int)
    if self._lengths is not None:
      self._spacing = float(np.mean(self._lengths / self._dims))

  @property
  def lengths(self):
    return self._lengths

  @lengths.setter
  def lengths(self, new_length):
    self._lengths = _as3(new_length, float)
    if self._dims is not None:
      self._spacing = float(np.mean(self._lengths / self._dims))

  @property
  def spacing(self):
    return self._spacing

  @property
  def traj(self):
    return self._traj

  @traj.setter
  def traj(self, the_traj):
    """Frame windows of the active trajectory (featurizer.py:229-241)."""
    self._traj = the_traj
    self.FRAMENUMBER = the_traj.n_frames
    self.SLICENUMBER = self.FRAMENUMBER // self.time_window
    self.FRAMESLICENUMBER = self.SLICENUMBER
    if self.SLICENUMBER == 0:
      logger.warning(f"{self.classname}: No frame slice is available. The trajectory have {self.FRAMENUMBER} frames and the time window is {self.time_window}.")
    if self.FRAMENUMBER % self.time_window != 0 and self.FRAMENUMBER != 1:
      logger.warning(f"{self.classname}: the number of frames ({self.FRAMENUMBER}) is not divisible by the time window ({self.time_window}). The last few frames will be ignored.")
    self.FRAMESLICES = [np.s_[i * self.time_window:(i + 1) * self.time_window] for i in range(self.SLICENUMBER)]

  @property
  def top(self):
    return self.traj.top

  # ---- registration -------------------------------------------------------------------------
  def register_feature(self, feature):
    feature.hook(self)
    self.FEATURESPACE.append(feature)
    self.FEATURENUMBER = len(self.FEATURESPACE)
    keys = [f.outkey for f in self.FEATURESPACE]
    if len(set(keys)) != len(keys):
      raise ValueError("The output keys for the features should be unique.")

  def register_features(self, features):
    if isinstance(features, dict):
      features = list(features.values())
    for feature in features:
      self.register_feature(feature)

  def register_trajloader(self, trajloader):
    self.TRAJLOADER = trajloader
    self.TRAJECTORYNUMBER = len(trajloader)
    printit(f"{self.classname}: Registered {self.TRAJECTORYNUMBER} trajectories")

  def register_focus(self, focus, format):
    """``mask`` (Amber masks), ``index`` (lists of atom indices), ``absolute`` ([P,3] points) or
    ``json`` (file: trajectory identity -> atom indices); featurizer.py:299-343."""
    if format == "absolute":
      focus = np.asarray(focus)
      assert focus.ndim == 2 and focus.shape[1] == 3, "The focus should be a 2D array with 3 columns"
      self.FOCALPOINTS = focus
    elif format in ("mask", "index", "json"):
      self.FOCALPOINTS = None
    else:
      raise ValueError(f"Unexpected focus format: {format}. Please choose from 'mask', 'absolute', 'index', 'function'")
    self.FOCALPOINTS_PROTOTYPE = focus
    self.FOCALPOINTS_TYPE = format

  def parse_focus(self):
    """Focal points [slice, focus, 3] in float32 for the active trajectory: the centre of
    geometry of the selection in the FIRST frame of every window (featurizer.py:345-397).
    Computed with numpy on the host trajectory, as the reference does; O(S*P*n_sel)."""
    proto = self.FOCALPOINTS_PROTOTYPE
    if self.FOCALPOINTS_TYPE == "json":
      with open(proto, "r") as fh:
        table = json.load(fh)
      key = self.traj.identity if self.traj.identity in table else str(self.traj.identity)[:4]
      proto = [np.array(table[key], dtype=int)]
    self.FOCALNUMBER = len(proto)
    self.FOCALPOINTS = np.full((self.SLICENUMBER, self.FOCALNUMBER, 3), 99999, dtype=np.float32)
    for midx, item in enumerate(proto):
      if self.FOCALPOINTS_TYPE == "mask":
        selection = self.traj.top.select(item)
        if len(selection) == 0:
          printit(f"{self.classname} Warning: The trajectory {self.traj.identity} does not have any atoms in the selection {item}")
          return False
      elif self.FOCALPOINTS_TYPE in ("index", "json"):
        selection = np.asarray(item, dtype=int)
      else:
        assert len(item) == 3, "The focus should be a 3D coordinate"
        self.FOCALPOINTS[:, midx] = item
        continue
      if self.SLICENUMBER:
        self.FOCALPOINTS[:, midx] = self._focal_means(selection)
    return 1

  def _focal_means(self, selection):
    """Centre of geometry of `selection` in the first frame of every window, [windows, 3]: the reference's per-frame
    np.mean(frame[selection], axis=0) (featurizer.py:363-365).  For a plain float32 / float64 trajectory array the native
    library adds the rows with numpy's arithmetic -- same order, same precision, same rounding, so the same bits
    (tests/test_host_api.py) -- without materialising the gathered copy; anything else takes numpy itself."""
    xyz = self.traj.xyz
    selection = np.asarray(selection)
    if (isinstance(xyz, np.ndarray) and xyz.ndim == 3 and xyz.dtype in (np.float32, np.float64) and xyz.flags.c_contiguous
        and selection.ndim == 1 and len(selection) > 0 and selection.dtype.kind in "iu" and selection.min() >= 0):
      frames = np.arange(0, self.SLICENUMBER * self.time_window, self.time_window, dtype=np.int32)
      sel = np.ascontiguousarray(selection, dtype=np.int32)
      out = np.empty((len(frames), 3), dtype=np.float32)
      _lib.check(_lib.lib().nrl_focal_means(_lib.addr(xyz), int(xyz.dtype == np.float64), xyz.shape[0], xyz.shape[1], _lib.addr(frames),
                                            len(frames), _lib.addr(sel), len(sel), _lib.addr(out)))
      return out
    # one gather: np.mean over axis 1 adds the selected rows in the same order as the per-frame np.mean(..., axis=0)
    firsts = xyz[0:self.SLICENUMBER * self.time_window:self.time_window]
    return np.mean(firsts[:, selection], axis=1)

  # ---- the loop -----------------------------------------------------------------------------
  def run(self, *_ignored, trajectory_indices=None, task_shard=None, shard_device_out=None, shard_device_xyz=None):
    """Featurise every registered trajectory (featurizer.py:399-469).  Positional arguments
    (the worker counts some scripts pass) are accepted and ignored.  `trajectory_indices`
    restricts the loop to a shard (multi-GPU: one process per GPU, see nearl_b200.sharding).  `task_shard` =
    (rank, world[, n_slabs]) shards INSIDE every trajectory instead -- this process computes its share of the (window,
    focal point) units, cut into slabs of the voxel grid when there are fewer units than ranks -- and leaves the parts in
    `self.SHARD_RESULTS` for `sharding.run_task_sharded` to collect (nothing is dumped by this call).  With
    `shard_device_out` (callable(lane, n_floats) -> device address) the parts stay in device memory, one buffer per lane,
    to be sent to rank 0 over NVLink; SHARD_RESULTS then describes them instead of holding them.  `shard_device_xyz`
    (callable(lane, xyz) -> device address of the whole trajectory) lets the caller bring the coordinates to the device
    itself (every rank uploads a share of the frames, an all-gather over NVLink does the rest).

    Trajectories are software-pipelined over `pipeline_depth` lanes of the native library: while
    the kernels of trajectory i run, the host lowers the next ones and their frames upload;
    results are dumped in trajectory order."""
    indices = list(range(self.TRAJECTORYNUMBER) if trajectory_indices is None else trajectory_indices)
    stats = {"trajectories": 0, "tasks": 0, "native_tasks": 0, "frames": 0, "seconds": 0.0}
    t_run = time.perf_counter()
    depth = max(1, min(int(self.pipeline_depth), 8))        # lanes of the native library (NRL_MAX_LANES)
    self.__dict__.pop("_fingerprint_memo", None)
    self.__dict__.pop("_spec_memo", None)
    self.__dict__.pop("_template_quick", None)
    self._task_shard = None
    if task_shard is not None:
      rank, world = int(task_shard[0]), int(task_shard[1])
      if not 0 <= rank < world:
        raise ValueError(f"task_shard: rank {rank} outside a world of {world}")
      self._task_shard = (rank, world, task_shard[2] if len(task_shard) > 2 else None)
      self._shard_device_out = shard_device_out
      self._shard_device_xyz = shard_device_xyz
      self.SHARD_RESULTS = {}
    with utils.hdf_session():                                # output files stay open until the run is over
      done = self._run_static_groups(indices, depth, stats) if (self._static_groupable() and self._task_shard is None) else 0
      self._run_pipelined(indices[done:], depth, stats)
    self._task_shard = self._shard_device_out = self._shard_device_xyz = None
    stats["seconds"] = time.perf_counter() - t_run
    self.last_run_stats = stats
    printit(f"{self.classname}: All trajectories and tasks are finished. \n")
    return stats

  def _run_pipelined(self, indices, depth, stats, loaded=None):
    """One native call per trajectory, up to `depth` of them in flight.  A lane (its page-locked result buffer, its
    table buffers, its device workspace) belongs to a job from submission until `_finish`: a job takes the lowest lane
    that no pending job holds -- a skipped trajectory takes none -- and the oldest job is finished first when none is
    free.  Large trajectories narrow the pipeline: only `pipeline_bytes // trajectory bytes` lanes are used for them, so
    that the lanes' workspaces (a few times the trajectory each, grow-only) stay bounded."""
    pending = []                                             # submitted, not yet finished, oldest first
    multi_frame = False
    for tid, traj in self._trajectories(indices, loaded):
      nbytes = int(getattr(getattr(traj, "xyz", None), "nbytes", 0) or 0)
      lanes = max(1, min(depth, int(self.pipeline_bytes) // max(nbytes, 1)))
      lane = None
      while lane is None:
        busy = {job["lane"] for job in pending}
        lane = next((k for k in range(lanes) if k not in busy), None)
        if lane is None:
          self._finish(pending.pop(0), stats)                # frees a lane (and its buffers)
      job = self._submit(tid, lane=lane, traj=traj)
      multi_frame |= self.traj.n_frames != 1
      if job is None:
        continue                                             # skipped (no focal point): the lane stays free
      pending.append(job)              # host-protocol tasks were evaluated inside _submit
    while pending:
      self._finish(pending.pop(0), stats)
    return multi_frame

  def _trajectories(self, indices, loaded=None):
    """(index, trajectory) in order: through the loader's read-ahead threads where it has them (file-backed trajectory
    types that split their constructor), else built one by one on this thread as `self.TRAJLOADER[tid]`."""
    loader, workers = self.TRAJLOADER, self.prefetch_workers
    if workers is None:
      import os
      ranks_here = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
      workers = max(2, min(8, (os.cpu_count() or 4) // (2 * ranks_here)))
    workers = int(workers)
    if loaded is None and workers > 0 and hasattr(loader, "prefetched"):
      yield from loader.prefetched(indices, workers=workers)
      return
    for tid in indices:
      traj = loaded.get(tid) if loaded is not None else None
      yield tid, (traj if traj is not None else loader[tid])

  # ---- single-frame systems (PDBBind-style): many systems per native call ------------------------------
  def _static_groupable(self):
    """Static featurisation of single-frame systems, every feature on the device path: `static_group` systems
    are concatenated into one pseudo-system per native call, each a slice of its atoms with its own focus box."""
    if not (self.use_native_batch and self.time_window == 1 and int(self.static_group) > 1 and self.FEATURENUMBER > 0):
      return False
    n_static = 0
    for feat in self.FEATURESPACE:
      if getattr(feat, "_native_kind", None) == "label":
        continue                      # per-system scalars (affinity labels ...): evaluated on the host, member by member
      if getattr(feat, "_native_kind", None) != "static" or not feat._is_builtin("query", "run", "cache"):
        return False
      if feat.selection is not None and not isinstance(feat.selection, str):
        return False
      if int(feat.frame_offset or 0) != 0:
        return False
      n_static += 1
    return n_static > 0

  def _run_static_groups(self, indices, depth, stats):
    """Groups of `static_group` single-frame systems per native call, `depth` calls in flight.  Returns how many of
    `indices` were featurised: a multi-frame trajectory switches grouping off, and run() takes the rest of the list
    through the pipelined one-call-per-trajectory loop (trajectory lists are homogeneous in practice)."""
    group = max(2, int(self.static_group))
    pending = []
    submitted = 0
    source = self._trajectories(indices)                     # file-backed systems are read ahead on the loader threads
    for g0 in range(0, len(indices), group):
      tids = indices[g0:g0 + group]
      if len(pending) >= depth:
        self._finish_static_group(pending.pop(0), stats)
      loaded = {}
      for _ in tids:
        tid, traj = next(source)
        loaded[tid] = traj
      lane = submitted % depth
      assert all(job["lane"] != lane for job in pending), "pipeline lane handed out while its job is pending"
      job = self._submit_static_group(tids, lane=lane, loaded=loaded)
      if job is None:
        # a multi-frame trajectory (or a feature without weights) in this group: one call per trajectory, reference
        # order kept, re-using the trajectories the attempt already loaded
        while pending:
          self._finish_static_group(pending.pop(0), stats)
        if self._run_pipelined(tids, depth, stats, loaded=loaded):
          source.close()                                     # the rest of the list goes through the pipelined loop's own source
          return g0 + len(tids)
        continue
      if job["batch"] is None:
        continue                        # every member was skipped: nothing was submitted, the lane stays free
      submitted += 1
      pending.append(job)
    while pending:
      self._finish_static_group(pending.pop(0), stats)
    return len(indices)

  def _submit_static_group(self, tids, lane, loaded=None):
    members = []
    n_focal = None
    for tid in tids:
      traj = loaded.get(tid) if loaded is not None else None
      if traj is None:
        traj = self.TRAJLOADER[tid]
        if loaded is not None:
          loaded[tid] = traj
      self.traj = traj
      if traj.n_frames != 1:
        return None
      if self.FOCALPOINTS_PROTOTYPE is not None:
        if not self.parse_focus():
          printit(f"{self.classname} Warning: Skipping the trajectory {traj.identity}(index {tid + 1}) because focal points parsing is failed. ")
          continue
        focal = np.array(self.FOCALPOINTS[0], dtype=np.float32)
      else:
        focal = np.zeros((1, 3), dtype=np.float32)
      for feat in self.FEATURESPACE:
        feat.cache(traj)
      specs = [feat._native_spec() for feat in self.FEATURESPACE]
      label_feats = [f for f, feat in enumerate(self.FEATURESPACE) if feat._native_kind == "label"]
      if any(sp is None for f, sp in enumerate(specs) if f not in label_feats):
        return None
      if n_focal is None:
        n_focal = len(focal)
      if len(focal) != n_focal:
        return None
      xyz = traj.xyz
      if not isinstance(xyz, np.ndarray) or xyz.dtype not in (np.float32, np.float64):
        xyz = np.asarray(xyz, dtype=np.float32)
      labels = {(pid, f): self.FEATURESPACE[f].run(*self.FEATURESPACE[f].query(traj.top, traj.xyz[0:1], focal[pid]))
                for pid in range(len(focal)) for f in label_feats}
      members.append({"tid": tid, "identity": traj.identity, "xyz": xyz, "focal": focal, "specs": specs, "labels": labels,
                      "resids": next(feat.resids for feat in self.FEATURESPACE if feat._native_kind != "label")})
    if not members:
      return {"members": [], "batch": None, "lane": lane}
    dtype = members[0]["xyz"].dtype
    if any(m["xyz"].dtype != dtype for m in members):
      return None
    sizes = np.array([m["xyz"].shape[1] for m in members], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    n_total = int(offsets[-1])
    cat = self._table_buffer(lane, "group_xyz_" + dtype.name, (1, n_total, 3), dtype)
    for m, o in zip(members, offsets):
      cat[0, o:o + m["xyz"].shape[1]] = m["xyz"][0]
    native_feats = [f for f, feat in enumerate(self.FEATURESPACE) if feat._native_kind != "label"]
    specs_cat = [None] * self.FEATURENUMBER
    for f in native_feats:
      sp = dict(members[0]["specs"][f])
      sp["weights"] = np.concatenate([np.asarray(m["specs"][f]["weights"], dtype=np.float32) for m in members])
      specs_cat[f] = sp
    residue_index = None
    if any(self.FEATURESPACE[f].byres for f in native_feats):
      dense, most = [], 1
      for m in members:
        _, inv = np.unique(m["resids"], return_inverse=True)
        dense.append(inv)
        most = max(most, int(inv.max()) + 1 if len(inv) else 1)
      residue_index = (np.concatenate(dense), most)
    starts = np.repeat(offsets[:-1], n_focal)
    stops = np.repeat(offsets[1:], n_focal)
    focal = np.concatenate([m["focal"] for m in members], axis=0)
    batch = self._lower_residue_batch(native_feats, specs_cat, dtype == np.float64, 0, starts, stops, np.zeros(len(starts), dtype=np.int64),
                                      focal, n_total, lane, residue_index=residue_index)
    buffers = self.__dict__.setdefault("_out_buffers", {})
    if lane not in buffers or buffers[lane].size < batch.out_floats:
      buffers[lane] = pinned_empty((batch.out_floats,), np.float32)
    out = buffers[lane][:batch.out_floats]
    batch.execute(_lib.ptr(cat), 1, _lib.ptr(out), no_sync=True, lane=lane)
    return {"members": members, "batch": batch, "out": out, "lane": lane, "n_focal": n_focal, "xyz": cat}

  def _finish_static_group(self, job, stats):
    if job["batch"] is None:
      return
    _lib.check(_lib.lib().nrl_lane_synchronize(job["lane"]))
    grids = job["batch"].split(job["out"], copy=self.keep_results)
    nf, n_focal = self.FEATURENUMBER, job["n_focal"]
    k = 0
    with utils.batched_output():
      for m in job["members"]:
        for pid in range(n_focal):
          for fidx in range(nf):
            feat = self.FEATURESPACE[fidx]
            if (pid, fidx) in m["labels"]:
              result = m["labels"][(pid, fidx)]
            else:
              result = grids[k]
              k += 1
            if self.result_sink is not None:
              self.result_sink(feat.outkey, m["tid"], 0, pid, result)
            if self.keep_results:
              self.RESULTS.setdefault(feat.outkey, []).append(result)
            if feat.outfile or (self.result_sink is None and not self.keep_results):
              feat.dump(result)
        stats["trajectories"] += 1
        stats["tasks"] += n_focal * nf
        stats["native_tasks"] += n_focal * nf - len(m["labels"])
        stats["frames"] += 1

  def main_loop(self, *args, **kwargs):
    """Name used by the README / older scripts for the same loop."""
    return self.run(*args, **kwargs)

  def _submit(self, tid, lane=0, traj=None):
    """Everything up to and including the asynchronous native call for trajectory `tid`."""
    self.traj = traj if traj is not None else self.TRAJLOADER[tid]
    printit(f"{self.classname}: Processing the trajectory {tid + 1} ({self.traj.identity}) with {self.SLICENUMBER} frame slices")
    st = time.perf_counter()
    if self.FOCALPOINTS_PROTOTYPE is not None:
      if not self.parse_focus():
        printit(f"{self.classname} Warning: Skipping the trajectory {self.traj.identity}(index {tid + 1}) because focal points parsing is failed. ")
        return None
    else:
      self.FOCALNUMBER = 0
    for feat in self.FEATURESPACE:
      feat.cache(self.traj)
    # task list in the reference's order: window -> focal point -> feature
    n_focal = self.FOCALNUMBER if self.FOCALNUMBER > 0 else 1
    specs = self._native_specs()
    native_feats = [f for f in range(self.FEATURENUMBER) if specs[f] is not None]
    slabs = None
    if self._task_shard is None:
      task_keys = [(bid, pid, fidx) for bid in range(self.SLICENUMBER) for pid in range(n_focal)
                   for fidx in range(self.FEATURENUMBER)]
    else:
      task_keys, slabs = self.shard_task_list(self._task_shard, specs)
    native_ids = [i for i, (_, _, fidx) in enumerate(task_keys) if specs[fidx] is not None]
    job = {"tid": tid, "traj": self.traj, "task_keys": task_keys, "native_ids": native_ids, "lane": lane, "start": st, "slabs": slabs,
           "generic_ids": [i for i, (_, _, fidx) in enumerate(task_keys) if specs[fidx] is None], "batch": None}
    job["generic_results"] = self._run_generic_tasks(job)
    if slabs is not None and self._shard_device_out is not None and self._task_shard[0] == 0:
      # rank 0 lays the whole grids out: what every rank sends, in the order it sends it
      job["all_parts"] = [self.shard_task_list((r,) + self._task_shard[1:], specs) for r in range(self._task_shard[1])]
      job["native_mask"] = [sp is not None for sp in specs]
    job["has_generic"] = any(sp is None for sp in specs)
    xyz_dev = None
    if slabs is not None and self._shard_device_xyz is not None and native_feats:
      # Grid slabs: several ranks read the same frames, so the coordinates go up split over the ranks and an all-gather
      # completes them on every device.  A collective: EVERY rank takes part for this trajectory, whether or not it has a
      # task of its own (the decision depends on the trajectory and the world only, never on the rank).
      units, world = self.SLICENUMBER * n_focal, self._task_shard[1]
      n_slabs = self._task_shard[2] or (1 if units >= world or units == 0 else -(-world // units))
      if n_slabs > 1:
        xyz = self.traj.xyz
        if not isinstance(xyz, np.ndarray) or xyz.dtype not in (np.float32, np.float64) or not xyz.flags.c_contiguous:
          xyz = np.ascontiguousarray(xyz, dtype=np.float64 if np.asarray(xyz).dtype == np.float64 else np.float32)
        xyz_dev = self._shard_device_xyz(lane, xyz)
    if native_ids:
      xyz = self.traj.xyz
      if not isinstance(xyz, np.ndarray) or xyz.dtype not in (np.float32, np.float64) or not xyz.flags.c_contiguous:
        xyz = np.ascontiguousarray(xyz, dtype=np.float64 if np.asarray(xyz).dtype == np.float64 else np.float32)
      task_list = None if slabs is None else tuple(task_keys[i] + slabs[i] for i in native_ids)
      batch = self.lower_batch(native_feats, specs, xyz.shape[1], xyz.dtype == np.float64, lane=lane, task_list=task_list)
      if slabs is not None and self._shard_device_out is not None:
        # task-sharded run whose parts go to rank 0 over NVLink: results stay on the device, in the caller's buffer of this lane
        if xyz_dev is not None:
          batch.execute(xyz_dev, xyz.shape[0], self._shard_device_out(lane, batch.out_floats), xyz_on_device=True,
                        out_on_device=True, no_sync=True, lane=lane)
        else:
          batch.execute(_lib.addr(xyz), xyz.shape[0], self._shard_device_out(lane, batch.out_floats), out_on_device=True, no_sync=True, lane=lane)
        job.update(batch=batch, out=None, xyz=xyz)
        return job
      buffers = self.__dict__.setdefault("_out_buffers", {})
      if lane not in buffers or buffers[lane].size < batch.out_floats:
        buffers[lane] = pinned_empty((max(batch.out_floats, 1),), np.float32)     # page-locked, reused across trajectories
      out = buffers[lane][:batch.out_floats]
      batch.execute(_lib.addr(xyz), xyz.shape[0], _lib.addr(out), no_sync=True, lane=lane)      # both live in `job`
      job.update(batch=batch, out=out, xyz=xyz)
    return job

  def _native_specs(self):
    """`_native_spec()` of every feature for the active trajectory (None: host protocol).  A spec is a function of the
    feature's parameters -- fixed inside a run -- and of the per-atom arrays `cache` derived from the topology, which are
    memoised per topology object: the spec of a (feature, arrays) pair is built once per run (run() clears the memo)."""
    memo = self.__dict__.setdefault("_spec_memo", {})
    specs = []
    for fidx, feat in enumerate(self.FEATURESPACE):
      if not (self.use_native_batch and hasattr(feat, "_native_spec")):
        specs.append(None)
        continue
      arr, sel = getattr(feat, "cached_array", None), getattr(feat, "selected", None)
      key = (fidx, id(arr), id(sel), _lib.OPTION_EPOCH)
      hit = memo.get(key)
      if hit is None or hit[0] is not arr or hit[1] is not sel or arr is None:
        if len(memo) > 1024 or hit is not None:
          # spec objects are about to be released: the identities `_template` keys its shortcut on may come back
          memo.clear()
          self.__dict__.pop("_template_quick", None)
        hit = memo[key] = (arr, sel, feat._native_spec())
      specs.append(hit[2])
    return specs

  def shard_task_list(self, task_shard, specs):
    """The (window, focal point, feature) tasks of the ACTIVE trajectory that rank `task_shard[0]` of `task_shard[1]`
    computes, with the grid slab (x_begin, x_count; 0, 0 = whole grid) of each: unit -> slab -> feature order.  A task that
    cannot be cut (host protocol, non-cubic grid) goes whole to the owner of the unit's first slab."""
    from . import sharding
    rank, world, n_slabs = (tuple(task_shard) + (None,))[:3]
    n_focal = self.FOCALNUMBER if self.FOCALNUMBER > 0 else 1
    task_keys, slabs = [], []
    for unit, cuts in sorted(sharding.task_parts(self.SLICENUMBER * n_focal, rank, world, n_slabs).items()):
      bid, pid = divmod(unit, n_focal)
      for slab, of in cuts:
        for fidx, feat in enumerate(self.FEATURESPACE):
          dims = [int(v) for v in feat.dims] if getattr(feat, "dims", None) is not None else []
          cuttable = specs[fidx] is not None and of > 1 and len(dims) == 3 and dims[0] == dims[1] == dims[2]
          xb, xc = sharding.slab_range(dims[0], slab, of) if cuttable else ((0, 0) if slab == 0 else (0, -1))
          if xc < 0 or (cuttable and xc == 0):
            continue
          task_keys.append((bid, pid, fidx))
          slabs.append((xb, xc))
    return task_keys, slabs

  def _run_generic_tasks(self, job):
    """The host-protocol tasks of `job` (label features, user subclasses that override query / run), evaluated at
    submit time -- while `self.traj`, the focal points and the features' caches are this trajectory's -- so that
    they do not stop the pipeline: {task index: result}."""
    out = {}
    frames_cache = {}
    task_keys = job["task_keys"]
    for i in job["generic_ids"]:
      bid, pid, fidx = task_keys[i]
      self.frame_slice = self.FRAMESLICES[bid]
      if bid not in frames_cache:
        frames_cache = {bid: self.traj.xyz[self.FRAMESLICES[bid]]}
      focal = self.FOCALPOINTS[bid, pid] if self.FOCALNUMBER > 0 else np.zeros(3, dtype=np.float32)
      feat = self.FEATURESPACE[fidx]
      out[i] = feat.run(*feat.query(self.top, frames_cache[bid], focal))
    return out

  def _finish(self, job, stats):
    """Wait for the native results of `job`, merge its host-protocol results, dump in order."""
    task_keys = job["task_keys"]
    results = [None] * len(task_keys)
    on_device = job["batch"] is not None and job.get("out", 0) is None
    if job["batch"] is not None:
      _lib.check(_lib.lib().nrl_lane_synchronize(job["lane"]))
      if not on_device:
        for i, arr in zip(job["native_ids"], job["batch"].split(job["out"], copy=self.keep_results)):
          results[i] = arr
    for i, res in job["generic_results"].items():
      results[i] = res
    tid = job["tid"]
    if job.get("slabs") is not None and (on_device or self._shard_device_out is not None):
      # a task-sharded run whose native parts stay in the lane's device buffer: describe them, keep the host-protocol ones
      self.SHARD_RESULTS[tid] = {
        "lane": job["lane"], "floats": job["batch"].out_floats if job["batch"] is not None else 0,
        "generic": [(bid, pid, fidx, 0, np.array(results[i])) for i, (bid, pid, fidx) in enumerate(task_keys) if i in job["generic_results"]],
        "all_parts": job.get("all_parts"), "native_mask": job.get("native_mask"), "has_generic": job["has_generic"],
        "dims": [tuple(int(v) for v in f.dims) if getattr(f, "dims", None) is not None else None for f in self.FEATURESPACE]}
      stats["trajectories"] += 1
      stats["tasks"] += len(task_keys)
      stats["native_tasks"] += len(job["native_ids"])
      stats["frames"] += job["traj"].n_frames // self.time_window * self.time_window
      return
    if job.get("slabs") is not None:
      # a task-sharded run: the parts wait for sharding.run_task_sharded (copies: the lane's buffer is reused)
      self.SHARD_RESULTS.setdefault(tid, []).extend(
        (bid, pid, fidx, xb, np.array(res)) for (bid, pid, fidx), (xb, _), res in zip(task_keys, job["slabs"], results))
      stats["trajectories"] += 1
      stats["tasks"] += len(task_keys)
      stats["native_tasks"] += len(job["native_ids"])
      stats["frames"] += job["traj"].n_frames // self.time_window * self.time_window
      return
    printit(f"{self.classname}: Trajectory {tid + 1} yields {len(task_keys)} frame-slices (tasks) for the featurization. ")
    with utils.batched_output():
      for (bid, pid, fidx), result in zip(task_keys, results):
        feat = self.FEATURESPACE[fidx]
        if self.result_sink is not None:
          self.result_sink(feat.outkey, tid, bid, pid, result)
        if self.keep_results:
          self.RESULTS.setdefault(feat.outkey, []).append(result)
        if feat.outfile or (self.result_sink is None and not self.keep_results):
          feat.dump(result)
    stats["trajectories"] += 1
    stats["tasks"] += len(task_keys)
    stats["native_tasks"] += len(job["native_ids"])
    stats["frames"] += job["traj"].n_frames // self.time_window * self.time_window
    printit(f"{self.classname}: Finished the trajectory {tid + 1} / {self.TRAJECTORYNUMBER} with {len(task_keys)} tasks in {time.perf_counter() - job['start']:.6f} seconds")

  def deliver_shard_results(self, parts):
    """`parts`: {trajectory index: [(window, focal point, feature index, x_begin, array), ...]} from every rank of a
    task-sharded run.  Slabs are put together along the first axis and the results are handed on in the reference's
    order (trajectory -> window -> focal point -> feature), as `_finish` does for an unsharded run."""
    with utils.hdf_session():
      for tid in sorted(parts):
        whole = {}
        for bid, pid, fidx, xb, arr in parts[tid]:
          whole.setdefault((bid, pid, fidx), []).append((xb, arr))
        with utils.batched_output():
          for (bid, pid, fidx) in sorted(whole):
            pieces = sorted(whole[(bid, pid, fidx)], key=lambda p: p[0])
            result = pieces[0][1] if len(pieces) == 1 else np.concatenate([a for _, a in pieces], axis=0)
            feat = self.FEATURESPACE[fidx]
            if self.result_sink is not None:
              self.result_sink(feat.outkey, tid, bid, pid, result)
            if self.keep_results:
              self.RESULTS.setdefault(feat.outkey, []).append(result)
            if feat.outfile or (self.result_sink is None and not self.keep_results):
              feat.dump(result)

  # ---- residue-classification data sets ------------------------------------------------------------
  def _residue_blocks(self, restype, tag_limit=0):
    """[(atom slice, label)] of the active trajectory in the reference's order
    (featurizer.py:487-532): residue names in constants.RES + RES_PATCH order, blocks in topology order."""
    names = list(constants.RES) + list(constants.RES_PATCH.keys())
    patched = lambda name: constants.RES_PATCH.get(name, name)
    blocks = []
    if restype == "single":
      for name in names:
        label = constants.RES2LAB[patched(name)]
        for sidx, s_ in enumerate(utils.find_block_single(self.traj, name)):
          if tag_limit > 0 and sidx >= tag_limit:
            break
          blocks.append((s_, label))
    elif restype == "dual":
      pairs = utils.dual_blocks_by_key(self.traj)         # one pass; same lists as find_block_dual(traj, key)
      for res1 in names:
        for res2 in names:
          label = constants.RES2LAB_DUAL.get(patched(res1) + patched(res2), "Unknown")
          for s_ in pairs.get(res1 + res2, ()):
            blocks.append((s_, label))
    else:
      raise ValueError(f"Unexpected residue type {restype!r}; choose 'single' or 'dual'")
    return blocks

  def loop_by_residue(self, restype, tag_limit=0, process_nr=None, chunk_floats=1 << 26, **_ignored):
    """Featurise every single residue (``restype="single"``) or every bonded residue pair (``"dual"``) of
    every trajectory as its own sample: the atoms of the block only, the box centred on the block's centre
    of geometry in the first frame of each window (/root/reference/nearl/featurizer.py:471-554).  Rows are
    appended per feature in the order window -> residue name -> block; the class labels go to the
    dataset ``label`` of the featuriser's ``outfile``.

    All (window, block, feature) tasks of a window go to the device in a few batched calls (the crop
    kernel reads the block's atoms only).  Two behaviours of the reference are reproduced on purpose:
    a feature is re-cached only when the block length changes (features.py:452-457), so a block is read
    with the per-atom tables of the last block of a different length that preceded one of its own length;
    and ``label`` receives the first n_blocks entries of the per-TASK label list (featurizer.py:545-549).
    `process_nr` (some scripts pass a worker count) is accepted and ignored."""
    with utils.hdf_session():                                # output files stay open until the loop is over
      return self._loop_by_residue(restype, tag_limit, chunk_floats)

  def _loop_by_residue(self, restype, tag_limit, chunk_floats):
    stats = {"trajectories": 0, "tasks": 0, "native_tasks": 0, "blocks": 0, "seconds": 0.0}
    t_run = time.perf_counter()
    for tid, traj in self._trajectories(range(self.TRAJECTORYNUMBER)):       # file-backed trajectories are read ahead
      self.traj = traj
      printit(f"{self.classname}: Start processing the trajectory {tid + 1} with {self.SLICENUMBER} frames")
      for feat in self.FEATURESPACE:
        feat.cache(self.traj)
      blocks = self._residue_blocks(restype, tag_limit)
      specs = [feat._native_spec() if (self.use_native_batch and hasattr(feat, "_native_spec") and feat._slice_invariant())
               else None for feat in self.FEATURESPACE]
      for f, sp in enumerate(specs):
        if sp is not None and sp["mode"] == 3 and sp.get("ranks") is None:
          specs[f] = None          # value-ordered crops cannot be sliced: host protocol
      native_feats = [f for f in range(self.FEATURENUMBER) if specs[f] is not None]
      generic_feats = [f for f in range(self.FEATURENUMBER) if specs[f] is None]
      xyz = self.traj.xyz
      if not isinstance(xyz, np.ndarray) or xyz.dtype not in (np.float32, np.float64) or not xyz.flags.c_contiguous:
        xyz = np.ascontiguousarray(xyz, dtype=np.float64 if np.asarray(xyz).dtype == np.float64 else np.float32)
      n_all = xyz.shape[1]
      labels = []
      # the reference's cache state: length of the slice the features were last cached on, and where it starts
      cached_len, src_first = n_all, 0
      pending = None
      submitted = 0                     # native chunks sent so far: with one job pending, alternate lanes never alias
      with utils.batched_output():
        for bid in range(self.SLICENUMBER):
          first = bid * self.time_window
          frames = xyz[self.FRAMESLICES[bid]]
          starts = np.array([s_.start for s_, _ in blocks], dtype=np.int64)
          stops = np.array([s_.stop for s_, _ in blocks], dtype=np.int64)
          shifts = np.zeros(len(blocks), dtype=np.int64)
          focal = np.zeros((len(blocks), 3), dtype=xyz.dtype)
          for k, (s_, label) in enumerate(blocks):
            if stops[k] - starts[k] != cached_len:
              cached_len, src_first = int(stops[k] - starts[k]), int(starts[k])
            shifts[k] = src_first - starts[k]
            focal[k] = np.mean(frames[0, s_], axis=0)                   # featurizer.py:503
            labels.extend([label] * self.FEATURENUMBER)
          # host-protocol features first (they follow the reference call by call, cache state included)
          generic = {}
          for k, (s_, label) in enumerate(blocks):
            for f in generic_feats:
              feat = self.FEATURESPACE[f]
              sub_top = self.traj.top[s_]
              generic[(k, f)] = feat.run(*feat.query(sub_top, frames[:, s_, :].copy(), focal[k]))
          # native features: chunks of blocks, pipelined over the two lanes
          per_block = sum(int(np.prod(self.FEATURESPACE[f].dims)) for f in native_feats)
          step = max(1, int(chunk_floats // max(per_block, 1))) if native_feats else max(len(blocks), 1)
          for c0 in range(0, len(blocks), step):
            c1 = min(c0 + step, len(blocks))
            job = {"bid": bid, "c0": c0, "c1": c1, "generic": generic, "native": None}
            if native_feats:
              lane = submitted % 2      # the pending job (if any) holds the other lane
              submitted += 1
              batch = self._lower_residue_batch(native_feats, specs, xyz.dtype == np.float64, first, starts[c0:c1], stops[c0:c1],
                                                shifts[c0:c1], focal[c0:c1], n_all, lane)
              buffers = self.__dict__.setdefault("_out_buffers", {})
              if lane not in buffers or buffers[lane].size < batch.out_floats:
                buffers[lane] = pinned_empty((batch.out_floats,), np.float32)
              out = buffers[lane][:batch.out_floats]
              batch.execute(_lib.ptr(xyz), xyz.shape[0], _lib.ptr(out), no_sync=True, lane=lane)
              job["native"] = (batch, out, lane)
            if pending is not None:
              self._finish_residue_chunk(pending, native_feats, generic_feats, stats)
            pending = job
        if pending is not None:
          self._finish_residue_chunk(pending, native_feats, generic_feats, stats)
          pending = None
      stats["blocks"] += len(blocks) * self.SLICENUMBER
      printit(f"{self.classname}: Tasks are finished, dumping the results to the feature space...")
      outfile = self.FEATURE_PARMS.get("outfile", None)
      if outfile is not None and self.FEATURENUMBER:
        lab = np.array(labels, dtype=int)[:len(labels) // self.FEATURENUMBER]
        opts = {"chunks": True}
        if (self.FEATURE_PARMS.get("hdf_compress_level", 0) or 0) > 0:
          opts.update(compression="gzip", compression_opts=int(self.FEATURE_PARMS["hdf_compress_level"]))
        utils.append_hdf_data(outfile, "label", lab, dtype=int, maxshape=(None,), **opts)
      if self.keep_results:
        self.RESULTS.setdefault("label", []).extend(labels[:len(labels) // max(self.FEATURENUMBER, 1)])
      stats["trajectories"] += 1
    stats["seconds"] = time.perf_counter() - t_run
    self.last_run_stats = stats
    printit(f"{self.classname}: All trajectories and tasks are finished")
    return stats

  def _finish_residue_chunk(self, job, native_feats, generic_feats, stats):
    """Dump the results of one chunk of blocks in the reference's order (block -> feature)."""
    native = {}
    if job["native"] is not None:
      batch, out, lane = job["native"]
      _lib.check(_lib.lib().nrl_lane_synchronize(lane))
      grids = batch.split(out, copy=self.keep_results)
      for i in range(job["c1"] - job["c0"]):
        for j, f in enumerate(native_feats):
          native[(job["c0"] + i, f)] = grids[i * len(native_feats) + j]
    for k in range(job["c0"], job["c1"]):
      for f in range(self.FEATURENUMBER):
        result = native[(k, f)] if (k, f) in native else job["generic"][(k, f)]
        feat = self.FEATURESPACE[f]
        if self.result_sink is not None:
          self.result_sink(feat.outkey, self.traj.identity, job["bid"], k, result)
        if self.keep_results:
          self.RESULTS.setdefault(feat.outkey, []).append(result)
        if feat.outfile or (self.result_sink is None and not self.keep_results):
          feat.dump(result)
        stats["tasks"] += 1
        stats["native_tasks"] += 1 if (k, f) in native else 0

  def _lower_residue_batch(self, native_feats, specs, is_f64, first_frame, starts, stops, shifts, focal, n_atoms, lane,
                           selection_rows=None, residue_index=None):
    """Crop / task tables for the blocks [starts, stops) of ONE window: per block one crop per crop class
    (features that share frames, padding, by-residue, selection and cap) and one task per native feature.
    `selection_rows` ({feature index: mask over all atoms}) and `residue_index` ((dense index per atom, count))
    override the features' own cached arrays (used when several systems are concatenated into one call)."""
    feats = [self.FEATURESPACE[f] for f in native_feats]
    sel_feats = [f for f in native_feats if self.FEATURESPACE[f].selection is not None and not specs[f].get("selection_folded")]
    sel_of = {f: i for i, f in enumerate(sel_feats)}
    rank_feats = [f for f in native_feats if specs[f]["mode"] == 3]
    rank_of = {f: len(native_feats) + i for i, f in enumerate(rank_feats)}
    classes, class_of = [], {}
    for f in native_feats:
      feat, spec = self.FEATURESPACE[f], specs[f]
      half = feat.spacing / 2
      lo = 0 - feat.padding - half
      lo_cmp = float(np.float32(lo)) if not is_f64 else lo
      hi = tuple(float(feat.lengths[k] + feat.padding - half) for k in range(3))
      static = spec["kind"] == "static"
      key = (first_frame + (spec["frame_offset"] if static else 0), 1 if static else self.time_window, feat.center.tobytes(), half,
             lo_cmp, hi, bool(feat.byres), sel_of.get(f, -1), spec["max_atoms"], bool(feat.force_recache))
      if key not in classes:
        classes.append(key)
      class_of[f] = classes.index(key)
    nb, nc, nf = len(starts), len(classes), len(native_feats)
    crops = np.zeros((nb, nc), dtype=np.dtype(_lib.CropStruct, align=True))
    for ci, (start, n_frames, center, half, lo_cmp, hi, byres, sid, max_atoms, recache) in enumerate(classes):
      col = crops[:, ci]
      col["frame_start"], col["n_frames"] = start, n_frames
      col["focal"] = focal.astype(np.float32)
      col["focal_f64"] = focal.astype(np.float64)
      col["center"] = np.frombuffer(center, dtype=np.float32)
      col["half_spacing_f32"], col["half_spacing_f64"] = np.float32(half), half
      col["lo"], col["hi"] = lo_cmp, hi
      col["byres"], col["selection"], col["max_atoms"], col["sort_table"] = int(byres), sid, max_atoms, -1
      col["atom_begin"], col["atom_end"] = starts, stops
      col["table_shift"] = 0 if recache else shifts
    tasks = np.zeros((nb, nf), dtype=np.dtype(_lib.TaskStruct, align=True))
    shapes = []
    for j, f in enumerate(native_feats):
      feat, spec = self.FEATURESPACE[f], specs[f]
      col = tasks[:, j]
      col["crop"] = np.arange(nb) * nc + class_of[f]
      col["mode"], col["agg"], col["weight_table"] = spec["mode"], spec["agg"], j
      col["dims"] = [int(v) for v in feat.dims]
      col["spacing"], col["cutoff"], col["sigma"] = float(feat.spacing), float(feat.cutoff), float(spec["sigma"])
      col["rank_table"] = rank_of.get(f, -1)
      shapes.append(tuple(int(v) for v in feat.dims))
    batch = LoweredBatch()
    batch.n_atoms, batch.is_f64 = n_atoms, bool(is_f64)
    batch._keep = (crops, tasks)       # the ctypes views below borrow these arrays
    batch.crops = crops.ctypes.data_as(ctypes.POINTER(_lib.CropStruct))
    batch.n_crops = nb * nc
    batch.tasks = tasks.ctypes.data_as(ctypes.POINTER(_lib.TaskStruct))
    batch.n_tasks = nb * nf
    batch.shapes = shapes * nb
    batch.out_floats = int(sum(sh[0] * sh[1] * sh[2] for sh in shapes)) * nb
    rows = [specs[f]["weights"] for f in native_feats] + [specs[f]["ranks"] for f in rank_feats]
    batch.weight_tables = self._table_buffer(lane, "weights", (len(rows), n_atoms), np.float32)
    for i, row in enumerate(rows):
      batch.weight_tables[i] = row
    batch.selections = None
    if sel_feats:
      batch.selections = self._table_buffer(lane, "selections", (len(sel_feats), n_atoms), np.uint8)
      for i, f in enumerate(sel_feats):
        batch.selections[i] = selection_rows[f] if selection_rows is not None else self.FEATURESPACE[f].selected
    batch.value_orders = None
    batch.residue_index, batch.n_residues = None, 0
    if any(c[6] for c in classes):
      if residue_index is not None:
        batch.residue_index, batch.n_residues = np.ascontiguousarray(residue_index[0], dtype=np.int32), int(residue_index[1])
      else:
        uniq, inverse = np.unique(feats[0].resids, return_inverse=True)
        batch.residue_index, batch.n_residues = np.ascontiguousarray(inverse, dtype=np.int32), len(uniq)
    return batch

  # ---- lowering to nrl_featurize_batch --------------------------------------------------------
  def _template(self, native_feats, specs, n_atoms, is_f64, task_list=None):
    """Crop / task tables for the current (features, windows, focal points) structure.  Only the
    focal points, weights and selections change from one trajectory to the next, so the tables
    are built once and re-used.  Crops are shared by every feature that looks at the same frames
    through the same box (same padding / by-residue / selection / cap / emission order)."""
    n_focal = self.FOCALNUMBER if self.FOCALNUMBER > 0 else 1

    memo = self.__dict__.setdefault("_fingerprint_memo", {})      # cleared by run(): feature parameters are fixed inside a run

    def fingerprint(fidx):
      f, sp = self.FEATURESPACE[fidx], specs[fidx]
      fixed = memo.get(fidx)
      if fixed is None:
        fixed = memo[fidx] = (fidx, tuple(int(v) for v in f.dims), float(f.spacing), float(f.cutoff), float(f.padding), bool(f.byres),
                              f.selection is not None)
      return fixed + (float(sp["sigma"]), bool(sp.get("selection_folded")), sp["kind"], sp["mode"], sp["agg"],
                      sp["frame_offset"], sp["max_atoms"], sp.get("ranks") is not None, sp.get("weight_sig"))

    # inside a run the spec OBJECTS are memoised (`_native_specs`): their identities name the template without rebuilding
    # the fingerprints (run() clears this shortcut together with the spec memo, which keeps the objects alive)
    quick = self.__dict__.setdefault("_template_quick", {})
    qkey = (tuple(id(specs[f]) for f in native_feats), tuple(native_feats), self.SLICENUMBER, n_focal, self.time_window, bool(is_f64), task_list)
    if "_spec_memo" in self.__dict__ and qkey in quick:
      return quick[qkey]
    key = (tuple(fingerprint(f) for f in native_feats), self.SLICENUMBER, n_focal, self.time_window, bool(is_f64), task_list)
    cache = self.__dict__.setdefault("_template_cache", {})
    if key in cache:
      if "_spec_memo" in self.__dict__:
        quick[qkey] = cache[key]
      return cache[key]
    # one weight-table row per distinct weight signature (features without one get their own row)
    table_feats, table_of, sig_row = [], {}, {}
    for f in native_feats:
      sig = specs[f].get("weight_sig")
      if sig is not None and sig in sig_row:
        table_of[f] = sig_row[sig]
        continue
      table_of[f] = len(table_feats)
      if sig is not None:
        sig_row[sig] = len(table_feats)
      table_feats.append(f)
    # distinct_count features whose values have <= 32 ranks: the ranks travel as extra weight-table rows
    rank_feats = [f for f in native_feats if specs[f]["mode"] == 3 and specs[f].get("ranks") is not None]
    rank_of = {f: len(table_feats) + i for i, f in enumerate(rank_feats)}
    sel_feats = [f for f in native_feats if self.FEATURESPACE[f].selection is not None and not specs[f].get("selection_folded")]
    sel_of = {f: i for i, f in enumerate(sel_feats)}
    crops, crop_ids, crop_slots, tasks, shapes = [], {}, [], [], []
    order_feats, any_byres = [], False
    if task_list is None:      # every (window, focal point, feature), whole grids
      task_list = [(bid, pid, fidx, 0, 0) for bid in range(self.SLICENUMBER) for pid in range(n_focal) for fidx in native_feats]
    for bid, pid, fidx, x_begin, x_count in task_list:
      feat, spec = self.FEATURESPACE[fidx], specs[fidx]
      wid = table_of[fidx]
      sid = sel_of.get(fidx, -1)
      sort_table = wid if (spec["mode"] == 3 and fidx not in rank_of) else -1
      if sort_table >= 0 and fidx not in order_feats:
        order_feats.append(fidx)
      start = bid * self.time_window
      if spec["kind"] == "static":
        start, n_frames = start + spec["frame_offset"], 1
      else:
        n_frames = self.time_window
      half = feat.spacing / 2
      lo = 0 - feat.padding - half                                    # python float: numpy compares it as float32 ...
      lo_cmp = float(np.float32(lo)) if not is_f64 else lo             # ... against float32 coordinates (weak scalar)
      hi = tuple(float(feat.lengths[k] + feat.padding - half) for k in range(3))   # np.float64: compared in double
      ckey = (bid, pid, start, n_frames, feat.center.tobytes(), half, lo_cmp, hi, bool(feat.byres), sid,
              spec["max_atoms"], sort_table)
      if ckey not in crop_ids:
        crop_ids[ckey] = len(crops)
        c = _lib.CropStruct()
        c.frame_start, c.n_frames = start, n_frames
        c.center[:] = [float(v) for v in feat.center]
        c.half_spacing_f32 = float(np.float32(half))
        c.half_spacing_f64 = float(half)
        c.lo[:] = [lo_cmp] * 3
        c.hi[:] = hi
        c.byres = 1 if feat.byres else 0
        c.selection = sid
        c.max_atoms = spec["max_atoms"]
        c.sort_table = sort_table
        any_byres |= bool(feat.byres)
        crops.append(c)
        crop_slots.append((bid, pid))
      t = _lib.TaskStruct()
      t.crop, t.mode, t.agg, t.weight_table = crop_ids[ckey], spec["mode"], spec["agg"], wid
      t.rank_table = rank_of.get(fidx, -1)
      t.x_begin, t.x_count = int(x_begin), int(x_count)
      t.dims[:] = [int(v) for v in feat.dims]
      t.spacing, t.cutoff, t.sigma = float(feat.spacing), float(feat.cutoff), float(spec["sigma"])
      tasks.append(t)
      shapes.append((int(x_count) if x_count else int(feat.dims[0]), int(feat.dims[1]), int(feat.dims[2])))
    tpl = {"crops": crops, "crop_slots": crop_slots, "tasks": (_lib.TaskStruct * len(tasks))(*tasks), "shapes": shapes,
           "sel_feats": sel_feats, "order_feats": order_feats, "rank_feats": rank_feats, "table_feats": table_feats,
           "table_of": table_of, "any_byres": any_byres,
           "out_floats": int(sum(sh[0] * sh[1] * sh[2] for sh in shapes))}
    cache[key] = tpl
    return tpl

  def _table_buffer(self, lane, name, shape, dtype):
    """Re-used page-locked staging array (per pipeline lane): the table uploads inside the native call are DMA."""
    if lane is None:
      return np.empty(shape, dtype=dtype)
    pool = self.__dict__.setdefault("_table_buffers", {})
    need = int(np.prod(shape))
    buf = pool.get((lane, name))
    if buf is None or buf.size < need or buf.dtype != np.dtype(dtype):
      buf = pool[(lane, name)] = pinned_empty((max(need, 1),), dtype)
    return buf[:need].reshape(shape)

  def lower_batch(self, native_feats, specs, n_atoms, is_f64, lane=None, task_list=None):
    """Crop and task tables of the C ABI for the active trajectory (tasks in the reference's
    order window -> focal point -> feature, restricted to `native_feats`).  With a `lane` the
    per-trajectory tables are written into that lane's page-locked buffers, which stay valid until
    the lane's next submission."""
    tpl = self._template(native_feats, specs, n_atoms, is_f64, task_list)
    batch = LoweredBatch()
    batch.n_atoms, batch.is_f64 = n_atoms, bool(is_f64)
    # the crop table of the template as a numpy record array (a ctypes array costs microseconds per element to copy and
    # to fill); only the focal points are per trajectory
    crops_np = tpl.get("crops_np")
    if crops_np is None:
      n = len(tpl["crops"])
      crops_np = tpl["crops_np"] = np.frombuffer((_lib.CropStruct * n)(*tpl["crops"]), dtype=np.dtype(_lib.CropStruct, align=True)).copy()
      tpl["crop_bid"] = np.array([bid for bid, _ in tpl["crop_slots"]], dtype=np.int64)
      tpl["crop_pid"] = np.array([pid for _, pid in tpl["crop_slots"]], dtype=np.int64)
    crops = crops_np.copy()
    if self.FOCALNUMBER > 0 and len(crops):
      focal = self.FOCALPOINTS[tpl["crop_bid"], tpl["crop_pid"]]
      crops["focal"] = focal
      crops["focal_f64"] = focal
    else:
      crops["focal"] = 0.0
      crops["focal_f64"] = 0.0
    batch._keep = crops                # the ctypes view below borrows this array
    batch.crops, batch.n_crops = crops.ctypes.data_as(ctypes.POINTER(_lib.CropStruct)), len(crops)
    batch.tasks, batch.shapes, batch.out_floats = tpl["tasks"], tpl["shapes"], tpl["out_floats"]
    batch.n_tasks_total = len(tpl["shapes"])
    rows = [specs[f]["weights"] for f in tpl["table_feats"]] + [specs[f]["ranks"] for f in tpl["rank_feats"]]
    batch.weight_tables = self._table_buffer(lane, "weights", (len(rows), n_atoms), np.float32)
    for i, row in enumerate(rows):
      batch.weight_tables[i] = row
    batch.selections = None
    if tpl["sel_feats"]:
      batch.selections = self._table_buffer(lane, "selections", (len(tpl["sel_feats"]), n_atoms), np.uint8)
      for i, f in enumerate(tpl["sel_feats"]):
        batch.selections[i] = self.FEATURESPACE[f].selected
    batch.value_orders = None
    if tpl["order_feats"]:
      batch.value_orders = self._table_buffer(lane, "orders", (len(rows), n_atoms), np.int32)
      for f in tpl["order_feats"]:
        batch.value_orders[tpl["table_of"][f]] = self.FEATURESPACE[f].value_order()
    batch.residue_index, batch.n_residues = None, 0
    if tpl["any_byres"]:
      uniq, inverse = np.unique(self.FEATURESPACE[native_feats[0]].resids, return_inverse=True)
      batch.residue_index, batch.n_residues = np.ascontiguousarray(inverse, dtype=np.int32), len(uniq)
    return batch


class LoweredBatch:
  """Crop / task tables of one trajectory, ready for nrl_featurize_batch."""

  def execute(self, xyz_ptr, n_frames_total, out_ptr, xyz_on_device=False, out_on_device=False, no_sync=False, lane=0):
    b = _lib.BatchStruct()
    b.xyz, b.xyz_is_f64, b.xyz_on_device = xyz_ptr, int(self.is_f64), int(xyz_on_device)
    b.n_frames_total, b.n_atoms = int(n_frames_total), int(self.n_atoms)
    # (the batch owns these arrays: plain addresses, no ctypes views)
    b.weight_tables, b.n_weight_tables = _lib.addr(self.weight_tables), len(self.weight_tables)
    b.selections = _lib.addr(self.selections) if self.selections is not None else None
    b.n_selections = len(self.selections) if self.selections is not None else 0
    b.residue_index = _lib.addr(self.residue_index) if self.residue_index is not None else None
    b.n_residues = self.n_residues
    b.value_orders = _lib.addr(self.value_orders) if self.value_orders is not None else None
    b.crops, b.n_crops = self.crops, getattr(self, "n_crops", None) or len(self.crops)
    b.tasks, b.n_tasks = self.tasks, getattr(self, "n_tasks", None) or len(self.tasks)
    b.out, b.out_on_device = out_ptr, int(out_on_device)
    b.atom_counts = None
    b.flags = 1 if no_sync else 0
    b.lane = int(lane)
    _lib.check(_lib.lib().nrl_featurize_batch(ctypes.byref(b)))

  def split(self, out, copy=True):
    """Per-task views of the flat result buffer (copied out of the re-used pinned buffer when the
    caller keeps them)."""
    flat = np.array(out[:self.out_floats]) if copy else out[:self.out_floats]
    if len(self.shapes) > 1 and self.shapes.count(self.shapes[0]) == len(self.shapes):
      return list(flat.reshape((len(self.shapes),) + tuple(self.shapes[0])))      # equal shapes: rows of one block
    results, pos = [], 0
    for shape in self.shapes:
      size = shape[0] * shape[1] * shape[2]
      results.append(flat[pos:pos + size].reshape(shape))
      pos += size
    return results
