"""numpy restatement of the HOST half of the reference's hot path, feeding the C oracle.

TEST INFRASTRUCTURE ONLY (see oracle/nearl_oracle.c).  It restates, independently of
``nearl_b200.features`` / ``nearl_b200.featurizer``, what the reference does between a
trajectory and the native calls:

  focal points      /root/reference/nearl/featurizer.py:345-397  (parse_focus)
  frame windows     /root/reference/nearl/featurizer.py:229-241
  loop order        /root/reference/nearl/featurizer.py:429-441  (window -> focal -> feature)
  translate + crop  /root/reference/nearl/features.py:92-123,426-483
  static query/run  /root/reference/nearl/features.py:485-508,599-605
  dynamic query     /root/reference/nearl/features.py:1329-1368 (9999-atom padded layout)
  NaN -> 0 at dump  /root/reference/nearl/features.py:521-523
  residue loop      /root/reference/nearl/featurizer.py:471-554 (loop_by_residue), with the feature
                    re-cache rule of /root/reference/nearl/features.py:452-457

Parity status: PINNED against the reference's own Python.  tests/golden/make_host_golden.py imports the
unmodified /root/reference/nearl package in the build container (inert stand-ins for h5py / pytraj / rdkit /
openbabel, which it only needs at import, and a recorder in place of the compiled `all_actions`), runs its
Featurizer.run and loop_by_residue on seeded systems and keeps what reaches the FFI boundary; the committed
record (tests/golden/host_vectors.npz) is replayed by tests/test_host_golden.py, which demands that this file
hand the kernels bit-identical coordinate / weight arrays, dimensions, scalars, task order and labels.
The kernels it calls are pinned as described in nearl_oracle.c.
"""
import numpy as np

from . import oracle as orc

MAX_ALLOWED_ATOMS = 9999        # features.py:1279
DEFAULT_COORD = 99999.0         # features.py:1280


def focal_points_from_selections(xyz, selections, time_window):
  """[S, P, 3] float32: mean position of each selection in the first frame of each window."""
  n_slices = xyz.shape[0] // time_window
  focal = np.full((n_slices, len(selections), 3), 99999, dtype=np.float32)
  for midx, sel in enumerate(selections):
    for fidx in range(n_slices):
      frame = xyz[fidx * time_window]
      focal[fidx, midx] = np.mean(frame[sel], axis=0)
  return focal


def crop_mask(points, upperbound, padding, spacing):
  x0 = points[:, 0] < upperbound[0] + padding - spacing / 2
  x1 = points[:, 0] > 0 - padding - spacing / 2
  y0 = points[:, 1] < upperbound[1] + padding - spacing / 2
  y1 = points[:, 1] > 0 - padding - spacing / 2
  z0 = points[:, 2] < upperbound[2] + padding - spacing / 2
  z1 = points[:, 2] > 0 - padding - spacing / 2
  return np.array(x0 * x1 * y0 * y1 * z0 * z1, dtype=bool)


def base_query(desc, resids, frame_coords, focal_point):
  dims = np.array([int(i) for i in desc["dims"]])
  spacing = float(desc["spacing"])
  center = np.array(dims * spacing, dtype=np.float32) / 2
  lengths = dims * spacing
  padding = float(desc["padding"] if desc.get("padding") is not None else desc["cutoff"])
  coords = frame_coords - focal_point + center - spacing / 2
  mask = crop_mask(coords, lengths, padding, spacing)
  if desc.get("byres"):
    res_inbox = np.unique(resids[mask])
    final = np.full(len(resids), False)
    for res in res_inbox:
      final[np.where(resids == res)] = True
  else:
    final = mask
  if desc.get("selection") is not None:
    final = final * desc["selection"]
  return final, np.ascontiguousarray(coords[final])


class OracleKernels:
  """Kernel back-end of run_task: the C restatement (default)."""
  def __init__(self, buff_dim=-1):
    self.buff_dim = buff_dim

  def _b(self, cutoff, spacing):
    return self.buff_dim(cutoff, spacing) if callable(self.buff_dim) else self.buff_dim

  def frame_voxelize(self, coords, weights, dims, spacing, cutoff, sigma):
    return orc.frame_voxelize(coords, weights, dims, spacing, cutoff, sigma, buff_dim=self._b(cutoff, spacing))

  def density_flow(self, coords, weights, dims, spacing, cutoff, sigma, agg):
    return orc.density_flow(coords, weights, dims, spacing, cutoff, sigma, agg, buff_dim=self._b(cutoff, spacing))

  def marching_observer(self, coords, weights, dims, spacing, cutoff, obs, agg):
    return orc.marching_observer(coords, weights, dims, spacing, cutoff, obs, agg)


class ModuleKernels:
  """Kernel back-end that calls an `all_actions`-compatible module exactly as
  /root/reference/nearl/commands.py:56-66,136-142,177-190 does (used with the reference's own
  compiled kernels from oracle/_ref on the GPU box)."""
  def __init__(self, module):
    self.m = module

  def frame_voxelize(self, coords, weights, dims, spacing, cutoff, sigma):
    return self.m.frame_voxelize(coords, weights, np.array(dims, dtype=int), float(spacing), float(cutoff), float(sigma), 0)

  def density_flow(self, coords, weights, dims, spacing, cutoff, sigma, agg):
    return self.m.density_flow(coords, weights, np.array(dims, dtype=int), float(spacing), float(cutoff), float(sigma), int(agg))

  def marching_observer(self, coords, weights, dims, spacing, cutoff, obs, agg):
    return self.m.marching_observer(coords, weights, np.asarray(dims, dtype=int), spacing, cutoff, obs, agg)


def run_task(desc, resids, frames, focal_point, buff_dim=-1, kernels=None):
  """One (feature, window, focal point) task -> [dims] float32, before the NaN scrub.
  `buff_dim`: -1 (IEEE evaluation), an int, or a callable (cutoff, spacing) -> int giving the
  value the device computes (see nearl_oracle.c)."""
  dims = np.array([int(i) for i in desc["dims"]])
  spacing, cutoff = float(desc["spacing"]), float(desc["cutoff"])
  k = kernels if kernels is not None else OracleKernels(buff_dim)
  weights_all = np.asarray(desc["weights"], dtype=np.float32)
  if desc["kind"] == "static":
    frame = frames[int(desc.get("frame_offset") or 0)]
    mask, coords = base_query(desc, resids, frame, focal_point)
    if len(coords) == 0:
      return np.zeros(dims, dtype=np.float32)
    out = k.frame_voxelize(coords.astype(np.float32), weights_all[mask], dims, spacing, cutoff, float(desc["sigma"]))
    return out.reshape(dims)
  F = len(frames)
  coords = np.full((F, MAX_ALLOWED_ATOMS, 3), DEFAULT_COORD, dtype=np.float32)
  weights = np.full((F, MAX_ALLOWED_ATOMS), 0.0, dtype=np.float32)
  max_n = 0
  for idx, frame in enumerate(frames):
    mask, inbox = base_query(desc, resids, frame, focal_point)
    n = min(np.count_nonzero(mask), MAX_ALLOWED_ATOMS)
    coords[idx, :n] = inbox[:n]
    weights[idx, :n] = weights_all[mask][:n]
    max_n = max(max_n, n)
  c = np.ascontiguousarray(coords[:, :max_n], dtype=np.float32)
  w = np.ascontiguousarray(weights[:, :max_n].flatten(), dtype=np.float32)
  if desc["kind"] == "density_flow":
    out = k.density_flow(c, w, dims, spacing, cutoff, float(desc["sigma"]), int(desc["agg"]))
  else:
    out = k.marching_observer(c, w, dims, spacing, cutoff, int(desc["obs"]), int(desc["agg"]))
  return out.reshape(dims)


def featurize(xyz, resids, descs, focal, time_window, buff_dim=-1, scrub_nan=True, kernels=None, max_slices=None):
  """All tasks of one trajectory in the reference's order; returns {outkey: [rows...]}."""
  n_slices = xyz.shape[0] // time_window
  if max_slices is not None:
    n_slices = min(n_slices, max_slices)
  out = {d["outkey"]: [] for d in descs}
  for bid in range(n_slices):
    frames = xyz[bid * time_window:(bid + 1) * time_window]
    for pid in range(focal.shape[1]):
      for d in descs:
        res = run_task(d, resids, frames, focal[bid, pid], buff_dim=buff_dim, kernels=kernels)
        if scrub_nan and np.isnan(res).any():
          res = np.nan_to_num(res).astype(res.dtype)
        out[d["outkey"]].append(res)
  return {k: np.array(v, dtype=np.float32) for k, v in out.items()}


def featurize_by_residue(xyz, resids, descs, blocks, time_window, buff_dim=-1, scrub_nan=True, kernels=None):
  """loop_by_residue for one trajectory: every block (atom slice, label) of every window is a sample made of
  the block's atoms only, centred on the block's centre of geometry in the window's first frame
  (featurizer.py:487-506).  `descs[i]["weights"]` / ["selection"] are whole-topology arrays of slice-invariant
  per-atom values; a feature is re-cached on the block -- i.e. takes the block's part of them -- only when the
  block length differs from the length it was last cached on (features.py:452-457), otherwise it keeps the
  arrays of that earlier block.  Returns ({outkey: rows}, labels as written to the `label` dataset)."""
  n_slices = xyz.shape[0] // time_window
  out = {d["outkey"]: [] for d in descs}
  state = [{"len": xyz.shape[1], "w": np.asarray(d["weights"]), "sel": d.get("selection"), "resids": resids} for d in descs]
  task_labels = []
  for bid in range(n_slices):
    frames = xyz[bid * time_window:(bid + 1) * time_window]
    for s_, label in blocks:
      sub = frames[:, s_, :]
      focal = np.mean(sub[0], axis=0)
      for d, st in zip(descs, state):
        if st["len"] != sub.shape[1]:
          st.update(len=sub.shape[1], w=np.asarray(d["weights"])[s_], resids=resids[s_],
                    sel=(np.asarray(d["selection"])[s_] if d.get("selection") is not None else None))
        res = run_task(dict(d, weights=st["w"], selection=st["sel"]), st["resids"], sub.copy(), focal, buff_dim=buff_dim, kernels=kernels)
        if scrub_nan and np.isnan(res).any():
          res = np.nan_to_num(res).astype(res.dtype)
        out[d["outkey"]].append(res)
        task_labels.append(label)
  labels = np.array(task_labels, dtype=int)[:len(task_labels) // max(len(descs), 1)]      # featurizer.py:545-549
  return {k: np.array(v, dtype=np.float32) for k, v in out.items()}, labels


def superpose(xyz, fit_index, ref=0):
  """Least-squares fit of every frame onto frame `ref` over the atoms `fit_index` (Kabsch, SVD in float64):
  the textbook statement of pt.superpose(traj, mask) that MisatoTraj asks for
  (/root/reference/nearl/io/traj.py:398-404; cpptraj's own arithmetic is not in the reference tree, so this
  function is pinned by the algorithm's definition only).  Returns (fitted coordinates, per-frame RMSD)."""
  xyz = np.asarray(xyz, dtype=np.float64)
  out = xyz.copy()
  target = xyz[ref][fit_index]
  ct = target.mean(axis=0)
  rmsd = np.zeros(len(xyz))
  for f in range(len(xyz)):
    a = xyz[f][fit_index]
    ca = a.mean(axis=0)
    h = (a - ca).T @ (target - ct)
    u, _, vt = np.linalg.svd(h)
    d = np.sign(np.linalg.det(vt.T @ u.T))
    rot = vt.T @ np.diag([1.0, 1.0, d]) @ u.T
    fitted = (xyz[f] - ca) @ rot.T + ct
    rmsd[f] = np.sqrt(np.mean(np.sum((fitted[fit_index] - target) ** 2, axis=1)))
    if f != ref:
      out[f] = fitted
  return out, rmsd
