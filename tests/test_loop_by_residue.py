"""Featurizer.loop_by_residue (/root/reference/nearl/featurizer.py:471-554): block enumeration and
labels on the CPU; on the GPU the batched device path against the numpy restatement of the
reference's loop feeding the CPU oracle (oracle/host_ref.featurize_by_residue) and against this
package's own call-by-call host protocol."""
import numpy as np
import pytest

import nearl_b200
from nearl_b200 import _lib, constants, features, featurizer, io, utils
from oracle import host_ref

nearl_b200.update_config(quiet=True)

PARMS = {"dimensions": 16, "lengths": 12, "time_window": 3, "cutoff": 3.0, "sigma": 1.0}


def rel_bad(new, ref, rtol=1e-5, atol=0.0):
  ref = np.asarray(ref, dtype=np.float64); new = np.asarray(new, dtype=np.float64)
  scale = np.maximum(np.abs(ref), 1e-5 * max(np.abs(ref).max(), 1e-30))
  return int((np.abs(new - ref) > rtol * scale + atol).sum())


def test_label_tables_follow_the_reference():
  # nearl/constants.py:218-237 (single), :240-243,283-286 (pairs, row-major over RES x RES), :233-237 (patches)
  assert constants.RES2LAB["ALA"] == 0 and constants.RES2LAB["VAL"] == 19 and constants.LAB2RES[8] == "HIS"
  assert constants.RES2LAB_DUAL["ALAALA"] == 0 and constants.RES2LAB_DUAL["ALAVAL"] == 19 and constants.RES2LAB_DUAL["ARGALA"] == 20
  assert constants.RES2LAB_DUAL["ILEVAL"] == 199 and constants.LAB2RES_DUAL[140] == "GLYALA" and len(constants.RES2LAB_DUAL) == 400
  assert constants.RES_PATCH["CYX"] == "CYS" and constants.RES_PATCH["HSP"] == "HIS" and len(constants.RES_PATCH) == 7


def test_residue_blocks_single_and_dual():
  traj = io.synthetic_peptide(60, 6, seed=3, break_every=7)
  res = list(traj.top.residues)
  assert res[0].first_atom_index == 0 and res[-1].last_atom_index == traj.n_atoms
  assert all(a.last_atom_index == b.first_atom_index for a, b in zip(res, res[1:]))
  fz = featurizer.Featurizer(PARMS)
  fz.traj = traj
  single = fz._residue_blocks("single")
  assert len(single) == 60
  # residue names in RES + RES_PATCH order, blocks of one name in topology order, patched names fold onto the parent
  order = list(constants.RES) + list(constants.RES_PATCH)
  names = [traj.top.resnames[s.start] for s, _ in single]
  assert [order.index(n) for n in names] == sorted(order.index(n) for n in names)
  for (s, label), name in zip(single, names):
    assert label == constants.RES2LAB[constants.RES_PATCH.get(name, name)]
  assert any(n in constants.RES_PATCH for n in names)
  limited = fz._residue_blocks("single", tag_limit=1)
  assert len(limited) == len(set(names))
  # pairs: consecutive residues with a peptide bond in frame 0; the chain is broken after every 7th residue
  by_key = utils.dual_blocks_by_key(traj)
  n_pairs = sum(len(v) for v in by_key.values())
  assert n_pairs == 59 - len([r for r in range(1, 60) if r % 7 == 0])
  for key, blocks in by_key.items():
    assert [(s.start, s.stop) for s in utils.find_block_dual(traj, key)] == [(s.start, s.stop) for s in blocks]
  dual = fz._residue_blocks("dual")
  assert len(dual) == n_pairs
  for s, label in dual:
    n1, n2 = traj.top.resnames[s.start], traj.top.resnames[s.stop - 1]
    assert label == constants.RES2LAB_DUAL[constants.RES_PATCH.get(n1, n1) + constants.RES_PATCH.get(n2, n2)]
  with pytest.raises(ValueError):
    fz._residue_blocks("triple")


def make_features():
  return [
    features.Mass(outkey="mass"),
    features.HeavyAtom(outkey="heavy", byres=True, padding=1.0),
    features.MarchingObservers(obs="direct_count", agg="max", weight_type="mass", outkey="count"),
    features.MarchingObservers(obs="distinct_count", agg="mean", weight_type="atomic_number", outkey="distinct"),
    features.MarchingObservers(obs="mean_distance", agg="standard_deviation", weight_type="mass", outkey="mean_distance"),
    features.DensityFlow(agg="mean", weight_type="electronegativity", outkey="flow"),
  ]


def descs_from(fz):
  descs = []
  for f in fz.FEATURESPACE:
    d = {"outkey": f.outkey, "dims": f.dims, "spacing": f.spacing, "cutoff": f.cutoff, "sigma": f.sigma, "padding": f.padding,
         "byres": f.byres, "selection": (f.selected if f.selection is not None else None), "weights": f.cached_array,
         "frame_offset": f.frame_offset}
    if isinstance(f, features.MarchingObservers):
      d.update(kind="marching_observers", obs=f.obs, agg=f.agg)
    elif isinstance(f, features.DensityFlow):
      d.update(kind="density_flow", agg=f.agg)
    else:
      d.update(kind="static")
    descs.append(d)
  return descs


@pytest.mark.gpu
@pytest.mark.parametrize("restype,dtype,per_chunk", [("single", np.float32, 10), ("dual", np.float32, 9), ("single", np.float64, 20),
                                                     ("single", np.float32, 7), ("single", np.float32, 64)])
def test_loop_by_residue_vs_reference_loop(tmp_path, restype, dtype, per_chunk):
  """`per_chunk` blocks per native call: 40 blocks x 2 windows make 4 / 2 / 6 / 1 chunks per window (an even count sent the
  first chunk of a window to the lane still holding the last chunk of the window before)."""
  traj = io.synthetic_peptide(40, 6, seed=11, dtype=dtype, break_every=9)
  out = str(tmp_path / "residues.h5")
  fz = featurizer.Featurizer({**PARMS, "outfile": out})
  fz.register_features(make_features())
  fz.register_trajloader([traj])
  launches = _lib.kernel_launches()
  calls = []
  real = featurizer.LoweredBatch.execute
  featurizer.LoweredBatch.execute = lambda self, *a, **kw: calls.append(kw.get("lane")) or real(self, *a, **kw)
  try:
    stats = fz.loop_by_residue(restype, process_nr=4, chunk_floats=per_chunk * 6 * 16 ** 3)
  finally:
    featurizer.LoweredBatch.execute = real
  assert calls == [i % 2 for i in range(len(calls))]                  # one job pending: strictly alternating lanes
  if restype == "single":
    assert len(calls) == 2 * -(-40 // per_chunk)
  fz.traj = traj
  blocks = fz._residue_blocks(restype)
  assert stats["tasks"] == 2 * len(blocks) * 6 and stats["native_tasks"] == stats["tasks"]
  assert _lib.kernel_launches() > launches
  for f in fz.FEATURESPACE:
    f.cache(traj)
  ref, labels = host_ref.featurize_by_residue(traj.xyz, fz.FEATURESPACE[0].resids, descs_from(fz), blocks, 3, buff_dim=_lib.buff_dim)
  with utils.open_hdf(out, "r") as h:
    got = {k: h[k][:] for k in h.keys() if k != "featurizer_parms"}
  assert np.array_equal(got.pop("label"), labels) and labels.dtype.kind == "i"
  for key, want in ref.items():
    assert got[key].shape == want.shape == (2 * len(blocks), 16, 16, 16)
    if key in ("count", "distinct"):
      assert np.allclose(got[key], want, rtol=1e-6, atol=1e-6) if key == "distinct" else np.array_equal(got[key], want), key
    else:
      atol = 1e-5 * float(np.abs(want).max()) if key in ("mean_distance", "flow") else 0.0     # reducers that cancel
      assert rel_bad(got[key], want, atol=atol) == 0, (key, np.abs(got[key] - want).max())
  # the reference re-caches a feature only when the block length changes: some blocks must have been read
  # with another block's tables (otherwise this test would not pin that rule)
  lens = [s.stop - s.start for s, _ in blocks]
  assert any(a == b for a, b in zip(lens, lens[1:])) and any(a != b for a, b in zip(lens, lens[1:]))


@pytest.mark.gpu
def test_loop_by_residue_host_protocol_matches_device_path(tmp_path):
  traj = io.synthetic_peptide(24, 3, seed=12, break_every=0)
  rows = {}
  for native in (True, False):
    fz = featurizer.Featurizer(PARMS)
    fz.use_native_batch = native
    fz.keep_results = True
    feats = [features.Mass(outkey="mass"), features.MarchingObservers(obs="existence", agg="mean", outkey="exist"),
             # residue_id numbers residues inside the slice: host protocol on both sides
             features.MarchingObservers(obs="distinct_count", agg="max", weight_type="residue_id", outkey="resid")]
    fz.register_features(feats)
    fz.register_trajloader([traj])
    stats = fz.loop_by_residue("dual", tag_limit=0)
    assert (stats["native_tasks"] > 0) == native and stats["native_tasks"] < stats["tasks"]
    rows[native] = {k: np.array(v) for k, v in fz.RESULTS.items()}
  assert set(rows[True]) == {"mass", "exist", "resid", "label"}
  assert np.array_equal(rows[True]["label"], rows[False]["label"])
  assert rel_bad(rows[True]["mass"], rows[False]["mass"]) == 0
  assert np.array_equal(rows[True]["exist"], rows[False]["exist"]) and np.array_equal(rows[True]["resid"], rows[False]["resid"])
  assert rows[True]["resid"].max() == 2.0          # a pair of residues: two distinct residue ids
