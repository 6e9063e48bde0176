"""Trajectory ingestion without pytraj: Amber parm7 topologies, the MISATO
MD.hdf5 layout (/root/reference/nearl/io/traj.py:287-423), the lazy loader, and the device superposition
(pt.superpose(traj, mask="@CA"), traj.py:398-404) against a float64 Kabsch fit."""
import gzip
import os

import numpy as np
import pytest

import nearl_b200
from nearl_b200 import _lib, constants, features, featurizer, io, utils
from oracle import host_ref

nearl_b200.update_config(quiet=True)


def write_parm7(path, top, with_atomic_number=True):
  """Minimal Amber parm7 with the sections the loader reads, in the file's own fixed-width formats."""
  blocks = top.residue_blocks()

  def sec(flag, fmt, vals, per, width, kind):
    out = [f"%FLAG {flag}", f"%FORMAT({fmt})"]
    for i in range(0, len(vals), per):
      row = vals[i:i + per]
      out.append("".join(f"{v:<{width}}" if kind == "a" else (f"{v:{width}d}" if kind == "i" else f"{v:{width}.8E}") for v in row))
    return out

  pointers = [0] * 31
  pointers[0], pointers[11] = top.n_atoms, len(blocks)
  lines = ["%VERSION  VERSION_STAMP = V0001.000  DATE = 01/01/24  00:00:00", "%FLAG TITLE", "%FORMAT(20a4)", "synthetic"]
  lines += sec("POINTERS", "10I8", pointers, 10, 8, "i")
  lines += sec("ATOM_NAME", "20a4", [str(x)[:4] for x in top.names], 20, 4, "a")
  lines += sec("CHARGE", "5E16.8", [float(q) * 18.2223 for q in top.charge], 5, 16, "e")
  lines += sec("MASS", "5E16.8", [float(constants.ATOMICMASS[int(z)]) for z in top.atomic_numbers], 5, 16, "e")
  if with_atomic_number:
    lines += sec("ATOMIC_NUMBER", "10I8", [int(z) for z in top.atomic_numbers], 10, 8, "i")
  lines += sec("RESIDUE_LABEL", "20a4", [b[0] for b in blocks], 20, 4, "a")
  lines += sec("RESIDUE_POINTER", "10I8", [b[1] + 1 for b in blocks], 10, 8, "i")
  opener = gzip.open if str(path).endswith(".gz") else open
  with opener(path, "wt") as fh:
    fh.write("\n".join(lines) + "\n")


def make_misato_dir(root, code, traj, solvent_atoms=9):
  """{root}/parameter_restart_files_MD/{code}/production.top.gz (with WAT / Na+ / Cl- appended) + {root}/MD.hdf5."""
  top = traj.top
  names = np.concatenate([top.names, np.array(["O", "H1", "H2"] * (solvent_atoms // 3), dtype=object), np.array(["Na+", "Cl-"], dtype=object)])
  n_w = solvent_atoms // 3
  last = int(top.resids.max())
  resids = np.concatenate([top.resids, np.repeat(np.arange(last + 1, last + 1 + n_w), 3), [last + n_w + 1, last + n_w + 2]])
  resnames = np.concatenate([top.resnames, np.array(["WAT"] * (3 * n_w) + ["Na+", "Cl-"], dtype=object)])
  z = np.concatenate([top.atomic_numbers, np.array([8, 1, 1] * n_w + [11, 17])])
  full = io.Topology(names, resids, resnames, z)
  os.makedirs(os.path.join(root, "parameter_restart_files_MD", code.lower()), exist_ok=True)
  write_parm7(os.path.join(root, "parameter_restart_files_MD", code.lower(), "production.top.gz"), full)
  with utils.open_hdf(os.path.join(root, "MD.hdf5"), "a") as hdf:
    grp = hdf.create_group(code.upper())
    grp.create_dataset("trajectory_coordinates", data=np.asarray(traj.xyz, dtype=np.float32))


def test_parm7_topology(tmp_path):
  traj = io.synthetic_peptide(14, 2, seed=5)
  for name, with_z in (("a.parm7", True), ("b.top.gz", False)):
    write_parm7(str(tmp_path / name), traj.top, with_atomic_number=with_z)
    top = io.load_amber_topology(str(tmp_path / name))
    assert top.n_atoms == traj.top.n_atoms
    assert list(top.names) == [str(n)[:4] for n in traj.top.names]
    assert np.array_equal(top.resids, traj.top.resids) and list(top.resnames) == list(traj.top.resnames)
    assert np.array_equal(top.atomic_numbers, traj.top.atomic_numbers)      # from the masses when the section is absent
  assert np.array_equal(top.select("@CA"), traj.top.select("@CA"))
  n = top.n_atoms
  top.strip(":LEU")
  assert top.n_atoms == n - np.count_nonzero(traj.top.resnames == "LEU") and "LEU" not in set(top.resnames)


def test_parm7_fast_parser_matches_the_line_parser(tmp_path):
  """The one-piece parser (sections found by position, fields through numpy, a gzip member inflated only as far as the
  last wanted section) against the line-by-line one, on the shapes of file it has to cope with."""
  traj = io.synthetic_peptide(57, 1, seed=6)
  path = str(tmp_path / "full.parm7")
  write_parm7(path, traj.top)
  text = open(path).read()
  # a long tail of sections the topology does not need, a %COMMENT line, CRLF line ends, a ragged section
  tail = "\n".join(["%FLAG BONDS_INC_HYDROGEN", "%FORMAT(10I8)"] + ["".join(f"{v:8d}" for v in range(k, k + 10)) for k in range(0, 200000, 10)]) + "\n"
  commented = text.replace("%FLAG CHARGE\n", "%FLAG CHARGE\n%COMMENT  Atomic charge in internal units\n")
  ragged = text.replace("%FLAG RESIDUE_POINTER\n%FORMAT(10I8)\n", "%FLAG RESIDUE_POINTER\n%FORMAT(10I8)\n       1\n", 1)
  assert ragged != text
  ragged = ragged.replace("%FLAG RESIDUE_POINTER\n%FORMAT(10I8)\n       1\n       1", "%FLAG RESIDUE_POINTER\n%FORMAT(10I8)\n       1\n        ", 1)
  variants = {"plain.parm7": text, "tail.top.gz": text + tail, "comment.parm7": commented, "crlf.parm7": text.replace("\n", "\r\n")}
  want = io.load_amber_topology(path)
  sec = io._parm7_sections(text)
  assert list(want.names) == io._parm7_values(sec["ATOM_NAME"])[:want.n_atoms]
  assert np.array_equal(want.charge, np.array(io._parm7_values(sec["CHARGE"])) / 18.2223)
  assert np.array_equal(want.atomic_numbers, np.array(io._parm7_values(sec["ATOMIC_NUMBER"])))
  assert all(type(n) is str for n in want.names[:3]) and all(type(n) is str for n in want.resnames[:3])
  for name, body in variants.items():
    opener = gzip.open if name.endswith(".gz") else open
    with opener(str(tmp_path / name), "wt", newline="") as fh:
      fh.write(body)
    got = io.load_amber_topology(str(tmp_path / name))
    assert list(got.names) == list(want.names) and list(got.resnames) == list(want.resnames), name
    assert np.array_equal(got.resids, want.resids) and np.array_equal(got.atomic_numbers, want.atomic_numbers), name
    assert np.array_equal(got.charge, want.charge), name
  # residues dropped while parsing = the topology strip() of each would leave (MisatoTraj: water and ions)
  solvated = str(tmp_path / "misato")
  make_misato_dir(solvated, "1abc", traj, solvent_atoms=12)
  topfile = os.path.join(solvated, "parameter_restart_files_MD", "1abc", "production.top.gz")
  stripped = io.load_amber_topology(topfile)
  assert stripped.n_atoms == want.n_atoms + 14
  for name in ("WAT", "Cl-", "Na+"):
    stripped.strip(f":{name}")
  dropped = io.load_amber_topology(topfile, drop_residues=("WAT", "Cl-", "Na+"))
  assert dropped.n_atoms == want.n_atoms and list(dropped.names) == list(stripped.names) and list(dropped.resnames) == list(stripped.resnames)
  assert np.array_equal(dropped.resids, stripped.resids) and np.array_equal(dropped.atomic_numbers, stripped.atomic_numbers)
  assert np.array_equal(dropped.charge, stripped.charge)
  assert io.load_amber_topology(topfile, drop_residues=("XYZ",)).n_atoms == want.n_atoms + 14
  head = io._parm7_head(str(tmp_path / "tail.top.gz"))
  assert len(head) < len(text + tail) and b"RESIDUE_POINTER" in head            # the tail was not inflated to its end
  # a section whose lines are not all full falls back to the line parser, section by section
  assert io._parm7_array(io._parm7_sections_bytes(ragged.encode(), ("RESIDUE_POINTER",))["RESIDUE_POINTER"]) is None
  open(str(tmp_path / "ragged.parm7"), "w").write(ragged)
  got = io.load_amber_topology(str(tmp_path / "ragged.parm7"))
  assert np.array_equal(got.resids, want.resids)


def test_loader_prefetch_reads_ahead_in_order(tmp_path):
  """TrajectoryLoader.prefetched: file halves of the constructors on loader threads, trajectories handed out in order,
  identical to the unthreaded loads; an unreadable entry raises when its turn comes; in-memory entries pass through."""
  root = str(tmp_path / "misato")
  codes = ["1abc", "2xyz", "3pqr", "4stu", "5vwx"]
  trajs = {code: io.synthetic_peptide(9 + i, 6, seed=40 + i) for i, code in enumerate(codes)}
  for code, traj in trajs.items():
    make_misato_dir(root, code, traj)
  memory = io.synthetic_peptide(7, 3, seed=50)
  loader = io.TrajectoryLoader([(c, root) for c in codes], trajtype=io.MisatoTraj, strides=[None, 2, None, None, 3], masks=[None, None, "@CA,C,N", None, None])
  loader.append([memory])
  order = [4, 0, 5, 2, 1, 3, 0]
  plain = [loader[i] for i in order]
  for workers in (0, 1, 3):
    got = list(loader.prefetched(order, workers=workers, ahead=2))
    assert [i for i, _ in got] == order
    for (_, a), b in zip(got, plain):
      assert a.identity == b.identity and a.xyz.dtype == b.xyz.dtype and np.array_equal(a.xyz, b.xyz)
      assert list(a.top.names) == list(b.top.names)
  assert got[2][1] is memory
  # several threads reading one open file at once give what one thread reads
  from concurrent.futures import ThreadPoolExecutor
  with ThreadPoolExecutor(6) as pool:
    raws = list(pool.map(lambda c: io.MisatoTraj.read_raw(c, root)["coord"], codes * 6))
  for code, coord in zip(codes * 6, raws):
    assert np.array_equal(coord, trajs[code].xyz.astype(np.float32))
  # the failure of a read surfaces at its own position, after the trajectories in front of it were delivered
  os.makedirs(os.path.join(root, "parameter_restart_files_MD", "9bad"))
  write_parm7(os.path.join(root, "parameter_restart_files_MD", "9bad", "production.top.gz"), trajs["1abc"].top)
  broken = io.TrajectoryLoader([("1abc", root), ("9bad", root), ("2xyz", root)], trajtype=io.MisatoTraj)
  seen = []
  with pytest.raises(ValueError, match="9BAD"):
    for i, t in broken.prefetched([0, 1, 2], workers=2):
      seen.append(i)
  assert seen == [0]


def test_misato_layout_and_loader(tmp_path):
  root = str(tmp_path / "misato")
  trajs = {code: io.synthetic_peptide(10 + i, 8, seed=20 + i) for i, code in enumerate(("1abc", "2XYZ"))}
  for code, traj in trajs.items():
    make_misato_dir(root, code, traj)
  loader = io.TrajectoryLoader([("1abc", root), ("2XYZ", root)], trajtype=io.MisatoTraj)
  assert len(loader) == 2
  for (code, _), got in zip(loader.trajs, loader):
    want = trajs[code]
    assert got.identity == code.lower() and got.n_atoms == want.n_atoms          # water and ions stripped
    assert np.array_equal(got.xyz, want.xyz.astype(np.float32)) and got.xyz.dtype == np.float32
    assert list(got.top.resnames) == list(want.top.resnames)
  strided = io.MisatoTraj("1abc", root, stride=3)
  assert strided.n_frames == 3 and np.array_equal(strided.xyz, trajs["1abc"].xyz[::3])
  picked = io.MisatoTraj("1abc", root, frame_indices=[5, 1], mask="@CA,C,N")
  assert picked.n_frames == 2 and picked.n_atoms == 30 and np.array_equal(picked.xyz[0], trajs["1abc"].xyz[5][trajs["1abc"].top.select("@CA,C,N")])
  os.makedirs(os.path.join(root, "parameter_restart_files_MD", "4kkk"))
  write_parm7(os.path.join(root, "parameter_restart_files_MD", "4kkk", "production.top.gz"), trajs["1abc"].top)
  with pytest.raises(ValueError):                 # a topology without coordinates in MD.hdf5
    io.MisatoTraj("4kkk", root)
  with pytest.raises(FileNotFoundError):
    io.MisatoTraj("9zzz", root)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_superpose_vs_kabsch(dtype):
  rng = np.random.default_rng(31)
  traj = io.synthetic_peptide(80, 1, seed=31, dtype=dtype)
  base = traj.xyz[0].astype(np.float64)
  frames = [base]
  for f in range(1, 12):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    rot = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                    [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                    [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    frames.append((base + rng.normal(0.0, 0.4, size=base.shape)) @ rot.T + rng.uniform(-30, 30, size=3))
  moved = io.Trajectory(np.ascontiguousarray(np.array(frames), dtype=dtype), traj.top, "moved")
  fit = moved.top.select("@CA")
  want, want_rmsd = host_ref.superpose(moved.xyz, fit, ref=0)
  before = moved.xyz.copy()
  rmsd = io.superpose(moved, mask="@CA")
  tol = 2e-4 if dtype == np.float32 else 1e-9            # float32 coordinates of magnitude ~100 A carry ~1e-5 A each
  assert np.array_equal(moved.xyz[0], before[0])            # the reference frame is untouched
  assert np.abs(moved.xyz - want).max() <= tol, np.abs(moved.xyz - want).max()
  assert np.allclose(rmsd, want_rmsd, atol=1e-3 if dtype == np.float32 else 1e-6)
  assert rmsd[0] <= 1e-3 and 0.3 < rmsd[1:].mean() < 1.0
  # idempotent: a second fit moves nothing beyond rounding
  again = moved.xyz.copy()
  io.superpose(moved, mask="@CA")
  assert np.abs(moved.xyz - again).max() <= 10 * tol


@pytest.mark.gpu
def test_misato_trajectory_through_the_featurizer(tmp_path):
  root = str(tmp_path / "misato")
  make_misato_dir(root, "3eov", io.synthetic_peptide(60, 20, seed=33))
  loader = io.TrajectoryLoader([("3EOV", root)], trajtype=io.MisatoTraj, superpose=True)
  fz = featurizer.Featurizer({"dimensions": 16, "lengths": 12, "time_window": 10, "cutoff": 3.0, "sigma": 1.0})
  fz.register_features([features.Mass(outkey="mass"), features.MarchingObservers(obs="existence", agg="mean", outkey="exist")])
  fz.register_focus([":ARG,LYS,GLU"], "mask")
  fz.register_trajloader(loader)
  fz.keep_results = True
  stats = fz.run()
  assert stats["native_tasks"] == 4 and np.array(fz.RESULTS["mass"]).shape == (2, 16, 16, 16)
  assert np.array(fz.RESULTS["mass"]).sum() > 0


@pytest.mark.gpu
def test_misato_run_with_and_without_loader_threads(tmp_path):
  """Featurizer.run() over file-backed MISATO trajectories (superposed on the device by their constructor, on the calling
  thread): the read-ahead threads change nothing in the results, whatever their number."""
  root = str(tmp_path / "misato")
  codes = [f"{i}abc" for i in range(1, 8)]
  for i, code in enumerate(codes):
    make_misato_dir(root, code, io.synthetic_peptide(40 + 3 * i, 20, seed=60 + i))
  outs = []
  for workers in (0, 1, 4):
    loader = io.TrajectoryLoader([(c, root) for c in codes], trajtype=io.MisatoTraj, superpose=True)
    fz = featurizer.Featurizer({"dimensions": 16, "lengths": 12, "time_window": 10, "cutoff": 3.0, "sigma": 1.0})
    fz.register_features([features.Mass(outkey="mass"), features.MarchingObservers(obs="mean_distance", agg="standard_deviation", outkey="md")])
    fz.register_focus([":ARG,LYS,GLU"], "mask")
    fz.register_trajloader(loader)
    fz.keep_results = True
    fz.prefetch_workers = workers
    stats = fz.run()
    assert stats["trajectories"] == len(codes)
    outs.append({k: np.array(v) for k, v in fz.RESULTS.items()})
  for other in outs[1:]:
    for key in outs[0]:
      assert np.array_equal(outs[0][key], other[key]), key


PDB_TEXT = """\
HEADER    TEST
MODEL        1
ATOM      1  N   ALA A   1      11.104   6.134  -6.504  1.00  0.00           N
ATOM      2  CA  ALA A   1      11.639   6.071  -5.147  1.00  0.00           C
ATOM      3  C   ALA A   1      10.573   5.700  -4.127  1.00  0.00           C
ATOM      4  HA AALA A   1      12.400   5.300  -5.100  0.50  0.00
ATOM      5  HA BALA A   1      12.410   5.310  -5.110  0.50  0.00
ATOM      6  N   GLY A   2       9.300   5.900  -4.500  1.00  0.00           N
ATOM      7  CA  GLY A   2       8.200   5.600  -3.600  1.00  0.00           C
ATOM      8  N   GLY A   2A      7.100   5.000  -3.000  1.00  0.00           N
TER       9      GLY A   2A
ATOM     10  N   SER B   2       1.000   2.000   3.000  1.00  0.00           N
HETATM   11 CA    CA B 101       0.500   0.250  -0.125  1.00  0.00
HETATM   12 CL    CL B 102       4.000   4.000   4.000  1.00  0.00          CL
HETATM   13  O   HOH B 201       5.000   5.000   5.000  1.00  0.00           O
HETATM   14 1H   HOH B 201       5.500   5.500   5.500  1.00  0.00
ENDMDL
MODEL        2
ATOM      1  N   ALA A   1      12.104   6.134  -6.504  1.00  0.00           N
ATOM      2  CA  ALA A   1      12.639   6.071  -5.147  1.00  0.00           C
ATOM      3  C   ALA A   1      11.573   5.700  -4.127  1.00  0.00           C
ATOM      4  HA AALA A   1      13.400   5.300  -5.100  0.50  0.00
ATOM      5  HA BALA A   1      13.410   5.310  -5.110  0.50  0.00
ATOM      6  N   GLY A   2      10.300   5.900  -4.500  1.00  0.00           N
ATOM      7  CA  GLY A   2       9.200   5.600  -3.600  1.00  0.00           C
ATOM      8  N   GLY A   2A      8.100   5.000  -3.000  1.00  0.00           N
TER       9      GLY A   2A
ATOM     10  N   SER B   2       2.000   2.000   3.000  1.00  0.00           N
HETATM   11 CA    CA B 101       1.500   0.250  -0.125  1.00  0.00
HETATM   12 CL    CL B 102       5.000   4.000   4.000  1.00  0.00          CL
HETATM   13  O   HOH B 201       6.000   5.000   5.000  1.00  0.00           O
HETATM   14 1H   HOH B 201       6.500   5.500   5.500  1.00  0.00
ENDMDL
END
"""


def test_pdb_and_netcdf_files_without_pytraj(tmp_path):
  """Trajectory("traj.nc", "top.parm7" | "top.pdb") and Trajectory("complex.pdb") (/root/reference/nearl/io/traj.py:97-106 hands
  these to pt.load): PDB by columns -- residues from (chain, number, insertion code, name), elements from columns 77-78
  or the name columns (' CA ' is a C-alpha, 'CA  ' in residue CA is calcium), one frame per MODEL --, Amber NetCDF through
  scipy.  stride wins over frame_indices, masks apply to coordinates and topology, the identity defaults to the
  trajectory file name."""
  from scipy.io import netcdf_file
  pdb = tmp_path / "multi.pdb"
  pdb.write_text(PDB_TEXT)
  traj = io.Trajectory(str(pdb))
  assert traj.n_frames == 2 and traj.n_atoms == 13 and traj.xyz.dtype == np.float64
  assert traj.identity == str(pdb) and traj.top_filename is None
  assert list(traj.top.names) == ["N", "CA", "C", "HA", "HA", "N", "CA", "N", "N", "CA", "CL", "O", "1H"]
  assert list(traj.top.resids) == [0, 0, 0, 0, 0, 1, 1, 2, 3, 4, 5, 6, 6]               # insertion code and chain start new residues
  assert list(traj.top.resnames[[0, 5, 8, 9, 10, 12]]) == ["ALA", "GLY", "SER", "CA", "CL", "HOH"]
  assert list(traj.top.atomic_numbers) == [7, 6, 6, 1, 1, 7, 6, 7, 7, 20, 17, 8, 1]
  assert np.array_equal(traj.xyz[0, 1], [11.639, 6.071, -5.147]) and np.array_equal(traj.xyz[1, 9], [1.5, 0.25, -0.125])
  assert np.array_equal(traj.xyz[1, :, 0] - traj.xyz[0, :, 0], np.ones(13))
  assert list(traj.top.select(":ALA@CA")) == [1] and list(traj.top.select(":HOH")) == [11, 12]
  assert len(traj.atoms) == 13 and len(traj.residues) == 7 and traj.residues[0].name == "ALA"
  lig = traj[":CA,CL"]
  assert lig.n_atoms == 2 and lig.n_frames == 2 and list(lig.top.resnames) == ["CA", "CL"]
  # NetCDF trajectory (float32, as sander / pmemd write it) + PDB topology, then + parm7 topology
  pep = io.synthetic_peptide(9, 7, seed=3)
  nc_path = str(tmp_path / "traj.nc")
  with netcdf_file(nc_path, "w") as nc:
    nc.Conventions = "AMBER"
    nc.createDimension("frame", pep.n_frames); nc.createDimension("atom", pep.n_atoms); nc.createDimension("spatial", 3)
    var = nc.createVariable("coordinates", "f", ("frame", "atom", "spatial"))
    var[:] = pep.xyz.astype(np.float32)
  top_pdb = str(tmp_path / "top.pdb")
  io.write_pdb(top_pdb, pep.xyz[0], pep.top)
  a = io.Trajectory(nc_path, top_pdb)
  assert a.xyz.dtype == np.float64 and a.xyz.shape == pep.xyz.shape and np.array_equal(a.xyz, pep.xyz.astype(np.float32).astype(np.float64))
  assert list(a.top.names) == [str(n) for n in pep.top.names] and np.array_equal(a.top.atomic_numbers, pep.top.atomic_numbers)
  assert np.array_equal(a.top.resids, pep.top.resids - pep.top.resids.min()) and list(a.top.resnames) == list(pep.top.resnames)
  assert a.identity == nc_path and a.top_filename == top_pdb and a.traj_filename == nc_path
  parm = str(tmp_path / "top.parm7")
  write_parm7(parm, pep.top)
  b = io.Trajectory(nc_path, parm, stride=3, frame_indices=[0, 1], dtype=np.float32, identity="pep")
  assert b.xyz.dtype == np.float32 and b.n_frames == 3 and np.array_equal(b.xyz, pep.xyz.astype(np.float32)[::3]) and b.identity == "pep"
  c = io.Trajectory(nc_path, parm, frame_indices=[5, 2], mask="@CA")
  sel = pep.top.select("@CA")
  assert c.n_frames == 2 and np.array_equal(c.xyz, pep.xyz.astype(np.float32).astype(np.float64)[[5, 2]][:, sel]) and list(c.top.names) == ["CA"] * len(sel)
  # restart-style file: one frame of doubles
  rst = str(tmp_path / "last.ncrst")
  with netcdf_file(rst, "w") as nc:
    nc.createDimension("atom", pep.n_atoms); nc.createDimension("spatial", 3)
    nc.createVariable("coordinates", "d", ("atom", "spatial"))[:] = pep.xyz[3].astype(np.float64)
  d = io.Trajectory(rst, parm)
  assert d.n_frames == 1 and np.array_equal(d.xyz[0], pep.xyz[3].astype(np.float64))
  # round trip through the PDB writer, frames and masks
  out = str(tmp_path / "frames.pdb")
  a.write_frame([0, 4], out, mask="@CA,C,N")
  e = io.Trajectory(out)
  keep = pep.top.select("@CA,C,N")
  assert e.n_frames == 2 and np.allclose(e.xyz, a.xyz[[0, 4]][:, keep], atol=5e-4) and list(e.top.names) == [str(n) for n in pep.top.names[keep]]
  with pytest.raises(ValueError):
    io.Trajectory(nc_path, str(pdb))                                # 13-atom topology, other coordinates
  with pytest.raises(ValueError):
    io.load_coordinates(str(tmp_path / "top.parm7"))


def test_pdb_column_parser_of_the_library_matches_the_line_parser(tmp_path):
  """nrl_pdb_parse (host code of the native library) + the per-distinct-value topology against io._load_pdb_lines, the
  definition: same coordinates to the bit, same names / residues / elements -- and the texts it declines (None) are left
  to the line parser."""
  def same(text, name="t.pdb"):
    path = str(tmp_path / name)
    with open(path, "wb") as fh:
      fh.write(text if isinstance(text, bytes) else text.encode())
    fast = io._load_pdb_native(open(path, "rb").read())
    assert fast is not None, name
    xyz, top = io._load_pdb_lines(path)
    assert fast[0].dtype == np.float64 and np.array_equal(fast[0], xyz), name
    assert list(fast[1].names) == list(top.names) and list(fast[1].resnames) == list(top.resnames), name
    assert np.array_equal(fast[1].resids, top.resids) and np.array_equal(fast[1].atomic_numbers, top.atomic_numbers), name
    assert all(type(v) is str for v in fast[1].names[:2])
    got = io.load_pdb(path)
    assert np.array_equal(got[0], xyz) and list(got[1].names) == list(top.names)
    return xyz, top
  xyz, top = same(PDB_TEXT)
  assert xyz.shape == (2, 13, 3)
  same(PDB_TEXT.replace("\n", "\r\n"), "crlf.pdb")
  same(PDB_TEXT.rstrip("\n"), "no_final_newline.pdb")
  same(PDB_TEXT.replace("ENDMDL\n", "").replace("END\n", "").split("MODEL        2")[0], "no_end.pdb")
  # lines cut inside / in front of the element columns, negative and unsigned-plus coordinates, no decimals
  atom = "ATOM      1 FE   HEM A   1     -11.104  +6.134      -6  1.00  0.00          FE"
  for cut in (54, 60, 66, 76, 77, 78, 79):
    same("\n".join(line[:cut] for line in (atom, atom.replace("FE   HEM", " CA  ALA"), atom.replace("FE   HEM", "CA    CA"))) + "\n", f"cut{cut}.pdb")
    same("\n".join(line[:cut] for line in (atom, atom.replace(" 1 FE", " 2 FE"))), f"cut{cut}_open.pdb")
  pep = io.synthetic_peptide(30, 3, seed=8, break_every=7)
  io.write_pdb(str(tmp_path / "pep.pdb"), pep.xyz, pep.top)
  same(open(str(tmp_path / "pep.pdb"), "rb").read(), "pep_again.pdb")
  # declined: the line parser decides (and words the error where there is one)
  for text in (atom.replace("-11.104", "-1.1e+01"), atom[:50], atom.replace("HEM", "H\u00c9M"), atom.replace("-11.104", "        "),
               atom + "\r" + atom):
    assert io._load_pdb_native(text.encode()) is None, text[:40]
  lopsided = PDB_TEXT.split("\n")
  del lopsided[[i for i, l in enumerate(lopsided) if l.startswith("HETATM   14")][1]]
  assert io._load_pdb_native("\n".join(lopsided).encode()) is None
  open(str(tmp_path / "lopsided.pdb"), "w").write("\n".join(lopsided))
  with pytest.raises(ValueError, match="same number of atoms"):
    io.load_pdb(str(tmp_path / "lopsided.pdb"))
  open(str(tmp_path / "empty.pdb"), "w").write("HEADER\nEND\n")
  with pytest.raises(ValueError, match="no ATOM"):
    io.load_pdb(str(tmp_path / "empty.pdb"))
  with gzip.open(str(tmp_path / "z.pdb.gz"), "wt") as fh:
    fh.write(PDB_TEXT)
  assert np.array_equal(io.load_pdb(str(tmp_path / "z.pdb.gz"))[0], xyz)


def test_trajectory_loader_options(tmp_path):
  """TrajectoryLoader (/root/reference/nearl/io/trajloader.py:12-247): lazy construction with the common loading options,
  per-trajectory strides / masks / identities, integer / list / slice indexing, append, ready-made trajectory objects."""
  pep = io.synthetic_peptide(8, 6, seed=9)
  pdb = str(tmp_path / "p.pdb")
  io.write_pdb(pdb, pep.xyz, pep.top)
  loader = io.TrajectoryLoader([(pdb,), (pdb,), [pdb]], trajids=["a", "b", None], stride=2, strides=[None, 3, None], masks=[None, None, "@CA"])
  assert len(loader) == 3 and loader.get_loading_options(1) == {"stride": 3, "identity": "b"}
  t0, t1, t2 = loader[0], loader[1], loader[-1]
  assert (t0.n_frames, t1.n_frames, t2.n_frames) == (3, 2, 3) and (t0.identity, t1.identity, t2.identity) == ("a", "b", pdb)
  assert t2.n_atoms == len(pep.top.select("@CA")) and t0.n_atoms == pep.n_atoms and loader.i_ == 2
  assert [t.identity for t in loader[[0, 1]]] == ["a", "b"] and [t.n_frames for t in loader[1:]] == [2, 3]
  assert [t.n_frames for t in loader] == [3, 2, 3]
  loader.append([(pdb,)], trajids=["late"], masks=["@N"])
  loader.push(pep)                                          # an object that already is a trajectory is handed out as it is
  assert len(loader) == 5 and loader[3].identity == "late" and loader[3].n_atoms == len(pep.top.select("@N")) and loader[4] is pep
  assert loader.loading_options["stride"] == 2 and loader.trajtype is io.Trajectory
  with pytest.raises(ValueError):
    io.TrajectoryLoader([(pdb,)], trajids=["a", "b"])
  with pytest.raises(ValueError):
    io.TrajectoryLoader([(pdb,)], strides=[1, 2])
  with pytest.raises(IndexError):
    loader["x"]
  # read ahead on loader threads: the same trajectories, identities and options; a subclass with a constructor of its own
  # does not take part (it is built on access, on the calling thread)
  ahead = list(loader.prefetched(range(len(loader)), workers=2))
  for i, t in ahead:
    same = loader[i]
    assert t.identity == same.identity and t.n_frames == same.n_frames and np.array_equal(t.xyz, same.xyz) and t.xyz.dtype == same.xyz.dtype
    assert t.traj_filename == same.traj_filename and list(t.top.names) == list(same.top.names)
  assert ahead[4][1] is pep

  class Own(io.Trajectory):
    def __init__(self, path, scale):
      super().__init__(path)
      self.xyz = self.xyz * scale
  own = io.TrajectoryLoader([(pdb, 2.0)], trajtype=Own)
  assert np.array_equal(next(iter(own.prefetched([0], workers=2)))[1].xyz, io.Trajectory(pdb).xyz * 2.0)


def test_partial_charge_weights(tmp_path):
  """PartialCharge (/root/reference/nearl/features.py:895-1031): charges of the topology (parm7 CHARGE / 18.2223), of a
  table keyed by the trajectory's identity, of a user function; keep_sign; the reference's errors."""
  from nearl_b200 import features, featurizer
  pep = io.synthetic_peptide(7, 2, seed=21)
  rng = np.random.default_rng(2)
  pep.top.charge = np.round(rng.normal(0, 0.4, size=pep.n_atoms), 4)
  parm = str(tmp_path / "q.parm7")
  write_parm7(parm, pep.top)
  top = io.load_amber_topology(parm)
  assert np.allclose(top.charge, pep.top.charge, atol=1e-6) and np.array_equal(top[3:9].charge, top.charge[3:9])
  traj = io.Trajectory(pep.xyz, top, identity="pep")
  fz = featurizer.Featurizer({"dimensions": 8, "lengths": 8, "time_window": 1, "cutoff": 2.0, "sigma": 1.0})
  feats = {"both": features.PartialCharge(outkey="q"), "pos": features.PartialCharge(keep_sign="p", outkey="qp"),
           "neg": features.PartialCharge(keep_sign="negative", outkey="qn"),
           "manual": features.PartialCharge(charge_type="manual", charge_parm={"pep": list(range(pep.n_atoms))}, outkey="qm"),
           "ext": features.PartialCharge(charge_type="external", charge_parm=lambda t: np.full(t.top.n_atoms, -0.5), keep_sign="p", outkey="qe")}
  fz.register_features(feats)
  for f in feats.values():
    f.cache(traj)
  q = top.charge.astype(np.float32)
  assert np.array_equal(feats["both"].cached_array, q) and feats["both"].cached_array.dtype == np.float32
  assert np.array_equal(feats["pos"].cached_array, np.maximum(q, 0)) and np.array_equal(feats["neg"].cached_array, np.minimum(q, 0))
  assert np.array_equal(feats["manual"].cached_array, np.arange(pep.n_atoms, dtype=np.float32)) and not feats["ext"].cached_array.any()
  assert str(feats["neg"]) == "PartialCharge <type:topology|sign:negative>"
  coords, weights = feats["both"].query(top, pep.xyz[0], np.asarray(pep.xyz[0].mean(axis=0), dtype=np.float32))
  assert len(coords) == len(weights) > 0 and set(np.round(weights, 4)) <= set(np.round(q, 4))
  with pytest.raises(ValueError):
    features.PartialCharge(outkey="z", **{}).cache(io.Trajectory(pep.xyz, io.Topology(top.names, top.resids, top.resnames, top.atomic_numbers)))
  with pytest.raises(ValueError):
    features.PartialCharge(charge_type="mulliken")
  with pytest.raises(ValueError):
    features.PartialCharge(keep_sign="x")
  cf = features.PartialCharge(charge_type="chargefw", charge_parm="eem", outkey="c")
  fz.register_feature(cf)
  with pytest.raises(NotImplementedError):
    cf.cache(traj)


@pytest.mark.gpu
def test_static_groups_from_pdb_files_with_and_without_loader_threads(tmp_path):
  """PDBBind-style run from PDB files (single-frame systems grouped per native call): the loader threads and the column
  parser of the library leave the results what the line parser and no threads give."""
  paths = []
  for i in range(37):
    c = io.synthetic_complex(300 + 7 * i, seed=70 + i)
    paths.append(str(tmp_path / f"c{i}.pdb"))
    io.write_pdb(paths[-1], c.xyz, c.top)
  outs = []
  for workers, parser in ((0, "lines"), (0, "native"), (4, "native")):
    if parser == "lines":
      trajs = [io.Trajectory(*io._load_pdb_lines(p), identity=p) for p in paths]
      loader = io.TrajectoryLoader(trajs)
    else:
      loader = io.TrajectoryLoader([(p,) for p in paths])
    fz = featurizer.Featurizer({"dimensions": 16, "lengths": 12, "time_window": 1, "cutoff": 2.5, "sigma": 1.5})
    fz.register_features([features.Mass(outkey="mass"), features.HeavyAtom(outkey="heavy"), features.AtomType(focus_element=7, outkey="n")])
    fz.register_focus([":LIG"], "mask")
    fz.register_trajloader(loader)
    fz.keep_results = True
    fz.prefetch_workers = workers
    fz.static_group = 8
    stats = fz.run()
    assert stats["trajectories"] == len(paths)
    outs.append({k: np.array(v) for k, v in fz.RESULTS.items()})
  for other in outs[1:]:
    for key in outs[0]:
      assert outs[0][key].shape[0] == len(paths) and np.array_equal(outs[0][key], other[key]), key
