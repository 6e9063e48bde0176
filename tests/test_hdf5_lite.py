"""hdf5_lite: appends cost O(log rows) (metadata updated in place, B-tree right edge only), files re-open lazily, and the
classic-layout reader agrees with a file written by libhdf5 itself."""
import glob
import os

import numpy as np
import pytest

from nearl_b200 import hdf5_lite, utils


def _rows(n, shape=(2, 2, 2), seed=0):
  return np.random.default_rng(seed).random((n, *shape)).astype(np.float32)


def test_append_is_in_place_and_logarithmic(tmp_path):
  """6 000 one-row chunks: a three-level B-tree (64 children per node).  The file holds the data once plus one node per 64
  chunks; what an append WRITES does not grow with the number of rows already there."""
  path = str(tmp_path / "big.h5")
  rows = _rows(6000)
  written = []
  for start in range(0, 6000, 100):
    with hdf5_lite.File(path, "a") as h:
      utils._append_now(h, "feat", rows[start:start + 100], np.float32, (None, 2, 2, 2), chunks=(1, 2, 2, 2))
      if start % 1000 == 0:
        utils._append_now(h, "gz", rows[start:start + 100], np.float32, (None, 2, 2, 2), chunks=(1, 2, 2, 2), compression="gzip",
                          compression_opts=4)
    written.append(h.stats["bytes_written"])
  plain = [w for i, w in enumerate(written) if i % 10]             # the appends that did not touch "gz" as well
  assert max(plain[27:]) <= 1.3 * max(plain[:9])                   # bytes per append are flat (a new internal node now and then)
  with hdf5_lite.File(path, "r") as h:
    read_at_open = h.stats["bytes_read"]
    assert h["feat"].shape == (6000, 2, 2, 2) and h["feat"].chunks == (1, 2, 2, 2) and h["feat"].maxshape == (None, 2, 2, 2)
    assert np.array_equal(h["feat"][4321], rows[4321])
    assert h.stats["bytes_read"] - read_at_open < 4 * h["feat"]._index.node_len + 64      # one node per level + the chunk
    assert np.array_equal(h["feat"][:], rows)
    assert np.array_equal(h["gz"][:], np.concatenate([rows[s:s + 100] for s in range(0, 6000, 1000)]))
    entries = list(h["feat"]._index.walk())
    assert [e[0][0] for e in entries] == list(range(6000))
    assert h["feat"]._index._node(h["feat"]._index.root).level == 2
  assert read_at_open < 2048                                      # headers only: no B-tree node, no raw data
  data = 6000 * 32 + 600 * 32
  nodes = (6000 // 64 + 6000 // 4096 + 4 + 600 // 64 + 2) * h["feat"]._index.node_len
  assert os.path.getsize(path) < data + nodes + 16384              # data once + one node per 64 chunks


def test_overwrite_gap_fill_and_group_growth(tmp_path):
  path = str(tmp_path / "o.h5")
  rows = _rows(300, seed=1)
  with hdf5_lite.File(path, "w") as h:
    d = h.create_dataset("a", (300, 2, 2, 2), dtype=np.float32, maxshape=(None, 2, 2, 2), chunks=(1, 2, 2, 2))
    d[200:300] = rows[200:300]          # out of order: the front of the dataset is a gap ...
    d[0:100] = rows[0:100]              # ... filled afterwards (tree rebuilt around the sorted entries)
    z = h.create_dataset("z", (10, 2, 2, 2), dtype=np.float32, maxshape=(None, 2, 2, 2), chunks=(4, 2, 2, 2), compression="gzip")
    z[0:10] = rows[0:10]
  size0 = os.path.getsize(path)
  with hdf5_lite.File(path, "a") as h:
    assert np.array_equal(h["a"][0:100], rows[0:100]) and np.array_equal(h["a"][200:300], rows[200:300])
    assert not h["a"][100:200].any()                                 # never written: fill value
    h["a"][10:20] = rows[110:120]                                    # same size, no filter: chunks rewritten where they are
    h["z"][3:6] = rows[50:53]                                        # partial chunks of a compressed dataset
  assert os.path.getsize(path) - size0 < 1024
  with hdf5_lite.File(path, "a") as h:
    assert np.array_equal(h["a"][10:20], rows[110:120]) and np.array_equal(h["a"][20:100], rows[20:100])
    expect = rows[0:10].copy()
    expect[3:6] = rows[50:53]
    assert np.array_equal(h["z"][:], expect)
    for i in range(60):                                              # outgrows the slack of the root header: it moves once or twice
      h.create_dataset(f"extra/long_dataset_name_{i:03d}", data=np.arange(i + 1))
      h[f"s{i}"] = f"text {i}"
  with hdf5_lite.File(path, "r") as h:
    assert len(h.keys()) == 63 and len(h["extra"]) == 60
    assert np.array_equal(h["extra/long_dataset_name_041"][:], np.arange(42)) and h["s17"][()] == "text 17"
    assert np.array_equal(h["a"][200:300], rows[200:300])
    with pytest.raises(OSError):
      h["a"][0] = 1.0


def test_session_keeps_one_handle_and_every_trajectory_leaves_a_complete_file(tmp_path):
  path = str(tmp_path / "s.h5")
  rows = _rows(40, seed=2)
  utils.dump_dict(path, "featurizer_parms", {"cutoff": 5.0, "outfile": path})
  opened = []
  real_open = utils._open
  utils._open = lambda p, m: opened.append((p, m)) or real_open(p, m)
  try:
    with utils.hdf_session():
      for t in range(4):
        with utils.batched_output():
          for i in range(10):
            utils.append_hdf_data(path, "k", rows[t * 10 + i][None], dtype=np.float32, maxshape=(None, 2, 2, 2), chunks=(1, 2, 2, 2))
        # what is on disk now is a whole file with this trajectory's rows (read through an independent handle)
        with utils.open_hdf(path, "r") as h:
          assert h["k"].shape[0] == 10 * (t + 1) and np.array_equal(h["k"][:], rows[:10 * (t + 1)])
  finally:
    utils._open = real_open
  assert [m for _, m in opened].count("a") == 1
  for i in range(3):                                                 # outside a session: open / append / close per call, as the reference does
    utils.append_hdf_data(path, "k", rows[i][None], dtype=np.float32, maxshape=(None, 2, 2, 2), chunks=(1, 2, 2, 2))
  with utils.open_hdf(path, "r") as h:
    assert h["k"].shape[0] == 43 and h["featurizer_parms/cutoff"][()] == 5.0 and h["featurizer_parms/outfile"][()] == path


def test_update_hdf_data_grows_and_overwrites(tmp_path):
  """utils.update_hdf_data (/root/reference/nearl/utils.py:524-559): created from the first slice, grown when a slice ends
  beyond the dataset, overwritten in place otherwise -- by path and on an open file."""
  path = str(tmp_path / "u.h5")
  for epoch in range(5):
    utils.update_hdf_data(path, "loss", np.array([epoch * 1.5], dtype=np.float64), np.s_[epoch:epoch + 1], dtype=np.float64, maxshape=(None,))
  utils.update_hdf_data(path, "loss", np.array([-1.0, -2.0]), np.s_[1:3])
  with utils.open_hdf(path, "a") as h:
    utils.update_hdf_data(h, "grid", np.ones((2, 3, 3), dtype=np.float32), np.s_[0:2], dtype=np.float32, maxshape=(None, 3, 3))
    utils.update_hdf_data(h, "grid", np.full((2, 3, 3), 7, dtype=np.float32), np.s_[3:5])
  with utils.open_hdf(path, "r") as h:
    assert np.array_equal(h["loss"][:], [0.0, -1.0, -2.0, 4.5, 6.0]) and h["loss"].dtype == np.float64
    assert h["grid"].shape == (5, 3, 3) and np.array_equal(h["grid"][0:2], np.ones((2, 3, 3))) and not h["grid"][2].any()
    assert np.array_equal(h["grid"][3:5], np.full((2, 3, 3), 7))


@pytest.mark.skipif(utils.hdf_backend() != "hdf5_lite", reason="h5py reads its own format")
def test_classic_layout_written_by_libhdf5():
  """A MATLAB 7.4 '-v7.3' file (libhdf5 1.6: 512-byte user block, version-0 superblock, symbol-table root group, local
  heap, version-1 object headers, contiguous float64, a fixed-length string attribute), shipped with scipy's test data,
  against the same variable in the MAT-5 twin of the file, read by scipy."""
  import scipy.io
  data = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data")
  hits = glob.glob(os.path.join(data, "testhdf5_7.4_GLNX86.mat"))
  if not hits:
    pytest.skip("scipy's test data is not installed")
  with hdf5_lite.File(hits[0], "r") as h:
    assert h._base == 512 and not h._own_format
    assert list(h.keys()) == ["testdouble"]
    d = h["testdouble"]
    assert d.dtype == np.dtype("<f8") and d.chunks is None
    assert d.attrs["MATLAB_class"] == b"double"
    got = d[:]
    assert np.array_equal(d[2:5], got[2:5]) and d[8, 0] == got[8, 0]
    # the WRITER's datatype message for float64 is, byte for byte, the one libhdf5 put into this file (the version-1
    # object header pads the message to a multiple of eight)
    messages, _ = hdf5_lite._Reader(h)._messages(d._addr)
    body = next(b for t, b, _ in messages if t == hdf5_lite.MSG_DATATYPE)
    mine = hdf5_lite._encode_dtype("<f8")
    assert bytes(body[:len(mine)]) == mine and not any(body[len(mine):])
  twin = scipy.io.loadmat(os.path.join(data, "testdouble_7.4_GLNX86.mat"))["testdouble"]
  assert np.array_equal(got.reshape(-1), twin.reshape(-1))            # MATLAB stores column-major: (9, 1) here, (1, 9) there
  with pytest.raises(OSError):
    hdf5_lite.File(hits[0], "a")


def test_classic_layout_chunked_read_path(tmp_path):
  """The reader's N-D chunk grid, shuffle + deflate + Fletcher-32 pipeline and version-1 messages, on a classic-layout
  file assembled here byte by byte (no libhdf5 on this image writes one: self-consistency only, the structures are those
  of the specification sections cited in hdf5_lite)."""
  import struct
  import zlib
  rng = np.random.default_rng(3)
  data = rng.integers(-1000, 1000, size=(5, 7, 3)).astype(">i4")           # big-endian on purpose
  chunks = (2, 4, 3)
  blob = bytearray(b"\x00" * 2048)

  def put(raw, align=8):
    addr = (len(blob) + align - 1) // align * align
    blob.extend(b"\x00" * (addr - len(blob)) + raw)
    return addr

  # chunks: shuffle, deflate, fletcher32 (checksum bytes are not verified by the reader, only stripped)
  entries = []
  for i in range(0, 5, 2):
    for j in range(0, 7, 4):
      c = np.zeros(chunks, dtype=">i4")
      part = data[i:i + 2, j:j + 4]
      c[:part.shape[0], :part.shape[1]] = part
      raw = c.tobytes()
      sh = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 4).T.tobytes()
      z = zlib.compress(sh) + b"\x00\x00\x00\x00"
      entries.append(((i, j, 0, 0), put(z), len(z)))
  node = b"TREE" + bytes([1, 0]) + struct.pack("<HQQ", len(entries), hdf5_lite.UNDEF, hdf5_lite.UNDEF)
  for offs, addr, size in entries:
    node += struct.pack("<II", size, 0) + struct.pack("<4Q", *offs) + struct.pack("<Q", addr)
  node += struct.pack("<II", 0, 0) + struct.pack("<4Q", 6, 8, 3, 4)
  btree = put(node)

  def msg(mtype, body):
    body += b"\x00" * (-len(body) % 8)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body

  space = bytes([1, 3, 0, 0, 0, 0, 0, 0]) + struct.pack("<3Q", 5, 7, 3)
  dtype = bytes([0x10, 0x09, 0x00, 0x00]) + struct.pack("<I", 4) + struct.pack("<HH", 0, 32)
  layout = bytes([1, 4, 2, 0, 0, 0, 0, 0]) + struct.pack("<Q", btree) + struct.pack("<4I", 2, 4, 3, 4)

  def filt(fid, name, cdata):
    name += b"\x00" * (8 - len(name) % 8)
    out = struct.pack("<HHHH", fid, len(name), 1, len(cdata)) + name + b"".join(struct.pack("<I", c) for c in cdata)
    return out + (b"\x00" * 4 if len(cdata) % 2 else b"")

  pipeline = bytes([1, 3, 0, 0, 0, 0, 0, 0]) + filt(2, b"shuffle", (4,)) + filt(1, b"deflate", (6,)) + filt(3, b"fletcher32", ())
  body2 = msg(0x0B, pipeline) + msg(0x08, layout)
  cont = put(body2)
  body1 = msg(0x01, space) + msg(0x03, dtype) + msg(0x10, struct.pack("<QQ", cont, len(body2)))
  dset = put(struct.pack("<BBHII4x", 1, 0, 4, 1, len(body1)) + body1)
  # root group: local heap with the name, one symbol-table node, a one-entry group B-tree
  names = b"\x00" * 8 + b"values\x00\x00"
  seg = put(names)
  heap = put(b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(names), hdf5_lite.UNDEF, seg))
  snod = put(b"SNOD" + bytes([1, 0]) + struct.pack("<H", 1) + struct.pack("<QQII16x", 8, dset, 0, 0))
  gtree = put(b"TREE" + bytes([0, 0]) + struct.pack("<HQQ", 1, hdf5_lite.UNDEF, hdf5_lite.UNDEF) + struct.pack("<QQQ", 0, snod, 8))
  rbody = msg(0x11, struct.pack("<QQ", gtree, heap))
  root = put(struct.pack("<BBHII4x", 1, 0, 1, 1, len(rbody)) + rbody)
  sb = hdf5_lite.SIGNATURE + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0)
  sb += struct.pack("<QQQQ", 0, hdf5_lite.UNDEF, len(blob), hdf5_lite.UNDEF) + struct.pack("<QQII16x", 0, root, 0, 0)
  blob[0:len(sb)] = sb
  path = str(tmp_path / "classic.h5")
  with open(path, "wb") as fh:
    fh.write(bytes(blob))
  with hdf5_lite.File(path, "r") as h:
    d = h["values"]
    assert d.shape == (5, 7, 3) and d.chunks == chunks and d.dtype == np.dtype(">i4") and d.compression == "gzip"
    assert np.array_equal(d[:], data) and np.array_equal(d[1:4], data[1:4]) and np.array_equal(d[4], data[4])
    assert np.array_equal(d[0:3, 2:6, 1], data[0:3, 2:6, 1])


def test_classic_layout_many_groups_are_read_lazily(tmp_path):
  """MISATO's MD.hdf5 is one group per PDB code (~17 000 in the root group).  A classic-layout file with 40 groups under a
  two-level group B-tree (5 symbol-table nodes), each holding a contiguous dataset: opening the file reads the names only,
  looking one group up reads that group only (self-assembled file: structures per the specification)."""
  import struct
  U = hdf5_lite.UNDEF
  blob = bytearray(b"\x00" * 1024)

  def put(raw, align=8):
    addr = (len(blob) + align - 1) // align * align
    blob.extend(b"\x00" * (addr - len(blob)) + raw)
    return addr

  def msg(mtype, body):
    body += b"\x00" * (-len(body) % 8)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body

  def header(messages):
    body = b"".join(messages)
    return put(struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body)

  def group(children):
    """children: sorted [(name, object header address)] -> object header address of a group holding them."""
    names = b"\x00" * 8
    offs = []
    for name, _ in children:
      offs.append(len(names))
      names += name.encode() + b"\x00"
      names += b"\x00" * (-len(names) % 8)
    seg = put(names)
    heap = put(b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(names), U, seg))
    snods = []
    for i in range(0, len(children), 8):
      part = list(zip(offs[i:i + 8], children[i:i + 8]))
      body = b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(part))
      for off, (_, addr) in part:
        body += struct.pack("<QQII16x", off, addr, 0, 0)
      snods.append((put(body), part[0][0], part[-1][0]))

    def tree(level, kids):            # kids: [(address, first key, last key)]
      body = b"TREE" + bytes([0, level]) + struct.pack("<HQQ", len(kids), U, U) + struct.pack("<Q", 0)
      for addr, _, last in kids:
        body += struct.pack("<QQ", addr, last)
      return put(body), kids[0][1], kids[-1][2]

    leaves = [tree(0, snods[i:i + 3]) for i in range(0, len(snods), 3)]
    root = leaves[0] if len(leaves) == 1 else tree(1, leaves)
    return header([msg(0x11, struct.pack("<QQ", root[0], heap))])

  def dataset(arr):
    raw = put(arr.tobytes())
    space = bytes([1, arr.ndim, 0, 0, 0, 0, 0, 0]) + struct.pack("<%dQ" % arr.ndim, *arr.shape)
    dtype = bytes([0x11, 0x20, 31, 0]) + struct.pack("<I", 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    layout = bytes([3, 1]) + struct.pack("<QQ", raw, arr.nbytes)
    return header([msg(0x01, space), msg(0x03, dtype), msg(0x08, layout)])

  rng = np.random.default_rng(4)
  codes = sorted(f"{i:03d}X" for i in range(40))
  data = {c: rng.random((3, 5, 3)).astype(np.float32) for c in codes}
  members = [(c, group([("trajectory_coordinates", dataset(data[c]))])) for c in codes]
  root = group(members)
  sb = hdf5_lite.SIGNATURE + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0)
  sb += struct.pack("<QQQQ", 0, U, len(blob), U) + struct.pack("<QQII16x", 0, root, 0, 0)
  blob[0:len(sb)] = sb
  path = str(tmp_path / "many.h5")
  with open(path, "wb") as fh:
    fh.write(bytes(blob))
  with hdf5_lite.File(path, "r") as h:
    at_open = h.stats["bytes_read"]
    assert sorted(h.keys()) == codes and len(h) == 40 and "017X" in h and "999X" not in h
    assert h.stats["bytes_read"] == at_open < 4096                       # names only: no member header was read
    got = h["017X/trajectory_coordinates"][:]
    assert np.array_equal(got, data["017X"])
    one = h.stats["bytes_read"] - at_open
    assert one < 1200                                                    # that group, its dataset header, its 180 bytes
    assert np.array_equal(h["031X"]["trajectory_coordinates"][1:3], data["031X"][1:3])
    assert sum(1 for v in dict.values(h._members) if isinstance(v, hdf5_lite._LazyRef)) == 38
    assert all(np.array_equal(g["trajectory_coordinates"][:], data[c]) for c, g in h.items())


def test_compressed_chunks_deflated_side_by_side_give_the_same_file(tmp_path, monkeypatch):
  """A spill of many pending chunks of a gzip dataset deflates them on a few threads: the file is the one a single
  thread writes, byte for byte, and reads back."""
  rows = _rows(300, shape=(8, 8, 8), seed=3) * (_rows(300, shape=(8, 8, 8), seed=4) > 0.7)

  def write(path):
    with hdf5_lite.File(path, "w") as h:
      for start in range(0, 300, 75):
        utils._append_now(h, "gz", rows[start:start + 75], np.float32, (None, 8, 8, 8), chunks=(1, 8, 8, 8), compression="gzip", compression_opts=4)
  write(str(tmp_path / "threads.h5"))

  class Serial:
    def map(self, fn, items):
      return map(fn, items)
  monkeypatch.setattr(hdf5_lite, "_deflate_pool", lambda: Serial())
  write(str(tmp_path / "serial.h5"))
  assert open(str(tmp_path / "threads.h5"), "rb").read() == open(str(tmp_path / "serial.h5"), "rb").read()
  with hdf5_lite.File(str(tmp_path / "threads.h5"), "r") as h:
    assert np.array_equal(h["gz"][:], rows)
    assert os.path.getsize(str(tmp_path / "threads.h5")) < rows.nbytes // 2
