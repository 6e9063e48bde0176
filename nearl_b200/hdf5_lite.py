"""A small, dependency-free HDF5 writer/reader for the output layout of the hot path.

The image this package is built on has neither h5py nor libhdf5, yet the path's product is
an HDF5 file (/root/reference/nearl/utils.py:491-570, /root/reference/nearl/features.py:510-534).
This module writes real HDF5 (version-2 superblock, version-2 object headers with compact
link storage, version-3 layout messages, version-1 chunk B-trees, a global heap for
variable-length strings) for exactly the objects that layout needs:

  * groups,
  * scalar and 1-D datasets of float64 / int64 / float32 / int32 / variable-length UTF-8
    strings (``featurizer_parms/*``, contiguous layout),
  * N-D float32 / int / float datasets extensible along axis 0 with chunks that span the
    trailing axes (``chunks=(c0, *shape[1:])``, ``maxshape=(None, ...)``), optionally gzip
    compressed (one feature dataset per ``outkey``).

and reads the same subset back (enough for tests and for appending to a file it wrote).
The API is the slice of h5py's that nearl uses: ``File(path, mode)`` as a context manager,
``in`` / ``[]`` / ``del`` / ``create_group`` / ``create_dataset`` / item assignment, and on
datasets ``shape`` / ``dtype`` / ``maxshape`` / ``chunks`` / ``resize`` / slicing on axis 0.

Appending never moves chunk data: new chunks go to the end of the file and the (small)
metadata -- object headers, B-tree nodes, heap -- is written afresh behind them, after which
the superblock is pointed at the new root header.  Format reference: "HDF5 File Format
Specification Version 3.0" (sections II.A superblock, III.A.1 v1 B-trees, III.E global heap,
IV.A.1.b / IV.A.2 object headers and messages).  Compatibility with libhdf5 could not be
exercised on this image (no HDF5 library available); the structures follow the
specification field by field and are checked by the module's own independent reader.
"""
import os
import struct
import zlib

import numpy as np

__all__ = ["File", "Group", "Dataset"]

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
UNLIMITED = 0xFFFFFFFFFFFFFFFF
BTREE_K = 32      # default "indexed storage internal node K": at most 2K = 64 children per chunk B-tree node

MSG_DATASPACE, MSG_LINKINFO, MSG_DATATYPE, MSG_FILLVALUE, MSG_LINK = 0x01, 0x02, 0x03, 0x05, 0x06
MSG_LAYOUT, MSG_GROUPINFO, MSG_FILTERS = 0x08, 0x0A, 0x0B


# ---- Jenkins lookup3 (hashlittle), the metadata checksum of HDF5 -------------------------------
def _rot(x, k):
  return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF


def lookup3(data, initval=0):
  M = 0xFFFFFFFF
  n = len(data)
  a = b = c = (0xDEADBEEF + n + initval) & M
  pos = 0
  while n > 12:
    a = (a + int.from_bytes(data[pos:pos + 4], "little")) & M
    b = (b + int.from_bytes(data[pos + 4:pos + 8], "little")) & M
    c = (c + int.from_bytes(data[pos + 8:pos + 12], "little")) & M
    a = (a - c) & M; a ^= _rot(c, 4); c = (c + b) & M
    b = (b - a) & M; b ^= _rot(a, 6); a = (a + c) & M
    c = (c - b) & M; c ^= _rot(b, 8); b = (b + a) & M
    a = (a - c) & M; a ^= _rot(c, 16); c = (c + b) & M
    b = (b - a) & M; b ^= _rot(a, 19); a = (a + c) & M
    c = (c - b) & M; c ^= _rot(b, 4); b = (b + a) & M
    pos += 12
    n -= 12
  if n == 0:
    return c
  tail = bytes(data[pos:pos + n]) + b"\x00" * (12 - n)
  a = (a + int.from_bytes(tail[0:4], "little")) & M
  b = (b + int.from_bytes(tail[4:8], "little")) & M
  c = (c + int.from_bytes(tail[8:12], "little")) & M
  c ^= b; c = (c - _rot(b, 14)) & M
  a ^= c; a = (a - _rot(c, 11)) & M
  b ^= a; b = (b - _rot(a, 25)) & M
  c ^= b; c = (c - _rot(b, 16)) & M
  a ^= c; a = (a - _rot(c, 4)) & M
  b ^= a; b = (b - _rot(a, 14)) & M
  c ^= b; c = (c - _rot(b, 24)) & M
  return c


# ---- datatypes ------------------------------------------------------------------------------------
class _VlenStr:
  """Marker dtype for variable-length UTF-8 strings."""
  def __repr__(self):
    return "vlen-str"


VLEN_STR = _VlenStr()


def _encode_dtype(dt):
  if dt is VLEN_STR:
    # class 9 (variable-length), version 1; type = string (1), padding = null terminated (0), charset = UTF-8 (1)
    # parent type as libhdf5 builds it for H5T_VARIABLE strings: unsigned 1-byte integer
    base = bytes([0x10, 0x00, 0x00, 0x00]) + struct.pack("<I", 1) + struct.pack("<HH", 0, 8)
    return bytes([0x19, 0x01, 0x01, 0x00]) + struct.pack("<I", 16) + base
  dt = np.dtype(dt)
  if dt.kind == "f":
    size = dt.itemsize
    exp_bits, mant_bits, bias = {4: (8, 23, 127), 8: (11, 52, 1023)}[size]
    bits0 = 0x20                                   # little endian, implied most-significant mantissa bit
    head = bytes([0x11, bits0, size * 8 - 1, 0x00]) + struct.pack("<I", size)
    props = struct.pack("<HHBBBBI", 0, size * 8, mant_bits, exp_bits, 0, mant_bits, bias)
    return head + props
  if dt.kind in "iu":
    size = dt.itemsize
    bits0 = 0x08 if dt.kind == "i" else 0x00       # little endian, two's complement signed
    return bytes([0x10, bits0, 0x00, 0x00]) + struct.pack("<I", size) + struct.pack("<HH", 0, size * 8)
  raise TypeError(f"hdf5_lite: unsupported dtype {dt}")


def _decode_dtype(buf):
  cls = buf[0] & 0x0F
  size = struct.unpack_from("<I", buf, 4)[0]
  if cls == 1:
    return np.dtype("<f%d" % size)
  if cls == 0:
    return np.dtype(("<i%d" if buf[1] & 0x08 else "<u%d") % size)
  if cls == 9 and (buf[1] & 0x0F) == 1:
    return VLEN_STR
  raise TypeError(f"hdf5_lite: unsupported datatype class {cls}")


# ---- object model -----------------------------------------------------------------------------------
class Dataset:
  def __init__(self, file, shape, dtype, maxshape=None, chunks=None, compression=None, compression_opts=None):
    self._file = file
    self.shape = tuple(int(s) for s in shape)
    self.dtype = dtype if dtype is VLEN_STR else np.dtype(dtype)
    self.maxshape = tuple(maxshape) if maxshape is not None else self.shape
    self.compression = compression
    self.compression_opts = 4 if (compression == "gzip" and compression_opts is None) else compression_opts
    extensible = any(m is None for m in self.maxshape) or self.maxshape != self.shape
    if chunks is True or (chunks is None and (extensible or compression)):
      chunks = (1,) + self.shape[1:] if len(self.shape) > 0 else None
    self.chunks = tuple(int(c) for c in chunks) if chunks else None
    if self.chunks is not None:
      if len(self.chunks) != len(self.shape) or self.chunks[1:] != self.shape[1:]:
        raise ValueError("hdf5_lite: chunks must span the trailing axes: (c0, *shape[1:])")
      self._disk_chunks = {}       # chunk index along axis 0 -> (address, stored size)
      self._mem_chunks = {}        # chunk index -> ndarray [c0, *trailing] not yet on disk
      self._value = None
    else:
      self._value = None if self.dtype is VLEN_STR else np.zeros(self.shape, dtype=self.dtype)
      self._contig_addr = None

  # -- h5py-like surface
  def __len__(self):
    return self.shape[0]

  @property
  def ndim(self):
    return len(self.shape)

  def resize(self, size, axis=None):
    if self.chunks is None:
      raise TypeError("Only chunked datasets can be resized")
    if axis is not None:
      new = list(self.shape); new[axis] = int(size); size = new
    size = tuple(int(s) for s in size)
    if size[1:] != self.shape[1:]:
      raise ValueError("hdf5_lite: only axis 0 can be resized")
    if self.maxshape[0] is not None and size[0] > self.maxshape[0]:
      raise ValueError("resize beyond maxshape")
    self.shape = size
    self._file._dirty = True

  def _row_range(self, key):
    """Split an index into (row start, row stop, row-is-scalar, index for the trailing axes)."""
    rest = ()
    if isinstance(key, tuple):
      if any(k is Ellipsis for k in key[1:]):
        raise IndexError("hdf5_lite: Ellipsis is only supported as the whole index")
      rest = tuple(key[1:])
      key = key[0] if key else slice(None)
    if key is Ellipsis:
      key = slice(None)
    if isinstance(key, (int, np.integer)):
      k = int(key) + (self.shape[0] if key < 0 else 0)
      if not 0 <= k < self.shape[0]:
        raise IndexError("index out of range")
      return k, k + 1, True, rest
    start, stop, step = key.indices(self.shape[0])
    if step != 1:
      raise IndexError("hdf5_lite: step slicing is not supported")
    return start, max(stop, start), False, rest

  def _load_chunk(self, ci):
    if ci in self._mem_chunks:
      return self._mem_chunks[ci]
    if ci in self._disk_chunks:
      addr, size = self._disk_chunks[ci]
      raw = self._file._read(addr, size)
      if self.compression == "gzip":
        raw = zlib.decompress(raw)
      return np.frombuffer(raw, dtype=self.dtype).reshape(self.chunks).copy()
    return np.zeros(self.chunks, dtype=self.dtype)

  def __getitem__(self, key):
    if self.ndim == 0:
      return self._value if self.dtype is VLEN_STR else self._value[()]
    if self.chunks is None:
      return self._value[key]
    r0, r1, scalar, rest = self._row_range(key)
    out = np.zeros((r1 - r0,) + self.shape[1:], dtype=self.dtype)
    c0 = self.chunks[0]
    for ci in range(r0 // c0, (r1 + c0 - 1) // c0 if r1 > r0 else r0 // c0):
      lo, hi = max(r0, ci * c0), min(r1, (ci + 1) * c0)
      out[lo - r0:hi - r0] = self._load_chunk(ci)[lo - ci * c0:hi - ci * c0]
    if scalar:
      return out[0][rest] if rest else out[0]
    return out[(slice(None),) + rest] if rest else out

  def __setitem__(self, key, value):
    self._file._dirty = True
    if self.ndim == 0 or self.chunks is None:
      if self.dtype is VLEN_STR:
        self._value = value
      else:
        self._value[key] = value
      return
    r0, r1, _, rest = self._row_range(key)
    if any(k != slice(None) for k in rest):
      raise IndexError("hdf5_lite: assignment supports slicing on axis 0 only")
    value = np.broadcast_to(np.asarray(value, dtype=self.dtype), (r1 - r0,) + self.shape[1:])
    c0 = self.chunks[0]
    for ci in range(r0 // c0, (r1 + c0 - 1) // c0 if r1 > r0 else r0 // c0):
      lo, hi = max(r0, ci * c0), min(r1, (ci + 1) * c0)
      if hi - lo == c0:
        chunk = np.ascontiguousarray(value[lo - r0:hi - r0])
      else:
        chunk = self._load_chunk(ci)
        chunk[lo - ci * c0:hi - ci * c0] = value[lo - r0:hi - r0]
      self._mem_chunks[ci] = chunk
      self._disk_chunks.pop(ci, None)
    if len(self._mem_chunks) * int(np.prod(self.chunks)) * self.dtype.itemsize > (256 << 20):
      self._file._spill(self)


class Group:
  def __init__(self, file):
    self._file = file
    self._members = {}

  def _resolve(self, path, create=False):
    node = self
    parts = [p for p in path.split("/") if p]
    for p in parts[:-1]:
      if p not in node._members:
        if not create:
          raise KeyError(path)
        node._members[p] = Group(self._file)
      node = node._members[p]
    return node, (parts[-1] if parts else "")

  def __contains__(self, path):
    try:
      node, leaf = self._resolve(path)
    except KeyError:
      return False
    return leaf in node._members

  def __getitem__(self, path):
    node, leaf = self._resolve(path)
    return node._members[leaf]

  def __delitem__(self, path):
    node, leaf = self._resolve(path)
    del node._members[leaf]
    self._file._dirty = True

  def keys(self):
    return self._members.keys()

  def items(self):
    return self._members.items()

  def __iter__(self):
    return iter(self._members)

  def create_group(self, path):
    node, leaf = self._resolve(path, create=True)
    if leaf in node._members:
      raise ValueError(f"Unable to create group (name already exists): {path}")
    node._members[leaf] = Group(self._file)
    self._file._dirty = True
    return node._members[leaf]

  def create_dataset(self, path, shape=None, dtype=None, data=None, maxshape=None, chunks=None, compression=None,
                     compression_opts=None, **_ignored):
    node, leaf = self._resolve(path, create=True)
    if leaf in node._members:
      raise ValueError(f"Unable to create dataset (name already exists): {path}")
    if data is not None:
      data = np.asarray(data)
      shape = data.shape if shape is None else shape
      dtype = data.dtype if dtype is None else dtype
    if isinstance(shape, (int, np.integer)):
      shape = (int(shape),)
    dset = Dataset(self._file, shape, dtype, maxshape=maxshape, chunks=chunks, compression=compression,
                   compression_opts=compression_opts)
    node._members[leaf] = dset
    self._file._dirty = True
    if data is not None:
      dset[...] = data
    return dset

  def __setitem__(self, path, value):
    """h5py's implicit dataset creation: float -> f64, int -> i64, str -> vlen UTF-8, sequences -> arrays."""
    node, leaf = self._resolve(path, create=True)
    if isinstance(value, (str, bytes)):
      dset = Dataset(self._file, (), VLEN_STR)
      dset._value = value.decode("utf-8") if isinstance(value, bytes) else value
    else:
      arr = np.asarray(value)
      if arr.dtype == bool:
        arr = arr.astype(np.int8)
      if arr.dtype.kind in "US":
        raise TypeError("hdf5_lite: string arrays are not supported")
      if arr.dtype.kind == "O":
        raise TypeError(f"hdf5_lite: cannot store object value {value!r}")
      dset = Dataset(self._file, arr.shape, arr.dtype)
      dset._value = arr.copy()
    node._members[leaf] = dset
    self._file._dirty = True


class File(Group):
  def __init__(self, path, mode="r"):
    Group.__init__(self, self)
    self.filename = path
    self.mode = mode
    self._dirty = False
    self._fh = None
    self._eof = 0
    exists = os.path.exists(path) and os.path.getsize(path) > 0
    if mode == "r":
      if not exists:
        raise FileNotFoundError(path)
      self._fh = open(path, "rb")
      _Reader(self).load()
    elif mode in ("a", "r+"):
      if exists:
        self._fh = open(path, "r+b")
        _Reader(self).load()
      elif mode == "r+":
        raise FileNotFoundError(path)
      else:
        self._fh = open(path, "w+b")
        self._dirty = True
    elif mode in ("w", "w-", "x"):
      if mode != "w" and exists:
        raise FileExistsError(path)
      self._fh = open(path, "w+b")
      self._dirty = True
    else:
      raise ValueError(f"invalid mode {mode!r}")

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    self.close()
    return False

  def _read(self, addr, size):
    self._fh.seek(addr)
    return self._fh.read(size)

  def _append(self, blob, align=8):
    """Write `blob` at the (aligned) end of file, return its address."""
    addr = (max(self._eof, 48) + align - 1) // align * align
    self._fh.seek(addr)
    self._fh.write(blob)
    self._eof = addr + len(blob)
    return addr

  def _spill(self, dset):
    """Move a dataset's pending chunks to disk (bounded memory while appending many rows)."""
    for ci in sorted(dset._mem_chunks):
      raw = np.ascontiguousarray(dset._mem_chunks[ci]).tobytes()
      if dset.compression == "gzip":
        raw = zlib.compress(raw, int(dset.compression_opts))
      dset._disk_chunks[ci] = (self._append(raw), len(raw))
    dset._mem_chunks.clear()

  def flush(self):
    if self._fh is None or self.mode == "r" or not self._dirty:
      return
    _Writer(self).write()
    self._dirty = False

  def close(self):
    if self._fh is None:
      return
    self.flush()
    self._fh.close()
    self._fh = None


# ---- writer -------------------------------------------------------------------------------------------
def _message(mtype, body, flags=0):
  return struct.pack("<BHB", mtype, len(body), flags) + body


def _object_header(messages):
  body = b"".join(messages)
  # flags: chunk-0 size stored in 4 bytes (bits 0-1 = 2); no times, no attribute phase change
  head = b"OHDR" + bytes([2, 0x02]) + struct.pack("<I", len(body))
  blob = head + body
  return blob + struct.pack("<I", lookup3(blob))


class _Writer:
  def __init__(self, file):
    self.f = file
    self.heap_objects = []     # payloads of the global heap collection of this session

  def write(self):
    f = self.f
    # 1. pending chunk data first (it never moves afterwards)
    for dset in self._datasets(f):
      if dset.chunks is not None:
        f._spill(dset)
    # 2. global heap for variable-length strings: its address must be known before the
    #    datasets that point into it are serialised
    strings = [d for d in self._datasets(f) if d.dtype is VLEN_STR]
    heap_addr = None
    if strings:
      payloads = [str(d._value).encode("utf-8") for d in strings]
      heap_addr = f._append(self._global_heap(payloads))
      for idx, (d, p) in enumerate(zip(strings, payloads)):
        d._heap_ref = struct.pack("<IQI", len(p), heap_addr, idx + 1)
    # 3. metadata bottom-up
    root_addr = self._write_group(f)
    # 4. superblock
    eof = f._eof
    sb = SIGNATURE + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, UNDEF, eof, root_addr)
    sb += struct.pack("<I", lookup3(sb))
    f._fh.seek(0)
    f._fh.write(sb)
    f._fh.flush()

  def _datasets(self, group):
    for obj in group._members.values():
      if isinstance(obj, Dataset):
        yield obj
      else:
        yield from self._datasets(obj)

  @staticmethod
  def _global_heap(payloads):
    body = b""
    for idx, p in enumerate(payloads):
      pad = (-len(p)) % 8
      body += struct.pack("<HHIQ", idx + 1, 1, 0, len(p)) + p + b"\x00" * pad
    size = max(4096, 16 + len(body) + 16)
    free = size - 16 - len(body)
    body += struct.pack("<HHIQ", 0, 0, 0, free) + b"\x00" * (free - 16)   # object 0: the free space
    return b"GCOL" + bytes([1, 0, 0, 0]) + struct.pack("<Q", size) + body

  def _write_group(self, group):
    links = []
    for name, obj in group._members.items():
      addr = self._write_dataset(obj) if isinstance(obj, Dataset) else self._write_group(obj)
      nb = name.encode("utf-8")
      # link message v1: flags = 0x10|size-of-length field (bit 4: character set present -> UTF-8), hard link
      size_code = 0 if len(nb) < 256 else 1
      body = bytes([1, 0x10 | size_code, 1]) + (struct.pack("<B", len(nb)) if size_code == 0 else struct.pack("<H", len(nb)))
      body += nb + struct.pack("<Q", addr)
      links.append(_message(MSG_LINK, body))
    link_info = _message(MSG_LINKINFO, bytes([0, 0]) + struct.pack("<QQ", UNDEF, UNDEF))
    group_info = _message(MSG_GROUPINFO, bytes([0, 0]))
    return self.f._append(_object_header([link_info, group_info] + links))

  def _write_dataset(self, d):
    f = self.f
    rank = len(d.shape)
    # dataspace v2
    if rank == 0:
      space = bytes([2, 0, 0, 0])
    else:
      has_max = d.maxshape != d.shape
      space = bytes([2, rank, 1 if has_max else 0, 1]) + b"".join(struct.pack("<Q", s) for s in d.shape)
      if has_max:
        space += b"".join(struct.pack("<Q", UNLIMITED if m is None else m) for m in d.maxshape)
    msgs = [_message(MSG_DATASPACE, space), _message(MSG_DATATYPE, _encode_dtype(d.dtype), flags=1)]
    if d.chunks is not None:
      msgs.append(_message(MSG_FILLVALUE, bytes([3, 0x0B])))          # incremental allocation, fill if set, default fill
      if d.compression == "gzip":
        filt = struct.pack("<HHHI", 1, 1, 1, int(d.compression_opts))   # deflate, optional, one client value
        msgs.append(_message(MSG_FILTERS, bytes([2, 1]) + filt))
      btree = self._write_chunk_btree(d)
      itemsize = d.dtype.itemsize
      layout = bytes([3, 2, rank + 1]) + struct.pack("<Q", btree)
      layout += b"".join(struct.pack("<I", c) for c in d.chunks) + struct.pack("<I", itemsize)
    else:
      msgs.append(_message(MSG_FILLVALUE, bytes([3, 0x0A])))          # late allocation, fill if set
      if d.dtype is VLEN_STR:
        raw = d._heap_ref
      else:
        raw = np.ascontiguousarray(d._value, dtype=d.dtype).tobytes()
      addr = f._append(raw) if raw else UNDEF
      layout = bytes([3, 1]) + struct.pack("<QQ", addr, len(raw))
    msgs.append(_message(MSG_LAYOUT, layout))
    return f._append(_object_header(msgs))

  def _write_chunk_btree(self, d):
    """Static bulk-loaded v1 B-tree (node type 1) over the chunks along axis 0."""
    f = self.f
    rank = len(d.shape)
    c0 = d.chunks[0]
    n_chunks = (d.shape[0] + c0 - 1) // c0

    def key(size, row):
      return struct.pack("<II", size, 0) + struct.pack("<Q", row) + b"\x00" * (8 * rank)

    key_len = 8 + 8 * (rank + 1)
    node_len = 24 + (2 * BTREE_K + 1) * key_len + 2 * BTREE_K * 8
    entries = []     # (first row, stored size, child address)
    for ci in range(n_chunks):
      if ci in d._disk_chunks:
        addr, size = d._disk_chunks[ci]
        entries.append((ci * c0, size, addr))
    level = 0
    if not entries:
      blob = b"TREE" + bytes([1, 0]) + struct.pack("<HQQ", 0, UNDEF, UNDEF)
      return f._append(blob + b"\x00" * (node_len - len(blob)))
    end_row = (entries[-1][0] // c0 + 1) * c0
    while True:
      groups = [entries[i:i + 2 * BTREE_K] for i in range(0, len(entries), 2 * BTREE_K)]
      base = (max(f._eof, 48) + 7) // 8 * 8
      addrs = [base + i * node_len for i in range(len(groups))]
      parents = []
      for gi, grp in enumerate(groups):
        left = addrs[gi - 1] if gi > 0 else UNDEF
        right = addrs[gi + 1] if gi + 1 < len(groups) else UNDEF
        blob = b"TREE" + bytes([1, level]) + struct.pack("<HQQ", len(grp), left, right)
        for row, size, child in grp:
          blob += key(size, row) + struct.pack("<Q", child)
        last_row = groups[gi + 1][0][0] if gi + 1 < len(groups) else end_row
        blob += key(0, last_row)
        blob += b"\x00" * (node_len - len(blob))
        got = f._append(blob)
        assert got == addrs[gi]
        parents.append((grp[0][0], grp[0][1], got))
      if len(parents) == 1:
        return parents[0][2]
      entries, level = parents, level + 1


# ---- reader (independent walk over the structures above) -------------------------------------------------
class _Reader:
  def __init__(self, file):
    self.f = file
    self.heaps = {}

  def load(self):
    f = self.f
    head = f._read(0, 48)
    if head[:8] != SIGNATURE or head[8] != 2:
      raise OSError("hdf5_lite can only reopen files with a version-2 superblock (its own output)")
    if struct.unpack_from("<I", head, 44)[0] != lookup3(head[:44]):
      raise OSError("superblock checksum mismatch")
    _base, _ext, eof, root = struct.unpack_from("<QQQQ", head, 12)
    f._eof = eof
    self._load_group(f, root)

  def _messages(self, addr):
    f = self.f
    head = f._read(addr, 16)
    if head[:4] != b"OHDR" or head[4] != 2:
      raise OSError("unsupported object header")
    flags = head[5]
    pos = 6
    if flags & 0x20:
      pos += 16
    if flags & 0x10:
      pos += 4
    width = 1 << (flags & 3)
    size = int.from_bytes(head[pos:pos + width], "little")
    pos += width
    blob = f._read(addr, pos + size + 4)
    if struct.unpack_from("<I", blob, pos + size)[0] != lookup3(blob[:pos + size]):
      raise OSError("object header checksum mismatch")
    end = pos + size
    out = []
    while pos + 4 <= end:
      mtype, msize, _mflags = struct.unpack_from("<BHB", blob, pos)
      pos += 4
      if flags & 0x04:
        pos += 2
      out.append((mtype, blob[pos:pos + msize]))
      pos += msize
    return out

  def _load_group(self, group, addr):
    for mtype, body in self._messages(addr):
      if mtype != MSG_LINK:
        continue
      flags = body[1]
      pos = 2
      if flags & 0x08:
        if body[pos] != 0:
          raise OSError("only hard links are supported")
        pos += 1
      if flags & 0x04:
        pos += 8
      if flags & 0x10:
        pos += 1
      width = 1 << (flags & 3)
      nlen = int.from_bytes(body[pos:pos + width], "little")
      pos += width
      name = body[pos:pos + nlen].decode("utf-8")
      child = struct.unpack_from("<Q", body, pos + nlen)[0]
      kinds = {m for m, _ in self._messages(child)}
      if MSG_LAYOUT in kinds:
        group._members[name] = self._load_dataset(child)
      else:
        sub = Group(self.f)
        self._load_group(sub, child)
        group._members[name] = sub

  def _load_dataset(self, addr):
    f = self.f
    shape, maxshape, dtype, layout, gzip_level = (), None, None, None, None
    for mtype, body in self._messages(addr):
      if mtype == MSG_DATASPACE:
        rank, flags = body[1], body[2]
        shape = struct.unpack_from("<%dQ" % rank, body, 4) if rank else ()
        if flags & 1:
          mx = struct.unpack_from("<%dQ" % rank, body, 4 + 8 * rank)
          maxshape = tuple(None if m == UNLIMITED else m for m in mx)
      elif mtype == MSG_DATATYPE:
        dtype = _decode_dtype(body)
      elif mtype == MSG_LAYOUT:
        layout = body
      elif mtype == MSG_FILTERS:
        if body[0] != 2 or body[1] != 1 or struct.unpack_from("<H", body, 2)[0] != 1:
          raise OSError("only a single deflate filter is supported")
        gzip_level = struct.unpack_from("<I", body, 8)[0]
    if layout[0] != 3:
      raise OSError("unsupported layout message version")
    if layout[1] == 2:
      rank = layout[2] - 1
      btree = struct.unpack_from("<Q", layout, 3)[0]
      chunks = struct.unpack_from("<%dI" % rank, layout, 11)
      d = Dataset(f, shape, dtype, maxshape=maxshape, chunks=chunks, compression="gzip" if gzip_level is not None else None,
                  compression_opts=gzip_level)
      self._walk_btree(btree, d)
      return d
    if layout[1] != 1:
      raise OSError("unsupported layout class")
    caddr, csize = struct.unpack_from("<QQ", layout, 2)
    d = Dataset(f, shape, dtype, maxshape=maxshape)
    raw = f._read(caddr, csize) if caddr != UNDEF and csize else b""
    if dtype is VLEN_STR:
      length, haddr, index = struct.unpack("<IQI", raw)
      d._value = self._heap_object(haddr, index)[:length].decode("utf-8")
    else:
      d._value = np.frombuffer(raw, dtype=dtype).reshape(shape).copy()
    return d

  def _walk_btree(self, addr, d):
    f = self.f
    rank = len(d.shape)
    key_len = 8 + 8 * (rank + 1)
    head = f._read(addr, 24)
    if head[:4] != b"TREE" or head[4] != 1:
      raise OSError("bad chunk B-tree node")
    level, used = head[5], struct.unpack_from("<H", head, 6)[0]
    body = f._read(addr + 24, used * (key_len + 8) + key_len)
    for i in range(used):
      off = i * (key_len + 8)
      size = struct.unpack_from("<I", body, off)[0]
      row = struct.unpack_from("<Q", body, off + 8)[0]
      child = struct.unpack_from("<Q", body, off + key_len)[0]
      if level == 0:
        d._disk_chunks[row // d.chunks[0]] = (child, size)
      else:
        self._walk_btree(child, d)

  def _heap_object(self, addr, index):
    f = self.f
    if addr not in self.heaps:
      head = f._read(addr, 16)
      if head[:4] != b"GCOL":
        raise OSError("bad global heap collection")
      size = struct.unpack_from("<Q", head, 8)[0]
      blob = f._read(addr, size)
      objs, pos = {}, 16
      while pos + 16 <= size:
        idx, _ref, _res, osize = struct.unpack_from("<HHIQ", blob, pos)
        if idx == 0:
          break
        objs[idx] = blob[pos + 16:pos + 16 + osize]
        pos += 16 + (osize + 7) // 8 * 8
      self.heaps[addr] = objs
    return self.heaps[addr][index]
