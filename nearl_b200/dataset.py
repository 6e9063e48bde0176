"""Consumer of the featuriser's HDF5 output: samples -> (optionally normalised, augmented) torch
mini-batches, with the interface of /root/reference/nearl/io/dataset.py (`Dataset`, `readdata`,
`data_augment`, `split_array`).  SURVEY.md section 8(f) rank 4: the step AFTER the hot path.

Differences from the reference, all on purpose: files are read through `utils.open_hdf` (h5py when it
is importable, else the self-contained reader), a batch is read file by file in one pass instead of
one `open()` per (sample, channel) in a 24-process pool, and `device=` moves the batch to the GPU
before the augmentation so that flips / rotations / rolls run there.
"""
import os

import numpy as np

from . import config, printit, utils

__all__ = ["Dataset", "data_augment", "readdata", "split_array", "rand_flip_axis", "rand_translate"]

# flips of a (batch, channel, x, y, z) tensor / of a single (x, y, z) grid (dataset.py:15-37)
FLIP_DICT_BATCH_LEVEL_TORCH = {0: [], 1: [2], 2: [3], 3: [4], 4: [2, 3], 5: [2, 4], 6: [3, 4], 7: [2, 3, 4]}
FLIP_DICT_GRID_LEVEL = {0: [], 1: [0], 2: [1], 3: [2], 4: [0, 1], 5: [0, 2], 6: [1, 2], 7: [0, 1, 2]}


def rand_flip_axis(level):
  rand_flip = int(np.random.random() * 1000) % 8
  return FLIP_DICT_BATCH_LEVEL_TORCH[rand_flip] if level == "batch" else FLIP_DICT_GRID_LEVEL[rand_flip]


def rand_translate(factor=1):
  return np.random.randint(low=-1 * factor, high=1 * factor + 1, size=3)


def _normalise(ret_data, normalize):
  """Per-sample normalisation of readdata (dataset.py:72-87).  Only the strings "minmax" and "zscore" do
  anything: the `normalize=True` a Dataset passes by default leaves the data as it is, as in the reference."""
  if isinstance(normalize, str) and normalize == "minmax":
    min_val, max_val = np.min(ret_data), np.max(ret_data)
    if np.isclose(min_val, max_val):
      return np.zeros_like(ret_data)
    return (ret_data - min_val) / (max_val - min_val)
  if isinstance(normalize, str) and normalize == "zscore":
    mean_val, std_val = np.mean(ret_data), np.std(ret_data)
    if np.isclose(std_val, 0) or np.isnan(std_val) or np.isnan(mean_val):
      return np.zeros_like(ret_data)
    return (ret_data - mean_val) / std_val
  return ret_data


_READ_POOLS = {}


def _read_pool(workers):
  pool = _READ_POOLS.get(workers)
  if pool is None:
    from concurrent.futures import ThreadPoolExecutor
    pool = _READ_POOLS[workers] = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="nearl-dataset-read")
  return pool


def readdata(input_file, keyword, theslice, normalize=False):
  """One sample (or slice) of dataset `keyword` (dataset.py:51-88)."""
  with utils.open_hdf(input_file, "r") as h5file:
    ret_data = np.asarray(h5file[keyword][theslice])
  return _normalise(ret_data, normalize)


def split_array(input_array, batch_size):
  """Consecutive batches of `batch_size`, the remainder last (dataset.py:97-108)."""
  bin_nr = (len(input_array) + batch_size - 1) // batch_size
  if len(input_array) == batch_size * bin_nr:
    return np.array_split(input_array, bin_nr)
  final_batch_size = len(input_array) % batch_size
  if bin_nr - 1 == 0:
    return [input_array[-final_batch_size:]]
  return np.array_split(input_array[:-final_batch_size], bin_nr - 1) + [input_array[-final_batch_size:]]


def data_augment(batch_array, trans=0, add_noise=False):
  """Random flips, 90-degree rotations about the three axes, zero-filled integer translations and
  optional 10 % noise on a (batch, channel, x, y, z) tensor (dataset.py:111-180); the random draws come
  from numpy's global generator in the reference's order."""
  import torch
  flip_axes = rand_flip_axis("batch")
  for axis in flip_axes:
    batch_array = batch_array.flip(dims=(axis,))
  for axis in range(3):
    if np.random.choice([0, 1]):
      k = int(np.random.choice([1, 2, 3]))
      dims = [3, 4] if axis == 0 else ([2, 4] if axis == 1 else [2, 3])
      batch_array = torch.rot90(batch_array, k, dims=dims)
  if trans > 0:
    for i, shift in enumerate(rand_translate(factor=trans)):
      shift = int(shift)
      if shift == 0:
        continue
      batch_array = torch.roll(batch_array, shifts=shift, dims=(i + 2))
      index = [slice(None)] * 5
      index[i + 2] = slice(None, shift) if shift > 0 else slice(shift, None)
      batch_array[tuple(index)] = 0
  if config.verbose() or config.debug():
    printit(f"Flipped axes: {flip_axes}; ")
  if add_noise:
    batch_array = batch_array + torch.randn_like(batch_array) * torch.max(batch_array) * 0.1
  return batch_array


class Dataset:
  """Samples spread over several featuriser output files (dataset.py:183-393).  `feature_keys` become the
  channels of a sample; `label_key` (default "label") the target, 0 when the files carry none; the grid size
  comes from ``featurizer_parms/dimensions`` unless `grid_dim` is given."""

  def __init__(self, files, grid_dim=None, label_key="label", feature_keys=[], normalize=True):
    self.FILELIST = list(files)
    self.sample_sizes = []
    self.total_entries = 0
    self.feature_keys = list(feature_keys)
    self.channel_nr = len(self.feature_keys)
    self.label_key = label_key
    self.label_dtype = int
    self.size = None
    self.normalize = normalize
    for filename in self.FILELIST:
      if not os.path.exists(filename):
        raise FileNotFoundError(f"File {filename} does not exist")
      with utils.open_hdf(filename, "r") as hdf:
        if grid_dim is None and self.size is None:
          self.size = np.array(hdf["featurizer_parms"]["dimensions"][()]).reshape(-1).astype(int)
          if len(self.size) == 1:
            self.size = np.repeat(self.size, 3)
        elif isinstance(grid_dim, (int, float, np.integer)):
          self.size = np.array([grid_dim, grid_dim, grid_dim], dtype=int)
        elif isinstance(grid_dim, (list, tuple, np.ndarray)):
          self.size = np.array(grid_dim, dtype=int)
        keys = list(hdf.keys())
        for k in self.feature_keys:
          if k not in keys:
            raise KeyError(f"Feature key {k} is not in the h5 file: {filename}")
        if self.label_key not in keys:
          self.label_key = None
        else:
          self.label_dtype = np.dtype(hdf[self.label_key].dtype)
        label_nr = hdf[self.feature_keys[0]].shape[0] if self.label_key is None else hdf[self.label_key].shape[0]
        if config.verbose():
          printit(f"Found {label_nr} labels in {filename}")
        self.sample_sizes.append(label_nr)
        self.total_entries += label_nr
    self.file_map = np.full(self.total_entries, 0, dtype=np.uint32)
    self.position_map = np.full(self.total_entries, 0, dtype=np.uint64)
    count = 0
    for fidx, label_nr in enumerate(self.sample_sizes):
      self.file_map[count:count + label_nr] = fidx
      self.position_map[count:count + label_nr] = np.arange(label_nr)
      count += label_nr

  def __len__(self):
    return self.total_entries

  def position(self, index):
    return int(self.position_map[index])

  def filename(self, index):
    return self.FILELIST[self.file_map[index]]

  def __getitem__(self, index):
    import torch
    if index >= self.total_entries:
      raise IndexError(f"Index {index} is out of range. The dataset has {self.total_entries} entries.")
    data = torch.zeros([self.channel_nr] + self.size.tolist(), dtype=torch.float32)
    for i, key in enumerate(self.feature_keys):
      data[i] = torch.from_numpy(np.asarray(readdata(self.filename(index), key, self.position(index), normalize=self.normalize), dtype=np.float32))
    if self.label_key is None:
      label = torch.tensor([0])
    else:
      label = torch.tensor([np.asarray(readdata(self.filename(index), self.label_key, self.position(index))).item()])
    return data, label

  def mini_batch_task(self, index):
    return [(self.filename(index), key, np.s_[self.position(index)], self.normalize) for key in self.feature_keys]

  def _read_batch(self, batch):
    """data [len(batch), channels, x, y, z] float32 and labels [len(batch), 1] for the global indices `batch`,
    each file opened once."""
    data = np.empty([len(batch), self.channel_nr] + self.size.tolist(), dtype=np.float32)
    labels = np.zeros((len(batch), 1), dtype=np.float32 if self.label_key is None else self.label_dtype)
    files = self.file_map[batch]
    for fidx in np.unique(files):
      rows = np.flatnonzero(files == fidx)
      hdf = self._reader(int(fidx))
      dsets = [hdf[key] for key in self.feature_keys]

      def read_rows(part):
        for r in part:
          for c, dset in enumerate(dsets):
            data[r, c] = _normalise(np.asarray(dset[self.position(batch[r])], dtype=np.float32), self.normalize)
      workers = max(1, int(getattr(self, "read_workers", 4)))
      if workers > 1 and len(rows) >= 2 * workers:
        # samples are independent chunks: reading one (file read, inflate) releases the interpreter lock
        list(_read_pool(workers).map(read_rows, np.array_split(rows, workers)))
      else:
        read_rows(rows)
      if self.label_key is not None:
        dset = hdf[self.label_key]
        for r in rows:
          labels[r, 0] = dset[self.position(batch[r])]
    return data, labels

  def _reader(self, fidx):
    """The file of index `fidx`, kept open between batches (what a reader has learnt about a file -- chunk index nodes,
    member headers -- serves every later batch); at most `max_open_files` handles, the least recently used one goes."""
    pool = self.__dict__.setdefault("_open_files", {})
    hdf = pool.pop(fidx, None)
    if hdf is None:
      while len(pool) >= max(1, int(getattr(self, "max_open_files", 64))):
        pool.pop(next(iter(pool))).close()
      hdf = utils.open_hdf(self.FILELIST[fidx], "r")
    pool[fidx] = hdf                       # most recently used last
    return hdf

  def close(self):
    for hdf in self.__dict__.pop("_open_files", {}).values():
      try:
        hdf.close()
      except Exception:
        pass

  def __del__(self):
    self.close()

  def mini_batches(self, batch_nr=None, batch_size=None, shuffle=True, process_nr=None, augment=False, augment_translation=0,
                   augment_add_noise=False, device=None, **kwargs):
    """Generator of (data, label) mini-batches (dataset.py:326-393).  `process_nr` is accepted and ignored;
    `device` (e.g. "cuda") moves each batch there before the augmentation."""
    import torch
    indices = np.arange(self.total_entries)
    if shuffle:
      np.random.shuffle(indices)
    if batch_size is not None:
      batches = split_array(indices, batch_size)
    elif batch_nr is not None:
      batch_size = self.total_entries // batch_nr
      batches = split_array(indices, batch_size)
    else:
      raise ValueError("Either batch_nr or batch_size should be specified")
    if config.verbose():
      printit(f"Iterating the dataset: {len(batches)} batches with batch size {batch_size}.")
    # the next batch is read on one reader thread while the caller works on the current one (reading a chunk -- the file
    # read, the inflate -- releases the interpreter lock); the random draws of shuffle and augmentation stay where they were
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=1, thread_name_prefix="nearl-dataset") as pool:
      ahead = pool.submit(self._read_batch, batches[0]) if len(batches) else None
      for k in range(len(batches)):
        data_numpy, labels_numpy = ahead.result()
        ahead = pool.submit(self._read_batch, batches[k + 1]) if k + 1 < len(batches) else None
        yield self._deliver(data_numpy, labels_numpy, device, augment, augment_translation, augment_add_noise)

  def _deliver(self, data_numpy, labels_numpy, device, augment, augment_translation, augment_add_noise):
    import torch
    data, label = torch.from_numpy(data_numpy), torch.from_numpy(labels_numpy)
    if device is not None:
      data, label = data.to(device, non_blocking=True), label.to(device, non_blocking=True)
    if augment:
      data = data_augment(data, trans=augment_translation, add_noise=augment_add_noise)
    return data, label
