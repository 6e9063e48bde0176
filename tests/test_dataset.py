"""Consumer side of the HDF5 layout (/root/reference/nearl/io/dataset.py): the files the featuriser
writes are read back into channel-stacked mini-batches; augmentation keeps what it must keep."""
import numpy as np
import pytest
import torch

import nearl_b200
from nearl_b200 import dataset, utils

nearl_b200.update_config(quiet=True)


def make_file(path, n, dim, seed, with_label=True):
  rng = np.random.default_rng(seed)
  utils.dump_dict(path, "featurizer_parms", {"dimensions": [dim, dim, dim], "cutoff": 3.0, "outfile": path})
  a = rng.random((n, dim, dim, dim), dtype=np.float32)
  b = rng.random((n, dim, dim, dim), dtype=np.float32) * 5 - 1
  for key, arr in (("mass", a), ("flow", b)):
    for row in arr:                       # the featuriser appends one row per task
      utils.append_hdf_data(path, key, row[None], dtype=np.float32, maxshape=(None, dim, dim, dim), chunks=(1, dim, dim, dim))
  labels = rng.integers(0, 20, size=n)
  if with_label:
    utils.append_hdf_data(path, "label", labels, dtype=int, maxshape=(None,), chunks=True)
  return a, b, labels


def test_dataset_reads_the_featurizer_layout(tmp_path):
  f1, f2 = str(tmp_path / "a.h5"), str(tmp_path / "b.h5")
  a1, b1, l1 = make_file(f1, 5, 8, 1)
  a2, b2, l2 = make_file(f2, 3, 8, 2)
  ds = dataset.Dataset([f1, f2], feature_keys=["mass", "flow"])
  assert len(ds) == 8 and list(ds.size) == [8, 8, 8] and ds.channel_nr == 2
  assert ds.filename(6) == f2 and ds.position(6) == 1
  data, label = ds[6]
  assert data.shape == (2, 8, 8, 8) and np.array_equal(data[0].numpy(), a2[1]) and np.array_equal(data[1].numpy(), b2[1]) and int(label[0]) == l2[1]
  with pytest.raises(IndexError):
    ds[8]
  with pytest.raises(KeyError):
    dataset.Dataset([f1], feature_keys=["nope"])
  got = list(ds.mini_batches(batch_size=3, shuffle=False))
  assert [len(d) for d, _ in got] == [3, 3, 2]
  allx = np.concatenate([a1, a2]); ally = np.concatenate([l1, l2])
  assert np.array_equal(torch.cat([d for d, _ in got])[:, 0].numpy(), allx)
  assert np.array_equal(torch.cat([l for _, l in got]).numpy().reshape(-1), ally)
  np.random.seed(3)
  shuffled = list(ds.mini_batches(batch_nr=2, shuffle=True))
  assert sorted(torch.cat([l for _, l in shuffled]).reshape(-1).tolist()) == sorted(ally.tolist())
  # per-sample normalisation only for the two named modes (normalize=True is a no-op, as in the reference)
  z = dataset.Dataset([f1], feature_keys=["flow"], normalize="zscore")[2][0][0].numpy()
  assert abs(z.mean()) < 1e-5 and abs(z.std() - 1) < 1e-4
  m = dataset.readdata(f1, "flow", 2, normalize="minmax")
  assert m.min() == 0.0 and m.max() == 1.0
  # files without labels
  f3 = str(tmp_path / "c.h5")
  make_file(f3, 2, 8, 3, with_label=False)
  nol = dataset.Dataset([f3], feature_keys=["mass"])
  assert nol.label_key is None and len(nol) == 2 and int(nol[0][1][0]) == 0


def test_split_array_matches_reference_shapes():
  assert [len(b) for b in dataset.split_array(np.arange(10), 5)] == [5, 5]
  assert [len(b) for b in dataset.split_array(np.arange(11), 5)] == [5, 5, 1]
  assert [len(b) for b in dataset.split_array(np.arange(3), 5)] == [3]


def test_augmentation_invariants():
  rng = np.random.default_rng(0)
  x = torch.from_numpy(rng.random((2, 3, 6, 6, 6), dtype=np.float32))
  np.random.seed(5)
  for _ in range(8):
    y = dataset.data_augment(x.clone())
    assert y.shape == x.shape
    # flips and quarter turns permute the voxels of every (sample, channel) grid
    assert torch.equal(torch.sort(y.reshape(6, -1), dim=1).values, torch.sort(x.reshape(6, -1), dim=1).values)
  np.random.seed(7)
  y = dataset.data_augment(x.clone(), trans=2)
  kept = y[y != 0]
  assert len(kept) <= x.numel() and set(np.round(kept.numpy(), 6)).issubset(set(np.round(x.numpy().reshape(-1), 6)))
  y = dataset.data_augment(x.clone(), add_noise=True)
  assert y.shape == x.shape and not torch.equal(torch.sort(y.reshape(-1)).values, torch.sort(x.reshape(-1)).values)


def test_batches_read_on_threads_are_the_batches_read_on_one(tmp_path):
  """A batch is read by a few threads (samples are independent chunks), the next batch while the caller works on the
  current one, the files stay open between batches: the same batches in the same order, whatever the number of threads
  and of open handles."""
  files, arrays = [], []
  for k in range(3):
    path = str(tmp_path / f"f{k}.h5")
    a, b, labels = make_file(path, 21 + k, 6, 10 + k)
    files.append(path)
    arrays.append((a, b, labels))
  want = None
  for workers, handles in ((1, 64), (4, 64), (3, 1)):
    ds = dataset.Dataset(files, feature_keys=["mass", "flow"], normalize="zscore")
    ds.read_workers, ds.max_open_files = workers, handles
    np.random.seed(5)
    got = [(d.numpy().copy(), l.numpy().copy()) for d, l in ds.mini_batches(batch_size=16, shuffle=True)]
    assert sum(len(d) for d, _ in got) == len(ds) == 66
    if want is None:
      want = got
    for (d0, l0), (d1, l1) in zip(want, got):
      assert np.array_equal(d0, d1) and np.array_equal(l0, l1)
    assert len(ds.__dict__.get("_open_files", {})) <= handles
    ds.close()
    assert not ds.__dict__.get("_open_files")
  # unshuffled, unnormalised: the rows of the files in order
  ds = dataset.Dataset(files, feature_keys=["mass", "flow"])
  flat = np.concatenate([d.numpy() for d, _ in ds.mini_batches(batch_size=25, shuffle=False)])
  assert np.array_equal(flat[:, 0], np.concatenate([a for a, _, _ in arrays])) and np.array_equal(flat[:, 1], np.concatenate([b for _, b, _ in arrays]))
