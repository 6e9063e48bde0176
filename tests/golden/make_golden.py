"""Generates tests/golden/reference_vectors.npz by running the REFERENCE's own compiled kernels
(oracle/_ref/all_actions*.so, built unmodified from /root/reference/src by oracle/build_ref.sh)
on small seeded inputs.  Needs a GPU (the reference has no CPU path):

  gpurun -- 'python tests/golden/make_golden.py gpurun_out/reference_vectors.npz'

then copy the file to tests/golden/.  Inputs are regenerated from the seeds by the tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, ROOT)


def inputs():
  """Seeded inputs shared by the generator and the tests."""
  rng = np.random.default_rng(20240713)
  dims = np.array([16, 16, 16], dtype=np.int32)
  F, N = 6, 400
  s, c, sig = 0.75, 3.0, 1.2
  coords = rng.uniform(-c, 16 * s + c, size=(F, N, 3)).astype(np.float32)
  coords[2, 390:] = 99999.0
  w_int = rng.integers(1, 7, size=F * N).astype(np.float32)
  w_flt = rng.uniform(0.5, 16.0, size=F * N).astype(np.float32)
  w_flt[2 * N + 390:3 * N] = 0.0
  series = rng.random((9, 4096)).astype(np.float32)
  return dict(dims=dims, F=F, N=N, s=s, c=c, sig=sig, coords=coords, w_int=w_int, w_flt=w_flt, series=series)


def main(path):
  import all_actions as ref
  x = inputs()
  dims, s, c, sig, N = x["dims"], x["s"], x["c"], x["sig"], x["N"]
  out = {}
  out["voxelize_f0"] = ref.frame_voxelize(x["coords"][0], x["w_flt"][:N], dims, s, c, sig, 0)
  assert np.any(out["voxelize_f0"]), "the reference returned zeros: no usable GPU"
  for obs in (1, 2, 3):
    out[f"obs{obs}_frame1"] = ref.frame_observation(x["coords"][1], x["w_int"][N:2 * N], dims, s, c, obs)
    for agg in (1, 2, 3, 5, 6, 8):
      out[f"mo_obs{obs}_agg{agg}"] = ref.marching_observer(x["coords"], x["w_int"], dims, s, c, obs, agg)
  for obs in (11, 12, 13, 14, 15, 16):
    out[f"obs{obs}_frame1"] = ref.frame_observation(x["coords"][1], x["w_flt"][N:2 * N], dims, s, c, obs)
    out[f"mo_obs{obs}_agg1"] = ref.marching_observer(x["coords"], x["w_flt"], dims, s, c, obs, 1)
  for agg in (1, 2, 5):
    out[f"flow_agg{agg}"] = ref.density_flow(x["coords"], x["w_flt"], dims, s, c, sig, agg)
  for agg in range(1, 9):
    out[f"aggregate_{agg}"] = ref.aggregate(x["series"], agg)
  out["summation"] = np.array([ref.summation(x["series"].reshape(-1))], dtype=np.float32)
  np.savez_compressed(path, **{k: np.asarray(v, dtype=np.float32) for k, v in out.items()})
  print("wrote", path, len(out), "vectors")


if __name__ == "__main__":
  main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "reference_vectors.npz"))
