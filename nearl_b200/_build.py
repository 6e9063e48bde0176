"""Builds libnearl_b200.so in-tree with nvcc for sm_100a (no torch needed to build)."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libnearl_b200.so")

NVCC_FLAGS = [
  "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
  # arithmetic contract of csrc/common.cuh: no implicit FMA contraction, flush-to-zero like the reference build
  "-fmad=false", "-ftz=true",
  "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
  for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
    if cand and os.path.exists(cand):
      return cand
  raise RuntimeError("nvcc not found: nearl_b200 needs the CUDA toolkit to build its sm_100a kernels")


def sources():
  out = [os.path.join(_HERE, "..", "include", "nearl_b200.h")]
  for name in sorted(os.listdir(CSRC)):
    if name.endswith((".cu", ".cuh", ".inc")):
      out.append(os.path.join(CSRC, name))
  return out


def needs_build():
  if not os.path.exists(LIB_PATH):
    return True
  t = os.path.getmtime(LIB_PATH)
  return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
  if not force and not needs_build():
    return LIB_PATH
  cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "nrl_api.cu")]
  if verbose:
    cmd.insert(1, "-Xptxas=-v")
  res = subprocess.run(cmd, capture_output=True, text=True)
  if res.returncode != 0:
    raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
  if verbose:
    print(res.stderr)
  return LIB_PATH
