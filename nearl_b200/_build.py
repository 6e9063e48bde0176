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
  """Compile when the library is missing or older than a source.  Safe with one process per GPU: the build runs under an
  exclusive file lock (the ranks that lose the race wait, then find the library up to date), nvcc writes to a temporary
  name and the finished file is renamed over the old one, so no process ever maps a half-written library."""
  if not force and not needs_build():
    return LIB_PATH
  import fcntl
  with open(LIB_PATH + ".lock", "w") as lock:
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
      if not force and not needs_build():          # another process built it while this one waited
        return LIB_PATH
      tmp = f"{LIB_PATH}.{os.getpid()}.tmp"
      cmd = [_nvcc()] + NVCC_FLAGS + ["-o", tmp, os.path.join(CSRC, "nrl_api.cu")]
      if verbose:
        cmd.insert(1, "-Xptxas=-v")
      res = subprocess.run(cmd, capture_output=True, text=True)
      if res.returncode != 0:
        if os.path.exists(tmp):
          os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
      os.replace(tmp, LIB_PATH)
      if verbose:
        print(res.stderr)
    finally:
      fcntl.flock(lock, fcntl.LOCK_UN)
  return LIB_PATH
