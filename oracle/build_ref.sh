#!/usr/bin/env bash
# Compiles the UNMODIFIED reference kernels + pybind module from where they lie under
# /root/reference/src into oracle/_ref/all_actions*.so, for sm_100, with the reference's
# own nvcc flags (src/Makefile:15: -std=c++17 -use_fast_math -O3).  No reference source
# is copied into the repo; the output directory is git-ignored but travels to the GPU
# box.  The module only does useful work on a machine with a GPU (the reference has no
# CPU path and never checks a CUDA error); it is used by tests/ -m gpu as a second
# oracle and by bench.py --impl reference.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${NEARL_REFERENCE_SRC:-/root/reference/src}"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then echo "reference sources not present at $REF; keeping prebuilt files" >&2; exit 0; fi
mkdir -p "$OUT"
PYINC="$(python3 -c 'import sysconfig; print(sysconfig.get_path("include"))')"
PBINC="$(python3 -c 'import pybind11; print(pybind11.get_include())')"
SUFFIX="$(python3 -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
TARGET="$OUT/all_actions$SUFFIX"
if [ -f "$TARGET" ] && [ "$TARGET" -nt "$REF/voxelize.cu" ] && [ "$TARGET" -nt "$REF/marching_observers.cu" ] \
   && [ "$TARGET" -nt "$REF/gpuutils.cu" ] && [ "$TARGET" -nt "$REF/actions_py.cpp" ]; then
  exit 0
fi
"$NVCC" -Xcompiler -fPIC -std=c++17 -use_fast_math -O3 -arch=sm_100 -w \
  -I"$REF" -I"$PYINC" -I"$PBINC" -shared -o "$TARGET" \
  -x cu "$REF/actions_py.cpp" "$REF/gpuutils.cu" "$REF/marching_observers.cu" "$REF/voxelize.cu"
echo "built $TARGET"
