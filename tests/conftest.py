import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_available():
  try:
    from nearl_b200 import _lib
    return _lib.device_count() > 0
  except Exception:
    return False


def pytest_collection_modifyitems(config, items):
  # `-m gpu` on a machine without a device must fail loudly, not skip: only auto-skip when the
  # user did not ask for gpu tests explicitly.
  if "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or ""):
    return
  if _cuda_available():
    return
  skip = pytest.mark.skip(reason="no CUDA device")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


@pytest.fixture(scope="session")
def ref_actions():
  """The reference's own kernels (oracle/_ref, built by oracle/build_ref.sh); GPU box only."""
  ref_dir = os.path.join(ROOT, "oracle", "_ref")
  if not os.path.isdir(ref_dir):
    pytest.skip("oracle/_ref not built")
  sys.path.insert(0, ref_dir)
  try:
    import all_actions
  except Exception as exc:
    pytest.skip(f"reference module not importable: {exc}")
  return all_actions
