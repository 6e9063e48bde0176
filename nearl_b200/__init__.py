"""nearl_b200 -- B200-native featurisation hot path behind Nearl's API.

Package surface mirrors the reference package root (/root/reference/nearl/__init__.py:8-100):
``CONFIG`` / ``config`` accessors, ``update_config`` and ``printit``.  Sub-modules
(``all_actions``, ``commands``, ``features``, ``featurizer``, ``io``, ``utils``) are imported
on demand, exactly like ``import nearl.featurizer`` in the reference.
"""
import os
import sys
import time
from datetime import datetime

__version__ = "0.1.0"

CONFIG = {
  "tempfolder": "/tmp/",
  "clear": True,
  "verbose": False,
  "usegpu": True,
  "debug": False,
  "reportdatetime": False,
  # nearl_b200 additions: silence per-message stdout and the nearl.log append of printit
  "quiet": False,
}


class config:
  """Static accessors of the global configuration (reference nearl/__init__.py:17-35)."""
  @staticmethod
  def verbose():
    return CONFIG.get("verbose", False)

  @staticmethod
  def tempfolder():
    return CONFIG.get("tempfolder", "/tmp")

  @staticmethod
  def debug():
    return CONFIG.get("debug", False)

  @staticmethod
  def usegpu():
    return CONFIG.get("usegpu", False)

  @staticmethod
  def clear():
    return CONFIG.get("clear", True)

  @staticmethod
  def reportdatetime():
    return CONFIG.get("reportdatetime", False)

  @staticmethod
  def quiet():
    return CONFIG.get("quiet", False)


def update_config(dict_to_update=None, **kwargs):
  """Unknown keys raise KeyError (reference nearl/__init__.py:37-48)."""
  merged = dict(dict_to_update or {})
  merged.update(kwargs)
  for key, value in merged.items():
    if key not in CONFIG:
      raise KeyError(f"Key {key} is not in the configuration file")
    CONFIG[key] = value


_start_time = time.perf_counter()
_log_lines = []


def _flush_log():
  if not _log_lines:
    return
  try:
    with open(os.path.join(os.path.abspath(config.tempfolder()), "nearl.log"), "a") as fh:
      fh.write("\n".join(_log_lines) + "\n")
  except OSError:
    pass
  _log_lines.clear()


def printit(*arg, **kwarg):
  """Timestamped message to stdout (warnings/errors to stderr) and to <tempfolder>/nearl.log
  (reference nearl/__init__.py:78-100); the log is buffered instead of reopened per message."""
  if config.quiet():
    return
  if config.verbose() or config.reportdatetime() or config.debug():
    prefix = f"{datetime.now().strftime('%y-%m-%dT%H:%M:%S')}: "
  else:
    prefix = f"Running {time.perf_counter() - _start_time:8.2f}: "
  text = " ".join(str(i) for i in arg)
  first = str(arg[0]) if arg else ""
  stream = sys.stderr if ("Warning" in first or "Error" in first) else sys.stdout
  print(prefix + text, file=stream, **kwarg)
  _log_lines.append(text)
  if len(_log_lines) >= 64:
    _flush_log()


import atexit as _atexit
_atexit.register(_flush_log)
