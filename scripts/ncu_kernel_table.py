"""Per-kernel table (first launch of each kernel) from `ncu -i X.ncu-rep --page raw --csv`:
DRAM traffic, warp instructions, issue-slot utilisation, duration.  bench.py reads the result
(profiles/kernel_table.json) for `roofline.traffic` and the compute-side roofline.
usage: python scripts/ncu_kernel_table.py raw.csv > profiles/kernel_table.json"""
import csv, json, re, sys

CATEGORY = {0: "tile_voxelize", 1: "tile_existence", 2: "tile_direct_count", 3: "tile_distinct_count", 11: "tile_mean_distance",
            12: "tile_cumulative_weight", 13: "tile_density", 14: "tile_dispersion", 15: "tile_eccentricity",
            16: "tile_radius_of_gyration", 100: "tile_fused"}

def category(name):
  m = re.search(r"nrl_(cell_)?tile_kernel<\(?(?:int\))?(\d+)", name)
  if m:
    return CATEGORY.get(int(m.group(2)), "tile_?")
  for key, cat in (("crop_pack", "crop_pack"), ("crop_stream", "crop_pack"), ("crop_place", "crop_pack"), ("bin_kernel", "cell_bin"), ("atom_norm", "atom_norm"), ("pack_legacy", "pack_legacy"),
                   ("aggregate", "aggregate")):
    if key in name:
      return cat
  return None

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, key):
  try:
    return float(r[ix[key]].replace(",", ""))
  except Exception:
    return None
units = rows[1]
def to_bytes(r, key):
  v = val(r, key)
  if v is None:
    return None
  return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[ix[key]], 1)
out = {}
seen_kernels = set()
for r in rows[2:]:
  cat = category(r[ix["Kernel Name"]])
  if cat is None or r[ix["Kernel Name"]] in seen_kernels:
    continue
  seen_kernels.add(r[ix["Kernel Name"]])
  if cat in out:
    # a category served by two kernels of one step (the streaming crop and its placement pass): add them up
    dur = val(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1.0)
    out[cat]["kernel"] += " + " + r[ix["Kernel Name"]]
    out[cat]["duration_us"] += dur
    out[cat]["dram_bytes"] += (to_bytes(r, "dram__bytes_read.sum") or 0) + (to_bytes(r, "dram__bytes_write.sum") or 0)
    out[cat]["warp_instructions"] += val(r, "smsp__inst_executed.sum") or 0
    continue
  dur = val(r, "gpu__time_duration.sum")
  dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1.0)
  out[cat] = {"kernel": r[ix["Kernel Name"]], "duration_us": dur_us,
              "dram_bytes": (to_bytes(r, "dram__bytes_read.sum") or 0) + (to_bytes(r, "dram__bytes_write.sum") or 0),
              "warp_instructions": val(r, "smsp__inst_executed.sum"),
              "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
              "regs": val(r, "launch__registers_per_thread"), "grid": val(r, "launch__grid_size")}
json.dump({"source": sys.argv[1], "kernels": out}, sys.stdout, indent=1)
