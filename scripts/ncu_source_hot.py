"""Hot SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv`.
usage: python scripts/ncu_source_hot.py src.csv '<kernel substring>' [min share %]"""
import csv, sys
path, pat = sys.argv[1], sys.argv[2]
share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
rows = list(csv.reader(open(path)))
sections, cur = [], None
for r in rows:
  if r and r[0] == "Kernel Name":
    cur = {"name": r[1], "rows": []}
    sections.append(cur)
  elif cur is not None:
    cur["rows"].append(r)
for s in sections:
  if pat not in s["name"]:
    continue
  hdr = s["rows"][0]
  ix = {h: i for i, h in enumerate(hdr)}
  body = [r for r in s["rows"][1:] if len(r) == len(hdr)]
  tot = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
  print(f"== {s['name']}: {len(body)} SASS instructions, {tot} warp instructions executed")
  for k, r in enumerate(body):
    n = int(r[ix["Instructions Executed"]] or 0)
    if n * 100.0 >= tot * share:
      print(f"{k:5d} {100.0*n/tot:6.2f}% smp={r[ix['# Samples']]:>6} {r[ix['Source']][:110]}")
  break
