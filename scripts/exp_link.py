"""What the host link sustains for the headline step's transfers: 24.3 MB up alone, with 5.2 MB coming down per upload on another
stream, and with the SMs busy as well (torch only: plumbing, no kernels of the library)."""
import json, os, sys, torch
up_host = torch.empty(24_320_000 // 4, dtype=torch.float32).pin_memory()
dn_host = torch.empty(5_242_880 // 4, dtype=torch.float32).pin_memory()
up_dev = [torch.empty_like(up_host, device="cuda") for _ in range(4)]
dn_dev = torch.empty_like(dn_host, device="cuda")
a = torch.rand(4096, 4096, device="cuda"); b = torch.rand(4096, 4096, device="cuda")
s_up, s_dn, s_k = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
K = 40
def run(down, busy):
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  with torch.cuda.stream(s_up):
    e0.record()
  for i in range(K):
    with torch.cuda.stream(s_up):
      up_dev[i % 4].copy_(up_host, non_blocking=True)
    if down:
      with torch.cuda.stream(s_dn):
        dn_host.copy_(dn_dev, non_blocking=True)
    if busy:
      with torch.cuda.stream(s_k):
        torch.mm(a, b)
  with torch.cuda.stream(s_up):
    e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / K
rows = []
for down, busy in ((0, 0), (1, 0), (1, 1), (0, 0)):
  run(down, busy)
  ms = min(run(down, busy) for _ in range(3))
  rows.append({"download_alongside": bool(down), "sms_busy": bool(busy), "upload_ms": round(ms, 4), "upload_gbs": round(24.32 / ms, 2)})
  print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/exp_link.json", "w"), indent=1)
