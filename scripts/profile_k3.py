"""One standalone K3 call per reducer family on a 100 x 128^3 series (for ncu): python scripts/profile_k3.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nearl_b200
from nearl_b200 import _lib
L = _lib.lib()
stream = torch.cuda.current_stream()
_lib.check(L.nrl_set_stream(ctypes.c_void_p(stream.cuda_stream), 1))
F, G = 100, 128 ** 3
s = torch.rand((F, G), device="cuda")
out = torch.empty(G, device="cuda")
for agg in (1, 2, 7, 2, 1):
  _lib.check(L.nrl_aggregate_dev(ctypes.c_void_p(s.data_ptr()), F, G, agg, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream.cuda_stream)))
torch.cuda.synchronize()
print("ok", float(out.sum()))
