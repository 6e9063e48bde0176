"""nrl_superpose on a device-resident trajectory (for ncu): python scripts/profile_superpose.py [F] [N]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nearl_b200 import _lib
F = int(sys.argv[1]) if len(sys.argv) > 1 else 200
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
L = _lib.lib()
xyz = torch.rand(F, N, 3, device="cuda") * 50
idx = np.arange(0, N, 16, dtype=np.int32)
st = torch.cuda.current_stream()
for _ in range(2):
  _lib.check(L.nrl_superpose(ctypes.c_void_p(xyz.data_ptr()), 0, F, N, _lib.ptr(idx), len(idx), 0, 1, None, ctypes.c_void_p(st.cuda_stream)))
torch.cuda.synchronize()
print("ok")
