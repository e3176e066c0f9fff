"""Which runs of 8 shifts in the slow segment are pathological?  Not a test."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ila_b200
from ila_b200 import _lib
lib = ila_b200.lib()
dev = torch.device("cuda:0")
col = np.ascontiguousarray(np.array([2j, 0, 0, 1, 0.7]), dtype=np.complex128)
stream = torch.cuda.current_stream().cuda_stream
n = 1 << 20
d_lam = torch.empty(n, dtype=torch.complex128, device=dev)
d_L = torch.empty(n, dtype=torch.complex128, device=dev); d_st = torch.empty(n, dtype=torch.int32, device=dev)
args = (3, 1, _lib.ptr(col), _lib.ptr(d_lam), n, 1, _lib.ptr(d_L), None, _lib.ptr(d_st), ctypes.c_void_p(stream))
x = np.linspace(-10, 7, 1024)
y = -0.92822265625

def t_of(lam):
    d_lam.copy_(torch.from_numpy(np.ascontiguousarray(lam)).to(dev))
    lib.ila_ql_toeplitz_l11_c64_dev(*args)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); lib.ila_ql_toeplitz_l11_c64_dev(*args); e1.record(); e1.synchronize()
    return round(e0.elapsed_time(e1) * 1e3, 1)

print("runs of 8 (columns 960..1023):", [t_of(np.tile(x[s:s + 8] + 1j * y, n // 8)) for s in range(960, 1024, 8)])
print("single points as a constant run (columns 1008..1023):", [t_of(np.full(n, x[c] + 1j * y)) for c in range(1008, 1024)])
for s in range(960, 1024, 8):
    lam = x[s:s + 8] + 1j * y
    L, st = ila_b200.ql_toeplitz_l11(col, lam)
    print(s, st, np.abs(L))
