"""Which transition inside the slow run of 8 shifts is pathological?  Not a test."""
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

def t_of(run8):
    d_lam.copy_(torch.from_numpy(np.ascontiguousarray(np.tile(np.asarray(run8, dtype=np.complex128), n // 8))).to(dev))
    lib.ila_ql_toeplitz_l11_c64_dev(*args)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); lib.ila_ql_toeplitz_l11_c64_dev(*args); e1.record(); e1.synchronize()
    return round(e0.elapsed_time(e1) * 1e3, 1)

p = x[992:1000] + 1j * y
print("actual run", t_of(p), "reversed", t_of(p[::-1]), "constant p0", t_of([p[0]] * 8))
print("first k points then constant:", [t_of(list(p[:k]) + [p[k - 1]] * (8 - k)) for k in range(1, 9)])
print("start at point k (constant after 2):", [t_of([p[k], p[k + 1]] + [p[k + 1]] * 6) for k in range(0, 7)])
q = x[984:992] + 1j * y
print("neighbour run 984:", t_of(q), " spacing x2:", t_of(x[992:1008:2] + 1j * y), " y shifted 1e-3:", t_of(p + 1e-3j))
