"""Per-row cost of K3 in the slow band (each row tiled to 1M points).  Not a test."""
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

def time_row(y):
    d_lam.copy_(torch.from_numpy(np.tile(x + 1j * y, 1024)).to(dev))
    lib.ila_ql_toeplitz_l11_c64_dev(*args)
    ts = []
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); lib.ila_ql_toeplitz_l11_c64_dev(*args); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3

for off in (0.0, 0.5):
    ys = -7 + (np.arange(384, 512) + off) * (15.0 / 1024)
    ts = np.array([time_row(y) for y in ys])
    order = np.argsort(-ts)[:6]
    print(f"offset {off}: median row {np.median(ts):.1f} us, top rows:", [(int(384 + i), round(float(ys[i]), 5), round(float(ts[i]), 1)) for i in order])
# columns of the slowest row: which x make it slow?  (tile 64-point segments)
off = 0.5
ys = -7 + (np.arange(384, 512) + off) * (15.0 / 1024)
ts = np.array([time_row(y) for y in ys])
ybad = ys[int(np.argmax(ts))]
seg = []
for s0 in range(0, 1024, 64):
    lam = np.tile(x[s0:s0 + 64] + 1j * ybad, n // 64)
    d_lam.copy_(torch.from_numpy(lam).to(dev))
    lib.ila_ql_toeplitz_l11_c64_dev(*args)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); lib.ila_ql_toeplitz_l11_c64_dev(*args); e1.record(); e1.synchronize()
    seg.append(round(e0.elapsed_time(e1) * 1e3, 1))
print("slowest row y =", ybad, "per 64-column segment (x from -10):", seg)
