"""Which rows of the Bull-head grid make a slab slow?  Times K3 on row bands (tiled to 1M points) for two y-offsets.  Not a test."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ila_b200
from ila_b200 import _lib
lib = ila_b200.lib()
dev = torch.device("cuda:0")
col = np.ascontiguousarray(np.array([2j, 0, 0, 1, 0.7]), dtype=np.complex128)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream

def time_lam(lam, reps=10):
    d_lam = torch.from_numpy(np.ascontiguousarray(lam)).to(dev)
    n = d_lam.numel()
    d_L = torch.empty(n, dtype=torch.complex128, device=dev); d_st = torch.empty(n, dtype=torch.int32, device=dev)
    args = (3, 1, _lib.ptr(col), _lib.ptr(d_lam), n, 1, _lib.ptr(d_L), None, _lib.ptr(d_st), ctypes.c_void_p(stream))
    for _ in range(3): lib.ila_ql_toeplitz_l11_c64_dev(*args)
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); lib.ila_ql_toeplitz_l11_c64_dev(*args); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3, int((d_st != 0).sum().item())

x = np.linspace(-10, 7, 1024)
for off in (0.0, 0.25, 0.5, 0.75):
    dy = 15.0 / 1023
    y = -7 + (np.arange(1024) + off) * dy * (1023 / (1023 + 1))      # 1024 rows inside [-7, 8)
    lam = (x[None, :] + 1j * y[:, None])
    t, bad = time_lam(lam.reshape(-1))
    row = []
    for b in range(8):
        band = np.tile(lam[b * 128:(b + 1) * 128], (8, 1)).reshape(-1)
        row.append(time_lam(band)[0])
    print(f"offset {off}: full {t:.1f} us (status!=0: {bad}); bands of 128 rows x8: " + " ".join(f"{v:.1f}" for v in row))
# x-offset instead
for off in (0.0, 0.5):
    dx = 17.0 / 1023
    xo = -10 + (np.arange(1024) + off) * dx * (1023 / 1024)
    y = np.linspace(-7, 8, 1024)
    t, bad = time_lam((xo[None, :] + 1j * y[:, None]).reshape(-1))
    print(f"x-offset {off}: full {t:.1f} us")
