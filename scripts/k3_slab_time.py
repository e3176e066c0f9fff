"""Device-resident K3 time on the bench slabs of several (world, rank) pairs.  Not a test."""
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
out = []
for world, rank in ((1, 0), (2, 1), (4, 1), (4, 2), (4, 3), (8, 4), (8, 7), (3, 1), (5, 2), (6, 5), (7, 3)):
    lam = ila_b200.sharding.grid_rows(1024, 1024 * world, rank, world)
    d_lam = torch.from_numpy(lam).to(dev); n = d_lam.numel()
    d_L = torch.empty(n, dtype=torch.complex128, device=dev); d_st = torch.empty(n, dtype=torch.int32, device=dev)
    args = (3, 1, _lib.ptr(col), _lib.ptr(d_lam), n, 1, _lib.ptr(d_L), None, _lib.ptr(d_st), ctypes.c_void_p(stream))
    for _ in range(3): lib.ila_ql_toeplitz_l11_c64_dev(*args)
    ts = []
    for _ in range(10):
        flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); lib.ila_ql_toeplitz_l11_c64_dev(*args); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    out.append(f"({world},{rank}) {np.median(ts)*1e3:.1f}")
print("slab times (us):", "  ".join(out))
