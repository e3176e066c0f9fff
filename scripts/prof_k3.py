"""K3 run for ncu: the launch bench.py times (compact form, 1024 x 1024 Bull-head grid)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ila_b200
from ila_b200 import _lib
lib = _lib.lib()
spt = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
_lib.check(lib.ila_ql_set_shifts_per_thread(spt))
dev = torch.device("cuda", 0)
col = np.ascontiguousarray(np.array([2j, 0, 0, 1, 0.7]), dtype=np.complex128)
x = np.linspace(-10, 7, 1024); y = np.linspace(-7, 8, rows)
lam = torch.from_numpy((x[None, :] + 1j * y[:, None]).reshape(-1)).to(dev)
n = lam.numel()
L = torch.empty(n, dtype=torch.float64, device=dev); st = torch.empty(n, dtype=torch.uint8, device=dev)
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(4):
    _lib.check(lib.ila_ql_toeplitz_l11re_c64_dev(3, 1, _lib.ptr(col), _lib.ptr(lam), n, 1, _lib.ptr(L), _lib.ptr(st), s))
torch.cuda.synchronize()
print("ok", int((st != 0).sum().item()))
