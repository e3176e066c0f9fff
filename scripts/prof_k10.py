"""K10 workload run (tridiagonal conjugation, 2048 operators x 16384 rows, per-operator X) for ncu."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ila_b200
from ila_b200._lib import lib, check, ptr
dev = torch.device("cuda:0")
b, n = 2048, 16384
g = torch.Generator(device=dev); g.manual_seed(1)
U = torch.rand((b, 3, n + 1), dtype=torch.float64, device=dev, generator=g) * 0.3
U[:, 0] += 1.0
X = torch.rand((b, 3, n + 1), dtype=torch.float64, device=dev, generator=g)
Y = torch.empty((b, 3, n), dtype=torch.float64, device=dev)
s = torch.cuda.current_stream()
for _ in range(3):
    check(lib().ila_triconj_bands_f64_dev(3, b, n, ptr(U), n + 1, ptr(X), n + 1, 0, ptr(Y), None, C.c_void_p(s.cuda_stream)))
torch.cuda.synchronize()
print("ok", float(Y.abs().max()))
