"""K3 run for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ila_b200
from ila_b200 import _lib
spt = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_lib.check(_lib.lib().ila_ql_set_shifts_per_thread(spt))
dev = torch.device("cuda", 0)
col = np.array([2j, 0, 0, 1, 0.7])
x = np.linspace(-10, 7, 1024); y = np.linspace(-7, 8, 1024)
lam = torch.from_numpy((x[None, :] + 1j * y[:, None]).reshape(-1)).to(dev)
L = torch.empty_like(lam); st = torch.empty(lam.numel(), dtype=torch.int32, device=dev)
for _ in range(3):
    ila_b200.ql_toeplitz_l11_dev(col, lam, L, st, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", int((st != 0).sum().item()))
