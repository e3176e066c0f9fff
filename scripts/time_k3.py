"""K3-only timings over shifts-per-thread settings (not a test). Run under gpurun."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ila_b200
from ila_b200 import _lib
dev = torch.device("cuda", 0)
col = np.array([2j, 0, 0, 1, 0.7])
x = np.linspace(-10, 7, 1024); y = np.linspace(-7, 8, 1024)
lam = torch.from_numpy((x[None, :] + 1j * y[:, None]).reshape(-1)).to(dev)
L = torch.empty_like(lam); st = torch.empty(lam.numel(), dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
def ev_time(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return np.median(ts), np.min(ts)
for spt in (0, 4, 8, 16):
    _lib.check(_lib.lib().ila_ql_set_shifts_per_thread(spt))
    med, mn = ev_time(lambda: ila_b200.ql_toeplitz_l11_dev(col, lam, L, st, stream=s))
    print(f"spt={spt}: median {med*1e3:.1f} us min {mn*1e3:.1f} us -> {lam.numel()/med/1e-3/1e9:.2f} Gpts/s; bad status {int((st!=0).sum())}")
