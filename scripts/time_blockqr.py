"""K11 adaptive block-banded QR: device-resident kernel time (CUDA events on the launching stream).  Not a test.
   python scripts/time_blockqr.py [batch]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ila_b200
from ila_b200._lib import lib, check, ptr

L = lib()
dev = torch.device("cuda:0")
s = torch.cuda.current_stream()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 592
if len(sys.argv) > 2:
    check(L.ila_qr_blockbanded_set_mode(int(sys.argv[2])))
maxblocks = 64
m = 80
T = torch.tensor(np.stack([np.full(m, 0.5), np.zeros(m), np.full(m, 0.5)]), device=dev)
lam = torch.linspace(4.0, 9.0, batch, dtype=torch.float64, device=dev)
al = torch.ones(batch, dtype=torch.float64, device=dev); be = torch.ones(batch, dtype=torch.float64, device=dev)
b = torch.ones((batch, 1), dtype=torch.float64, device=dev)
per = np.zeros(1, dtype=np.int64)
check(L.ila_qr_blockbanded_work_doubles(maxblocks, ptr(per)))
xld = maxblocks * (maxblocks + 1) // 2
work = torch.empty(batch * int(per[0]), dtype=torch.float64, device=dev)
qtb = torch.zeros((batch, xld), dtype=torch.float64, device=dev); x = torch.empty((batch, xld), dtype=torch.float64, device=dev)
ncols = torch.zeros(batch, dtype=torch.int64, device=dev); nblk = torch.zeros(batch, dtype=torch.int32, device=dev)
st = torch.zeros(batch, dtype=torch.int32, device=dev)
fn = lambda: check(L.ila_qr_blockbanded_solve_f64_dev(batch, ptr(al), ptr(be), ptr(lam), ptr(T), ptr(T), m, 1, ptr(b), 1e-30, 300, maxblocks,
                                                      ptr(work), ptr(qtb), ptr(x), xld, ptr(ncols), ptr(nblk), ptr(st), C.c_void_p(s.cuda_stream)))
fn(); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(s); fn(); e1.record(s); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
nb_ = nblk.cpu().numpy().astype(np.int64)
# Householder flops of the factorization: block column J costs J columns x 4 (2J+1-j)(3J+3-j) (dot + update)
flops = sum(sum(4.0 * (2 * J + 1 - j) * (3 * J + 3 - j) for j in range(J)) for K in nb_ for J in range(1, K + 1))
cols = int(sum(K * (K + 1) // 2 for K in nb_))
print(f"batch {batch}: {ms:.3f} ms, blocks {nb_.min()}..{nb_.max()}, bad {int((st != 0).sum())}, {cols/ms/1e3:.2f} Mcolumns/s, "
      f"{flops/ms/1e9:.2f} TFLOP/s FP64 (Householder flops), {batch/ms*1e3:.0f} operators/s")

if os.environ.get("K11_PROFILE"):
    xc = x.cpu().numpy()
    i = int(np.argmax(nb_))
    print("phase cycles (operator with most blocks): load %.0f  first-norm %.0f  columns %.0f  write-out %.0f  back-substitution %.0f" % tuple(xc[i, -5:]))
    print("panel loop work before the barrier: warp 0 (panel) %.0f cycles, warp 1 (trailing) %.0f cycles" % (xc[i, -6], xc[i, -7]))
    print("warp 0 per panel step (%d steps): apply2 %.0f  two-warp barrier %.0f  panel_factor %.0f  block barrier wait %.0f cycles" %
          (xc[i, -12], xc[i, -8] / xc[i, -12], xc[i, -9] / xc[i, -12], xc[i, -10] / xc[i, -12], xc[i, -11] / xc[i, -12]))
