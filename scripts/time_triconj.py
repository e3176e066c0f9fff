"""K10 tridiagonal conjugation: device-resident kernel time (CUDA events on the launching stream) and achieved GB/s.
Not a test. Run under gpurun:  python scripts/time_triconj.py"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ila_b200
from ila_b200._lib import lib, check, ptr

L = lib()
dev = torch.device("cuda:0")
s = torch.cuda.current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s); fn(); e1.record(s); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


for nbands, batch, n, shared in ((3, 2048, 16384, True), (3, 2048, 16384, False), (2, 2048, 16384, True), (3, 64, 1 << 20, True)):
    m = n + 1
    g = torch.Generator(device=dev); g.manual_seed(1)
    U = torch.rand((batch, nbands, m), dtype=torch.float64, device=dev, generator=g) * 0.3
    U[:, 0] += 1.0
    X = torch.rand((3, m) if shared else (batch, 3, m), dtype=torch.float64, device=dev, generator=g)
    Y = torch.empty((batch, 3, n), dtype=torch.float64, device=dev)
    fn = lambda: check(L.ila_triconj_bands_f64_dev(nbands, batch, n, ptr(U), m, ptr(X), m, 1 if shared else 0, ptr(Y), None, C.c_void_p(s.cuda_stream)))
    ms = timeit(fn)
    rows = batch * n
    by = rows * 8 * (nbands + 3 + (0 if shared else 3))
    print(f"bands nb={nbands} batch={batch} n={n} X {'shared' if shared else 'per-op'}: {ms:.3f} ms, {rows/ms/1e6:.2f} Grows/s, {by/ms/1e6:.0f} GB/s algorithmic")

# fused Cholesky -> conjugation through the host entry (wall time)
import time
n, m = 4096, 4110
k = np.arange(1, m + 1, dtype=np.float64)
e = k / np.sqrt(4 * k * k - 1)
X = np.stack([e, np.zeros(m), e])
sigma = np.linspace(1.05, 3.0, 2048)
head = np.stack([-e[: n + 6], np.zeros(n + 6)])
p = np.array([[0.0, 0.0]]); q = np.zeros((1, 2))
for _ in range(3):
    t0 = time.perf_counter()
    gq = ila_b200.chol_tridiagonal_conjugation(1, p, q, X, n, shift=-sigma, head=head)
    dt = time.perf_counter() - t0
print(f"fused chol->conj 2048 shifts x n={n}: {dt*1e3:.2f} ms host call, bad {int((gq['status'] != 0).sum())}")
