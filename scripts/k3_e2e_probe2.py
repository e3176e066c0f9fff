"""Interleaved end-to-end timing of the K3 host call variants (the host of a GPU box is shared: copy rates drift between
runs, so every variant is timed in the same round-robin loop and reported as median / min)."""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ila_b200
from ila_b200 import _lib

lib = ila_b200.lib()
dev = torch.device("cuda", 0)
col = np.ascontiguousarray(np.array([2j, 0.0, 0.0, 1.0, 0.7]), dtype=np.complex128)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
x = np.linspace(-10.0, 7.0, 1024); y = np.linspace(-7.0, 8.0, 1024)
lam = (x[None, :] + 1j * y[:, None]).reshape(-1); n = lam.size
pin = lambda t: t.pin_memory()
P = dict(lam=pin(torch.from_numpy(lam.copy())), Lc=pin(torch.empty(n, dtype=torch.complex128)), st=pin(torch.empty(n, dtype=torch.int32)),
         Lr=pin(torch.empty(n, dtype=torch.float64)), s8=pin(torch.empty(n, dtype=torch.uint8)))
Q = dict(lam=torch.from_numpy(lam.copy()), Lc=torch.empty(n, dtype=torch.complex128), st=torch.empty(n, dtype=torch.int32),
         Lr=torch.empty(n, dtype=torch.float64), s8=torch.empty(n, dtype=torch.uint8))
d_in = torch.empty(16 * n, dtype=torch.uint8, device=dev); d_out = torch.empty(9 * n, dtype=torch.uint8, device=dev)
hb_in = pin(torch.empty(16 * n, dtype=torch.uint8)); hb_out = pin(torch.empty(9 * n, dtype=torch.uint8))
s2 = torch.cuda.Stream(dev)
def call_c(B, mode=0):
    def f():
        lib.ila_set_host_copy_mode(mode)
        _lib.check(lib.ila_ql_toeplitz_l11_c64(3, 1, _lib.ptr(col), _lib.ptr(B["lam"]), n, 1, _lib.ptr(B["Lc"]), None, _lib.ptr(B["st"]), 1))
    return f
def call_r(Bin, Bout, mode=0, status=True):
    def f():
        lib.ila_set_host_copy_mode(mode)
        _lib.check(lib.ila_ql_toeplitz_l11re_c64(3, 1, _lib.ptr(col), _lib.ptr(Bin["lam"]), n, 1, _lib.ptr(Bout["Lr"]), _lib.ptr(Bout["s8"]) if status else None, 1))
    return f
def copy_both():
    d_in.copy_(hb_in, non_blocking=True)
    with torch.cuda.stream(s2):
        hb_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
def copy_h2d():
    d_in.copy_(hb_in, non_blocking=True); torch.cuda.synchronize()
def copy_d2h():
    hb_out.copy_(d_out, non_blocking=True); torch.cuda.synchronize()
def call_chunk(minc, div):
    def f():
        lib.ila_set_host_copy_mode(0); lib.ila_set_host_chunking(minc, div)
        _lib.check(lib.ila_ql_toeplitz_l11re_c64(3, 1, _lib.ptr(col), _lib.ptr(P["lam"]), n, 1, _lib.ptr(P["Lr"]), _lib.ptr(P["s8"]), 1))
        lib.ila_set_host_chunking(0, 0)
    return f
variants = {
    "pinned real+uint8 chunks min16K div2": call_chunk(1 << 14, 2),
    "pinned real+uint8 chunks min16K div4": call_chunk(1 << 14, 4),
    "pinned real+uint8 chunks min32K div3": call_chunk(1 << 15, 3),
    "pinned real+uint8 chunks min64K div3": call_chunk(1 << 16, 3),
    "pinned real+uint8 chunks min128K div8": call_chunk(1 << 17, 8),
    "pinned real+uint8 chunks min8K div3": call_chunk(1 << 13, 3),
    "pinned complex+int32 (36 B)": call_c(P),
    "pinned real+uint8 (25 B)": call_r(P, P),
    "pinned real NaN-coded (24 B)": call_r(P, P, status=False),
    "pageable driver (mode1) real+uint8": call_r(Q, Q, 1),
    "pageable staged (mode0) real+uint8": call_r(Q, Q, 0),
    "pageable staged complex+int32": call_c(Q, 0),
    "pageable in, pinned out (shim) real+uint8": call_r(Q, P, 0),
    "bare copies H2D 16n + D2H 9n concurrently": copy_both,
    "bare copy H2D 16n": copy_h2d,
    "bare copy D2H 9n": copy_d2h,
}
for f in variants.values():
    for _ in range(3): f()
T = {k: [] for k in variants}
for rep in range(int(os.environ.get("REPS", "25"))):
    for k, f in variants.items():
        flush.fill_(1); torch.cuda.synchronize()
        t0 = time.perf_counter(); f(); T[k].append((time.perf_counter() - t0) * 1e3)
lib.ila_set_host_copy_mode(0)
out = {k: {"median_ms": float(np.median(v)), "min_ms": float(np.min(v)), "mean_ms": float(np.mean(v))} for k, v in T.items()}
for k, v in out.items(): print(f"{k:48s} median {v['median_ms']:.3f}  min {v['min_ms']:.3f}  mean {v['mean_ms']:.3f} ms")
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_k3_e2e_probe2.json"), "w"), indent=1)
