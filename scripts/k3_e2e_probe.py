"""K3 probes (run under gpurun): (1) kernel time against slab size and shifts-per-thread — the sub-wave regime a
strong-scaled 1024^2 grid puts every GPU in; (2) end-to-end host-call time for pinned / pageable caller buffers,
ComplexF64 + int32 output against the compact Float64 (+ uint8) output, direct copies against the staging ring."""
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
stream = torch.cuda.current_stream(dev).cuda_stream
out = {}

def grid(nx, ny):
    x = np.linspace(-10.0, 7.0, nx); y = np.linspace(-7.0, 8.0, ny)
    return (x[None, :] + 1j * y[:, None]).reshape(-1)

def dev_time(n_rows, nx, spt, reps=10, real=True):
    lam = torch.from_numpy(grid(nx, n_rows)).to(dev)
    n = lam.numel()
    L = torch.empty(n, dtype=torch.float64 if real else torch.complex128, device=dev)
    st = torch.empty(n, dtype=torch.uint8 if real else torch.int32, device=dev)
    _lib.check(lib.ila_ql_set_shifts_per_thread(spt))
    fn = lib.ila_ql_toeplitz_l11re_c64_dev if real else lib.ila_ql_toeplitz_l11_c64_dev
    args = (3, 1, _lib.ptr(col), _lib.ptr(lam), n, 1, _lib.ptr(L), _lib.ptr(st), ctypes.c_void_p(stream)) if real else \
        (3, 1, _lib.ptr(col), _lib.ptr(lam), n, 1, _lib.ptr(L), None, _lib.ptr(st), ctypes.c_void_p(stream))
    for _ in range(3):
        fn(*args)
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(*args); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    _lib.check(lib.ila_ql_set_shifts_per_thread(0))
    return float(np.median(ts)), float(np.min(ts))

print("== kernel time (µs, median/min) vs slab rows x 1024 and spt")
tab = {}
for rows in (16, 32, 64, 128, 256, 512, 1024, 4096):
    for spt in (0, 1, 2, 4, 8, 16):
        med, mn = dev_time(rows, 1024, spt)
        tab[f"{rows}x1024 spt{spt}"] = [med, mn]
        print(f"rows {rows:5d} spt {spt:2d}: {med:8.1f} {mn:8.1f} µs   {rows*1024/med/1e3:7.2f} G/s", flush=True)
med, mn = dev_time(8192, 8192, 0, reps=3)
tab["8192x8192 spt0"] = [med, mn]
print(f"8192x8192 auto: {med:.1f} µs  {8192*8192/med/1e3:.2f} G/s")
med, mn = dev_time(1024, 1024, 0, real=False)
tab["1024x1024 complex-out spt0"] = [med, mn]
print(f"1024x1024 complex+int32 output: {med:.1f} µs")
out["kernel_us"] = tab

print("== e2e host call (ms, median of 10) on the 1024x1024 grid")
lam = grid(1024, 1024); n = lam.size
def host_time(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts)), float(np.min(ts))
e2e = {}
def pinned(t): return t.pin_memory()
for bufkind in ("pinned", "pageable"):
    mk = pinned if bufkind == "pinned" else (lambda t: t)
    h_lam = mk(torch.from_numpy(lam.copy())); h_Lc = mk(torch.empty(n, dtype=torch.complex128)); h_st = mk(torch.empty(n, dtype=torch.int32))
    h_Lr = mk(torch.empty(n, dtype=torch.float64)); h_s8 = mk(torch.empty(n, dtype=torch.uint8))
    for mode in ((0,) if bufkind == "pinned" else (1, 2)):
        _lib.check(lib.ila_set_host_copy_mode(mode))
        name = f"{bufkind} mode{mode}"
        e2e[name + " complex+int32 (36 B)"] = host_time(lambda: _lib.check(lib.ila_ql_toeplitz_l11_c64(3, 1, _lib.ptr(col), _lib.ptr(h_lam), n, 1, _lib.ptr(h_Lc), None, _lib.ptr(h_st), 1)))
        e2e[name + " real+uint8 (25 B)"] = host_time(lambda: _lib.check(lib.ila_ql_toeplitz_l11re_c64(3, 1, _lib.ptr(col), _lib.ptr(h_lam), n, 1, _lib.ptr(h_Lr), _lib.ptr(h_s8), 1)))
        e2e[name + " real, NaN-coded status (24 B)"] = host_time(lambda: _lib.check(lib.ila_ql_toeplitz_l11re_c64(3, 1, _lib.ptr(col), _lib.ptr(h_lam), n, 1, _lib.ptr(h_Lr), None, 1)))
    _lib.check(lib.ila_set_host_copy_mode(0))
# cudaHostRegister cost for the three buffers (what a register-on-entry strategy would pay per call)
a = np.empty(n, dtype=np.complex128); b = np.empty(n, dtype=np.float64)
t0 = time.perf_counter(); _lib.check(lib.ila_host_register(_lib.ptr(a), a.nbytes)); _lib.check(lib.ila_host_register(_lib.ptr(b), b.nbytes)); t1 = time.perf_counter()
_lib.check(lib.ila_host_unregister(_lib.ptr(a))); _lib.check(lib.ila_host_unregister(_lib.ptr(b))); t2 = time.perf_counter()
e2e["cudaHostRegister 16.8+8.4 MB (ms)"] = [(t1 - t0) * 1e3, (t2 - t1) * 1e3]
# raw copy rates
h = torch.empty(16 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty(16 << 20, dtype=torch.uint8, device=dev)
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); e2e["H2D pinned 16 MiB GB/s"] = 20 * 16.777216e-3 / (time.perf_counter() - t0)
t0 = time.perf_counter()
for _ in range(20): h.copy_(d, non_blocking=True)
torch.cuda.synchronize(); e2e["D2H pinned 16 MiB GB/s"] = 20 * 16.777216e-3 / (time.perf_counter() - t0)
hp = torch.empty(16 << 20, dtype=torch.uint8)
t0 = time.perf_counter()
for _ in range(10): d.copy_(hp)
torch.cuda.synchronize(); e2e["H2D pageable 16 MiB GB/s"] = 10 * 16.777216e-3 / (time.perf_counter() - t0)
src = np.ones(16 << 20, dtype=np.uint8); dst = np.empty_like(src)
t0 = time.perf_counter()
for _ in range(10): np.copyto(dst, src)
e2e["host memcpy 1 thread GB/s"] = 10 * 16.777216e-3 / (time.perf_counter() - t0)
e2e["host_cpus"] = os.cpu_count()
for k, v in e2e.items(): print(f"{k:55s} {v}")
out["e2e_ms"] = e2e
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_k3_e2e_probe.json"), "w"), indent=1)
