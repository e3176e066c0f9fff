"""Where the K3 host call's time goes: the same byte movements issued by hand with torch streams and timed with CUDA events
(H2D alone, H2D + concurrent D2H, H2D chunks + kernels + D2H in the library's order), next to the library call."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ila_b200
from ila_b200 import _lib
lib = ila_b200.lib(); dev = torch.device("cuda", 0)
col = np.ascontiguousarray(np.array([2j, 0.0, 0.0, 1.0, 0.7]), dtype=np.complex128)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
x = np.linspace(-10.0, 7.0, 1024); y = np.linspace(-7.0, 8.0, 1024)
lam = (x[None, :] + 1j * y[:, None]).reshape(-1); n = lam.size
h_lam = torch.from_numpy(lam).pin_memory(); h_L = torch.empty(n, dtype=torch.float64).pin_memory(); h_s = torch.empty(n, dtype=torch.uint8).pin_memory()
d_lam = torch.empty(n, dtype=torch.complex128, device=dev); d_L = torch.empty(n, dtype=torch.float64, device=dev); d_s = torch.empty(n, dtype=torch.uint8, device=dev)
s_in, s_k, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)

def chunks(minc=1 << 16, div=3):
    c0 = 0; out = []
    while c0 < n:
        cc = max(((n - c0) // div + 1023) & ~1023, minc)
        if c0 + cc > n or n - (c0 + cc) < minc // 2: cc = n - c0
        out.append((c0, cc)); c0 += cc
    return out

def manual(with_kernel=True, with_d2h=True, ch=None):
    ch = ch or chunks()
    for (c0, cc) in ch:
        with torch.cuda.stream(s_in):
            d_lam[c0:c0 + cc].copy_(h_lam[c0:c0 + cc], non_blocking=True)
            e = torch.cuda.Event(); e.record(s_in)
        if with_kernel:
            s_k.wait_event(e)
            _lib.check(lib.ila_ql_toeplitz_l11re_c64_dev(3, 1, _lib.ptr(col), ctypes.c_void_p(d_lam.data_ptr() + 16 * c0), cc, 1,
                                                         ctypes.c_void_p(d_L.data_ptr() + 8 * c0), ctypes.c_void_p(d_s.data_ptr() + c0), ctypes.c_void_p(s_k.cuda_stream)))
            e = torch.cuda.Event(); e.record(s_k)
        if with_d2h:
            s_out.wait_event(e)
            with torch.cuda.stream(s_out):
                h_L[c0:c0 + cc].copy_(d_L[c0:c0 + cc], non_blocking=True)
                h_s[c0:c0 + cc].copy_(d_s[c0:c0 + cc], non_blocking=True)
    torch.cuda.synchronize()

def libcall():
    _lib.check(lib.ila_ql_toeplitz_l11re_c64(3, 1, _lib.ptr(col), _lib.ptr(h_lam), n, 1, _lib.ptr(h_L), _lib.ptr(h_s), 1))

def med(f, reps=15):
    for _ in range(3): f()
    ts = []
    for _ in range(reps):
        flush.fill_(1); torch.cuda.synchronize(); t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts)), float(np.min(ts))
one = [(0, n)]
print("library call                              ", med(libcall))
print("manual: H2D only, one copy                ", med(lambda: manual(False, False, one)))
print("manual: H2D only, library's chunks        ", med(lambda: manual(False, False)))
print("manual: H2D chunks + kernels              ", med(lambda: manual(True, False)))
print("manual: H2D chunks + kernels + D2H (torch)", med(lambda: manual(True, True)))
for minc, div in ((1 << 17, 2), (1 << 18, 2), (1 << 16, 2), (1 << 15, 4)):
    print(f"manual full, chunks min {minc} div {div}: ", med(lambda: manual(True, True, chunks(minc, div))), len(chunks(minc, div)))
print("chunks:", chunks())

# ---- event timeline of one manual pipelined call
def timeline(ch):
    ev0 = torch.cuda.Event(enable_timing=True)
    marks = []
    flush.fill_(1); torch.cuda.synchronize()
    for s in (s_in, s_k, s_out): s.wait_stream(torch.cuda.current_stream())
    ev0.record(s_in)
    for (c0, cc) in ch:
        with torch.cuda.stream(s_in):
            d_lam[c0:c0 + cc].copy_(h_lam[c0:c0 + cc], non_blocking=True)
            e1 = torch.cuda.Event(enable_timing=True); e1.record(s_in)
        s_k.wait_event(e1)
        _lib.check(lib.ila_ql_toeplitz_l11re_c64_dev(3, 1, _lib.ptr(col), ctypes.c_void_p(d_lam.data_ptr() + 16 * c0), cc, 1,
                                                     ctypes.c_void_p(d_L.data_ptr() + 8 * c0), ctypes.c_void_p(d_s.data_ptr() + c0), ctypes.c_void_p(s_k.cuda_stream)))
        e2 = torch.cuda.Event(enable_timing=True); e2.record(s_k)
        s_out.wait_event(e2)
        with torch.cuda.stream(s_out):
            h_L[c0:c0 + cc].copy_(d_L[c0:c0 + cc], non_blocking=True)
            e3 = torch.cuda.Event(enable_timing=True); e3.record(s_out)
            h_s[c0:c0 + cc].copy_(d_s[c0:c0 + cc], non_blocking=True)
            e4 = torch.cuda.Event(enable_timing=True); e4.record(s_out)
        marks.append((cc, e1, e2, e3, e4))
    torch.cuda.synchronize()
    for cc, e1, e2, e3, e4 in marks:
        print(f"  chunk {cc:7d}: H2D done {ev0.elapsed_time(e1)*1e3:7.1f} us  kernel done {ev0.elapsed_time(e2)*1e3:7.1f}  D2H L done {ev0.elapsed_time(e3)*1e3:7.1f}  D2H status done {ev0.elapsed_time(e4)*1e3:7.1f}")
for _ in range(2): manual()
print("timeline (library's chunk schedule):"); timeline(chunks())
print("timeline (4 chunks):"); timeline(chunks(1 << 17, 2))
