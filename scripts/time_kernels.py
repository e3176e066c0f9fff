"""Kernel-only timings with CUDA events (not a test). Run under gpurun."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ila_b200
from ila_b200 import _lib
from scipy.special import jv

dev = torch.device("cuda", 0)
lat = np.zeros(4); _lib.check(_lib.lib().ila_fp64_latency_probe(_lib.ptr(lat)))
print("latency cycles: dfma %.1f, dmul+dadd %.1f, sqrt+add %.1f, div+add %.1f" % tuple(lat))
print("fp64 peak", ila_b200.fp64_peak_tflops())

def ev_time(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return np.median(ts), np.min(ts)

# K3
col = np.array([2j, 0, 0, 1, 0.7])
x = np.linspace(-10, 7, 1024); y = np.linspace(-7, 8, 1024)
lam = torch.from_numpy((x[None, :] + 1j * y[:, None]).reshape(-1)).to(dev)
L = torch.empty_like(lam); st = torch.empty(lam.numel(), dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
f = lambda: ila_b200.ql_toeplitz_l11_dev(col, lam, L, st, stream=s)
med, mn = ev_time(f, reps=20)
print(f"K3 1M points: median {med*1e3:.1f} us min {mn*1e3:.1f} us -> {lam.numel()/med/1e-3/1e9:.2f} Gpts/s")

# QR
for name, zs in (("64 z in [1e3,1e4]", np.linspace(1e3, 1e4, 64)), ("4096 z in [1e3,1e5] (C2)", 1e3 + (1e5 - 1e3) * np.arange(4096) / 4095),
                 ("32768 z=2000", np.full(32768, 2000.0))):
    p, q, r = ila_b200.bessel_operator(zs)
    caps = ila_b200.bessel_ncap(zs)
    plan = ila_b200.QrBandedPlan(1, 1, p, q, r, b=jv(1, zs)[:, None], ncap_per=caps)
    if "C2" in name or "64 z" in name:
        _lib.check(_lib.lib().ila_qr_set_arith_mode(1)); plan.run(); torch.cuda.synchronize()
        x_ref = plan.d_x.clone(); nc_ref = plan.d_ncols.clone(); work_ref = plan.d_work.clone()
        tm1, _ = ev_time(plan.run, reps=2, warm=0)
        _lib.check(_lib.lib().ila_qr_set_arith_mode(0)); plan.run(); torch.cuda.synchronize()
        tot = int(plan.d_xoff[-1].item())
        print("  mode0 vs mode1: ncols equal", bool(torch.equal(nc_ref, plan.d_ncols)), "x bit-equal", bool(torch.equal(x_ref[:tot].view(torch.int64), plan.d_x[:tot].view(torch.int64))), "records bit-equal (swept part) checked via x; intrinsics-mode time %.2f ms" % tm1)
        del x_ref, work_ref
    mf, _ = ev_time(plan.run_factor, reps=5, warm=1)
    mb, _ = ev_time(plan.run_backsolve, reps=5, warm=1)
    mt, _ = ev_time(plan.run, reps=5, warm=1)
    ncols = plan.d_ncols.cpu().numpy(); nsw = plan.d_nswept.cpu().numpy()
    print(f"QR {name}: fwd {mf:.2f} ms bwd {mb:.2f} ms total {mt:.2f} ms; sum ncols {ncols.sum()} swept {nsw.sum()} max {nsw.max()}"
          f" -> {nsw.sum()/mt/1e-3/1e9:.3f} Gcol/s; fwd cycles/col(longest) {mf*1e-3*1.9e9/nsw.max():.0f} bwd {mb*1e-3*1.9e9/nsw.max():.0f}; GB/s {nsw.sum()*72/mt/1e-3/1e9:.0f}")

# K3 shifts-per-thread sweep
for spt in (1, 2, 4, 8, 16, 0):
    _lib.check(_lib.lib().ila_ql_set_shifts_per_thread(spt))
    med, mn = ev_time(f, reps=20)
    print(f"K3 spt={spt}: median {med*1e3:.1f} us min {mn*1e3:.1f} us -> {lam.numel()/med/1e-3/1e9:.2f} Gpts/s")

# C4 Cholesky phases
sigma = 0.01 + (10 - 0.01) * np.arange(2048) / 2047
p4, q4 = ila_b200.pentadiagonal_operator(sigma)
for name, pp, qq in (("C4 2048 shifts", p4, q4), ("C4-like 32768 shifts", np.tile(p4, (16, 1)), np.tile(q4, (16, 1)))):
    plan4 = ila_b200.QrBandedPlan(0, 2, pp, qq, b=[1.0], ncap=8000, kind="chol")
    mf, _ = ev_time(plan4.run_factor, reps=10, warm=3)
    mb, _ = ev_time(plan4.run_backsolve, reps=10, warm=3)
    nsw = plan4.d_nswept.cpu().numpy()
    print(f"CHOL {name}: fwd {mf*1e3:.0f} us bwd {mb*1e3:.0f} us; swept {nsw.sum()} max {nsw.max()}; fwd cycles/col(longest) {mf*1e-3*1.965e9/nsw.max():.0f} bwd {mb*1e-3*1.965e9/nsw.max():.0f}; {nsw.sum()/(mf+mb)/1e-3/1e9:.2f} Gcol/s")
