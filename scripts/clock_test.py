import sys, os, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ila_b200
from ila_b200 import _lib
dev = torch.device("cuda", 0)
def smclk():
    return subprocess.run(["nvidia-smi","--query-gpu=clocks.sm,clocks.max.sm,power.draw","--format=csv,noheader"],capture_output=True,text=True).stdout.strip()
def ev_time(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return np.median(ts)
sigma = 0.01 + (10 - 0.01) * np.arange(2048) / 2047
p4, q4 = ila_b200.pentadiagonal_operator(sigma)
plan4 = ila_b200.QrBandedPlan(0, 2, p4, q4, b=[1.0], ncap=8000, kind="chol")
col = np.array([2j, 0, 0, 1, 0.7])
x = np.linspace(-10, 7, 1024); y = np.linspace(-7, 8, 1024)
lam = torch.from_numpy((x[None, :] + 1j * y[:, None]).reshape(-1)).to(dev)
L = torch.empty_like(lam); st = torch.empty(lam.numel(), dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
f = lambda: ila_b200.ql_toeplitz_l11_dev(col, lam, L, st, stream=s)
print("idle clk", smclk())
print("cold: chol fwd %.0f us, K3 %.1f us" % (ev_time(plan4.run_factor)*1e3, ev_time(f)*1e3), smclk())
a = torch.randn(8192, 8192, device=dev)
t0 = time.time()
while time.time() - t0 < 1.0:
    b = a @ a
torch.cuda.synchronize()
print("after burn clk", smclk())
print("hot: chol fwd %.0f us, K3 %.1f us" % (ev_time(plan4.run_factor)*1e3, ev_time(f)*1e3), smclk())
# many back-to-back launches
for _ in range(300): plan4.run_factor()
torch.cuda.synchronize()
print("after 300 launches: chol fwd %.0f us" % (ev_time(plan4.run_factor)*1e3), smclk())
lat = np.zeros(4); _lib.check(_lib.lib().ila_fp64_latency_probe(_lib.ptr(lat))); print("latency probe", lat)
