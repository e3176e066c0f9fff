"""Host<->device copy bandwidth with pinned buffers (context for the e2e numbers; not a test). Run under gpurun."""
import time, torch
dev = torch.device("cuda", 0)
for mb in (2, 8, 16, 64):
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev); d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(kind, reps=20):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
    for kind in ("h2d", "d2h", "both"):
        run(kind, 3); t = run(kind)
        tot = n * (2 if kind == "both" else 1)
        print(f"{mb:3d} MiB {kind:5s}: {t*1e3:.3f} ms  {tot/t/1e9:.1f} GB/s total")
