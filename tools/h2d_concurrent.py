"""Per-GPU pinned host->device copy rate while N processes copy at once (the e2e ceiling of a multi-GPU box).
usage: python tools/h2d_concurrent.py <device> <n_concurrent>   (started N times in parallel by tools/run_h2d_scaling.sh)"""
import sys, time
import torch
dev, n = int(sys.argv[1]), int(sys.argv[2])
torch.cuda.set_device(dev)
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(1 << 30, dtype=torch.uint8, device=f"cuda:{dev}")
d.copy_(h, non_blocking=True); torch.cuda.synchronize()
time.sleep(max(0.0, 2.0 - (time.time() % 2.0)))       # crude alignment of the N processes on a 2 s boundary
t0 = time.perf_counter()
for _ in range(12):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"concurrent={n} device={dev}: {12 * (1 << 30) / dt / 1e9:.1f} GB/s", flush=True)
