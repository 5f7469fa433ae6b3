"""PCIe ceiling of the box: H2D alone, D2H alone, both directions at once, H2D split over two streams (pinned buffers, 256 MiB transfers)."""
import os, sys, time, torch
n = 256 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=8):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
def h2d_split():
    half = n // 2
    with torch.cuda.stream(s1): d1[:half].copy_(h1[:half], non_blocking=True)
    with torch.cuda.stream(s2): d1[half:].copy_(h1[half:], non_blocking=True)
print("H2D alone        %.1f GB/s" % (n / t(h2d) / 1e9))
print("D2H alone        %.1f GB/s" % (n / t(d2h) / 1e9))
dt = t(both); print("H2D + D2H        %.1f GB/s each direction (%.1f total)" % (n / dt / 1e9, 2 * n / dt / 1e9))
print("H2D two streams  %.1f GB/s" % (n / t(h2d_split) / 1e9))
try:
    import subprocess
    print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current", "--format=csv"], capture_output=True, text=True).stdout)
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
except Exception as e: print(e)
