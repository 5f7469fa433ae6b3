"""Development probe: does the NUMA node of the pinned host buffers limit the host<->device copies of the end-to-end path?"""
import glob
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


nodes = {}
for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    nodes[int(d.rsplit("node", 1)[1])] = cpulist(open(d + "/cpulist").read())
print("numa nodes:", {k: "%d cpus (%d..%d)" % (len(v), v[0], v[-1]) for k, v in nodes.items()})
p = torch.cuda.get_device_properties(0)
bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
try:
    gnode = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
except Exception as e:
    gnode = None
    print("numa_node lookup failed:", e)
print("gpu0 pci", bus, "numa node", gnode, " current affinity: %d cpus" % len(os.sched_getaffinity(0)))

n = 66 * 1024 * 1024
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
all_cpus = sorted(os.sched_getaffinity(0))


def bw(h, h2, label):
    for name, fn in [("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))]:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        print("  %-22s %s %.2f ms, %.1f GB/s" % (label, name, dt * 1e3, n / dt / 1e9))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print("  %-22s both directions at once: %.2f ms per pair, %.1f GB/s each way" % (label, dt * 1e3, n / dt / 1e9))


for label, cpus in [("default affinity", all_cpus)] + [("bound to node %d" % k, v) for k, v in nodes.items()]:
    cp = [c for c in cpus if c in all_cpus]
    if not cp:
        continue
    os.sched_setaffinity(0, cp)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(1); h2.fill_(2)
    bw(h, h2, label)
    del h, h2
os.sched_setaffinity(0, all_cpus)

# ---- does splitting one H2D transfer over several streams (copy engines) help?
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(3)
streams = [torch.cuda.Stream() for _ in range(8)]
for parts in (1, 2, 4, 8):
    step = n // parts
    def go():
        for i in range(parts):
            with torch.cuda.stream(streams[i]):
                d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        go()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print("  H2D split over %d streams: %.2f ms, %.1f GB/s" % (parts, dt * 1e3, n / dt / 1e9))
# ---- a kernel reading the pinned buffer directly (zero-copy) instead of the copy engine
import ctypes
hp = torch.empty(n, dtype=torch.uint8).pin_memory()
hp.fill_(5)
rt = ctypes.CDLL("libcudart.so.12")
dptr = ctypes.c_void_p()
rc = rt.cudaHostGetDevicePointer(ctypes.byref(dptr), ctypes.c_void_p(hp.data_ptr()), 0)
print("  cudaHostGetDevicePointer rc", rc, "same pointer" if dptr.value == hp.data_ptr() else "different pointer")
if rc == 0:
    import numpy as np
    # wrap the device-visible alias as a CUDA tensor and let an elementwise kernel do the reading
    class _Holder:
        pass
    hold = _Holder()
    hold.__cuda_array_interface__ = {"shape": (n // 16, 2), "typestr": "<i8", "data": (dptr.value, False), "version": 3}
    try:
        src = torch.as_tensor(hold, device="cuda")
        dst = torch.empty_like(src)
        for _ in range(3):
            dst.copy_(src)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            dst.copy_(src)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        print("  zero-copy kernel read of pinned memory: %.2f ms, %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
    except Exception as e:
        print("  zero-copy probe failed:", repr(e)[:200])
