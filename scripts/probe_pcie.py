import os, sys
import torch, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = 66 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in [("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))]:
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(name, "%.2f ms, %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
import kmc_b200, numpy as np
n_rec = 1 << 26
ctx = kmc_b200.Stage2Context(kmc_b200.Stage2Params(31, True, 2, 10 ** 9, 255, 7), device=0, n_slots=3)
hbs = [fast_bin(1 + j, 31, n_rec) for j in range(2)]
cap = ctx.out_capacity(n_rec) + 64
pin_bins = [torch.from_numpy(hb.data.copy()).pin_memory() for hb in hbs]
for nslots in (1, 2, 3):
    outs = [torch.zeros(cap, dtype=torch.uint8).pin_memory() for _ in range(nslots)]
    luts = [torch.zeros(ctx.lut_entries, dtype=torch.int64).pin_memory() for _ in range(nslots)]
    def run(steps):
        for i in range(steps):
            s = i % nslots
            if i >= nslots: ctx.wait_bin(s)
            hb = hbs[i % 2]
            ctx.submit_bin(s, pin_bins[i % 2].data_ptr(), hb.size, n_rec, hb.pack_bytes, outs[s].data_ptr(), cap, luts[s].data_ptr())
        for i in range(max(steps - nslots, 0), steps): ctx.wait_bin(i % nslots)
    run(6); torch.cuda.synchronize(); t0 = time.perf_counter(); run(30); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 30
    print("slots=%d: %.2f ms per bin, %.2f G k-mers/s" % (nslots, dt * 1e3, n_rec / dt / 1e9))
