import os, sys
"""Development probe: where does the host-buffer path spend its time?"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, kmc_b200

n_rec = 1 << 26
ctx = kmc_b200.Stage2Context(kmc_b200.Stage2Params(31, True, 2, 10 ** 9, 255, 7), device=0, n_slots=2)
hb = fast_bin(1, 31, n_rec)
cap = ctx.out_capacity(n_rec) + 64
pin_bin = torch.from_numpy(hb.data.copy()).pin_memory()
pin_out = torch.zeros(cap, dtype=torch.uint8).pin_memory()
pin_lut = torch.zeros(ctx.lut_entries, dtype=torch.int64).pin_memory()
for it in range(4):
    t0 = time.perf_counter()
    ctx.submit_bin(0, pin_bin.data_ptr(), hb.size, n_rec, hb.pack_bytes, pin_out.data_ptr(), cap, pin_lut.data_ptr())
    t1 = time.perf_counter()
    nb, st = ctx.wait_bin(0)
    t2 = time.perf_counter()
    print("single slot: submit %.2f ms wait %.2f ms  bytes=%d  stages=%s" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), nb, ctx.stage_times(0)))
pin_out2 = torch.zeros(cap, dtype=torch.uint8).pin_memory()
pin_lut2 = torch.zeros(ctx.lut_entries, dtype=torch.int64).pin_memory()
outs = [pin_out, pin_out2]; luts = [pin_lut, pin_lut2]
for it in range(3):
    t0 = time.perf_counter()
    ctx.submit_bin(0, pin_bin.data_ptr(), hb.size, n_rec, hb.pack_bytes, outs[0].data_ptr(), cap, luts[0].data_ptr())
    t1 = time.perf_counter()
    ctx.submit_bin(1, pin_bin.data_ptr(), hb.size, n_rec, hb.pack_bytes, outs[1].data_ptr(), cap, luts[1].data_ptr())
    t2 = time.perf_counter()
    ctx.wait_bin(0)
    t3 = time.perf_counter()
    ctx.wait_bin(1)
    t4 = time.perf_counter()
    print("two slots: submit0 %.2f submit1 %.2f wait0 %.2f wait1 %.2f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3)))
    print("   stages slot0", ctx.stage_times(0), "\n   stages slot1", ctx.stage_times(1))
