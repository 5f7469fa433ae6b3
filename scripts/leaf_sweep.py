"""Sweep of the leaf-table parameters (environment knobs read at kmcb200_create) on one resident bin: per-config step time and
stage intervals.  Usage (GPU box): python scripts/leaf_sweep.py [n_rec_log2] [k]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kmc_b200
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from kmc_testlib import fast_bin
if os.environ.get("KMCB200_LIB"):
    kmc_b200.LIB_PATH = os.path.abspath(os.environ["KMCB200_LIB"])

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
K = int(sys.argv[2]) if len(sys.argv) > 2 else 31
P = {31: 7, 55: 7, 28: 8}.get(K, K % 4 if K % 4 else 4)
n_rec = 1 << lg
dev = torch.device("cuda", 0)
bins = [fast_bin(1000 + j, K, n_rec) for j in range(2)]
d_bins = []
for hb in bins:
    t = torch.zeros(hb.size + 64, dtype=torch.uint8, device=dev)
    t[:hb.size] = torch.from_numpy(hb.data).to(dev)
    d_bins.append(t)
configs = [(10, 100, 0), (9, 150, 0)]
if len(sys.argv) > 3:
    configs = [tuple(int(x) for x in c.split(":")) for c in sys.argv[3:]]
for sb, pct, var in configs:
    os.environ["KMCB200_LEAF_SLOT_BITS"] = str(sb)
    os.environ["KMCB200_LEAF_ROUND_PCT"] = str(pct)
    os.environ["KMCB200_LEAF_VARIANT"] = str(var)
    ctx = kmc_b200.Stage2Context(kmc_b200.Stage2Params(K, True, 2, 10 ** 9, 255, P), device=0, n_slots=1)
    cap = ctx.out_capacity(n_rec) + 64
    d_out = torch.zeros(cap, dtype=torch.uint8, device=dev)
    d_lut = torch.zeros(ctx.lut_entries, dtype=torch.int64, device=dev)
    d_res = torch.zeros(8, dtype=torch.int64, device=dev)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)

    def step(i):
        hb = bins[i % 2]
        ctx.dev_process_bin(0, d_bins[i % 2].data_ptr(), hb.size, n_rec, hb.pack_bytes, d_out.data_ptr(), cap, d_lut.data_ptr(), d_res.data_ptr(), st.cuda_stream)

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        step(3 + i)
    e1.record()
    torch.cuda.synchronize()
    r = d_res.cpu().numpy()
    s = ctx.stage_times(0)
    iv = dict(zip(s.get("pass_names") or [], s["pass_ms"]))
    print("slot_bits=%d round_pct=%d variant=%d: %.3f ms/step  expand %.3f  leaves %.3f  fallback=%d emitted=%d stats=%s  %s" % (
        sb, pct, var, e0.elapsed_time(e1) / 10, s["expand_ms"], iv.get("leaf_count", 0.0), int(r[7]), int(r[4]), list(map(int, r[:3])),
        " ".join("%s=%.3f" % (k, v) for k, v in iv.items())), flush=True)
    ctx.close()
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
