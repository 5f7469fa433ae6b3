"""Development probe: per-stage CUDA-event times of one HBM-resident bin.  usage: probe_bin.py [n_rec] [k] [iterations] [genome_div]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, kmc_b200
from concurrent.futures import ThreadPoolExecutor
import bench

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 117440512
k = int(sys.argv[2]) if len(sys.argv) > 2 else 31
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
ctx = kmc_b200.Stage2Context(kmc_b200.Stage2Params(k, True, 2, 10 ** 9, 255, 7), device=0, n_slots=1)
_cache = "/dev/shm/probe_bin_%d_%d.npz" % (n_rec, k)          # (several runs of one gpurun call share the generated bin)
if os.path.exists(_cache):
    from kmc_testlib import Bin
    z = np.load(_cache)
    hb = Bin(data=z["data"], n_rec=n_rec, n_super_kmers=int(z["nsk"]), pack_bytes=z["pack_bytes"], pack_recs=z["pack_recs"], k=k)
else:
    with ThreadPoolExecutor(32) as ex:
        hb = bench.gen_bin(4004, k, n_rec, ex)
    try:
        np.savez(_cache, data=hb.data, nsk=hb.n_super_kmers, pack_bytes=hb.pack_bytes, pack_recs=hb.pack_recs)
    except OSError:
        pass
cap = ctx.out_capacity(n_rec) + 64
dev = torch.device("cuda", 0)
d_bin = torch.zeros(hb.size + 64, dtype=torch.uint8, device=dev); d_bin[:hb.size] = torch.from_numpy(hb.data).to(dev)
d_out = torch.zeros(cap, dtype=torch.uint8, device=dev)
d_lut = torch.zeros(ctx.lut_entries, dtype=torch.int64, device=dev)
d_res = torch.zeros(8, dtype=torch.int64, device=dev)
st = torch.cuda.Stream(device=dev)
acc = {}
for it in range(iters):
    l0 = ctx.kernel_launches()
    ctx.dev_process_bin(0, d_bin.data_ptr(), hb.size, n_rec, hb.pack_bytes, d_out.data_ptr(), cap, d_lut.data_ptr(), d_res.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    t = ctx.stage_times(0)
    tot = t["expand_ms"] + t["sort_ms"] + t["count_ms"]
    if it:
        for nm, x in [("expand", t["expand_ms"])] + list(zip(t["pass_names"], t["pass_ms"])) + [("total", tot)]:
            acc.setdefault(nm, []).append(x)
    print("it%d total %.3f ms (%.2f G k-mers/s) launches %d | expand=%.3f %s" % (it, tot, n_rec / tot / 1e6, ctx.kernel_launches() - l0, t["expand_ms"],
          " ".join("%s=%.3f" % (nm, x) for nm, x in zip(t["pass_names"], t["pass_ms"]))))
print("MEAN " + " ".join("%s=%.3f" % (nm, sum(v) / len(v)) for nm, v in acc.items()))
print("result", d_res.cpu().numpy())
