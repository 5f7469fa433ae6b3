"""One large bin on 1 GPU vs split over all visible GPUs (kmcb200_process_bin_multi).  usage: multi_gpu_bin.py [LG] [k]
Run under `gpurun --gpus N`.  Prints host-to-host times (pageable host buffers: the H2D of the bin and the D2H of the records are inside)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
import kmc_b200, bench

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
K = int(sys.argv[2]) if len(sys.argv) > 2 else 31
n_dev = torch.cuda.device_count()
with ThreadPoolExecutor(32) as ex:
    sk = bench.gen_bin(4712, K, 1 << lg, ex)
sp = kmc_b200.Stage2Params(K, True, 2, 10 ** 9, 255, 7)
ctxs = [kmc_b200.Stage2Context(sp, device=g, n_slots=1) for g in range(n_dev)]
pin = torch.from_numpy(sk.data.copy()).pin_memory()
sk.data = pin.numpy()
cap = ctxs[0].out_capacity(sk.n_rec) + 64
out = torch.empty(cap, dtype=torch.uint8).pin_memory().numpy()
lut = np.empty(ctxs[0].lut_entries, dtype=np.uint64)


def timed(f, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t0)
    return r, min(ts)


one, t1 = timed(lambda: ctxs[0].process_bin(sk, out=out, lut=lut))
ref_payload, ref_lut, ref_stats = one.payload.tobytes(), one.lut.copy(), one.stats
print("bin 2^%d k-mers (k=%d, %.2f GB): 1 GPU %.1f ms = %.3g k-mers/s" % (lg, K, sk.size / 1e9, t1 * 1e3, sk.n_rec / t1), flush=True)
for n in sorted({2, 4, n_dev}):
    if n > n_dev or n < 2:
        continue
    r, t = timed(lambda: kmc_b200.Stage2Context.process_bin_multi(ctxs[:n], sk, out=out, lut=lut))
    same = r.stats == ref_stats and np.array_equal(r.lut, ref_lut) and r.payload.tobytes() == ref_payload
    print("  split over %d GPUs: %.1f ms = %.3g k-mers/s (%.2fx), byte-identical: %s" % (n, t * 1e3, sk.n_rec / t, t1 / t, same), flush=True)
    assert same
