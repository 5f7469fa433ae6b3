"""Large-bin checks on the GPU box (too big for the oracle): size-independent properties, and the oversized-bin path (key blocks)
against the one-shot path on the same bin, byte for byte.  Usage: python scripts/big_bin_check.py [lg_n=27] [lg_oversized=29]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import kmc_b200
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from kmc_testlib import fast_bin

K, P = 31, 7
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 27
lg_big = int(sys.argv[2]) if len(sys.argv) > 2 else 29


def properties(ctx, r, n_rec):
    assert r.n_total == n_rec, (r.n_total, n_rec)
    n_emit = r.payload.size // ctx.out_rec_bytes
    assert int(r.lut.sum()) == n_emit == r.n_unique - r.n_cutoff_min - r.n_cutoff_max
    rec = r.payload.reshape(n_emit, ctx.out_rec_bytes)
    assert rec[:, -1].min() >= 2
    prefix = np.repeat(np.arange(ctx.lut_entries, dtype=np.uint64), r.lut.astype(np.int64))
    suf = np.zeros(n_emit, dtype=np.uint64)
    for j in range(6):
        suf = (suf << np.uint64(8)) | rec[:, j].astype(np.uint64)
    full = (prefix << np.uint64(48)) | suf
    assert np.all(full[1:] > full[:-1]), "emitted k-mers are not strictly increasing"
    return n_emit


def run(n_rec, seed, env):
    for k_, v in env.items():
        os.environ[k_] = str(v)
    ctx = kmc_b200.Stage2Context(kmc_b200.Stage2Params(K, True, 2, 10 ** 9, 255, P), device=0, n_slots=1)
    for k_ in env:
        del os.environ[k_]
    sk = fast_bin(seed, K, n_rec)
    t0 = time.perf_counter()
    r = ctx.process_bin(sk)
    dt = time.perf_counter() - t0
    n_emit = properties(ctx, r, n_rec)
    ctx.close()
    return r, n_emit, dt


a, na, ta = run(1 << lg, 4711, {})
b, nb, tb = run(1 << lg, 4711, {"KMCB200_MAX_BLOCK_RECORDS": 1 << (lg - 3), "KMCB200_MAX_CHUNK_BYTES": 1 << 24})
assert na == nb and a.stats == b.stats and np.array_equal(a.lut, b.lut) and a.payload.tobytes() == b.payload.tobytes(), "key blocks differ from the one-shot path"
print("2^%d k-mers: one shot %.0f ms, key blocks (<= 2^%d k-mers, 16 MiB chunks) %.0f ms: %d records, byte-identical" % (lg, ta * 1e3, lg - 3, tb * 1e3, na))
c, nc, tc = run(1 << lg_big, 4712, {})
print("2^%d k-mers (oversized by default): %.0f ms, %d records, properties ok, stats %s" % (lg_big, tc * 1e3, nc, c.stats))
