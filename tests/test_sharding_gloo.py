"""CPU, world_size 2 (gloo): the multi-GPU host logic - LPT bin sharding + the all_reduce of the four statistics.
The per-bin work is done by the oracle here (this is a test of the sharding, not of the kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kmc_b200.sharding import assign_bins, reduce_stats


def test_assign_bins_lpt():
    costs = [5, 100, 7, 30, 30, 1, 0, 64]
    parts = assign_bins(costs, 3)
    assert sorted(b for p in parts for b in p) == list(range(len(costs)))
    loads = [sum(costs[b] for b in p) for p in parts]
    assert max(loads) == 100 and min(loads) >= 60           # LPT keeps ranks balanced
    assert parts[0][0] == 1                                  # largest bin first (get_sorted_req_sizes order)
    assert assign_bins(costs, 1) == [sorted(range(len(costs)), key=lambda b: (-costs[b], b))]


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__)))
    from kmc_testlib import Oracle, Params, synth_bin
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prm = Params(k=31, cutoff_min=2, lut_prefix_len=7)
    sizes = [900, 50, 400, 0, 1200, 300, 10]
    bins = [synth_bin(70 + i, 31, n, genome_len=max(n, 300)) for i, n in enumerate(sizes)]
    mine = assign_bins([b.n_rec for b in bins], world)[rank]
    O = Oracle()
    local = np.zeros(4, dtype=np.int64)
    for b in mine:
        local += np.array(O.process_bin(bins[b], prm).stats, dtype=np.int64)
    total = reduce_stats(local)
    if rank == 0:
        ref = np.zeros(4, dtype=np.int64)
        for b in bins:
            ref += np.array(O.process_bin(b, prm).stats, dtype=np.int64)
        q.put((total, tuple(int(x) for x in ref), mine))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, ref, mine = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == ref
    assert len(mine) >= 1
