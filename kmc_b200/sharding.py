"""Bin → GPU sharding for a multi-GPU stage 2 (one process per GPU, no collective on the data path).

Bins never share a k-mer (a k-mer's bin is a function of its minimizer signature, kmc_core/splitter.cpp:611-637), so ranks simply
own disjoint sets of bins.  The order mirrors the reference: bins are taken in descending memory requirement
(CBinDesc::get_sorted_req_sizes, kmc_core/queues.h:499-558) and each goes to the least-loaded rank (LPT), which is what a shared
CBinQueue drained by N sorter objects converges to.  The only cross-bin state, the four statistics summed by the completer
(kb_completer.cpp:206-209), is reduced with one tiny all_reduce.
"""
from typing import List, Sequence


def assign_bins(bin_costs: Sequence[int], world_size: int) -> List[List[int]]:
    """Returns, for every rank, the list of bin ids it processes (in processing order)."""
    order = sorted(range(len(bin_costs)), key=lambda b: (-int(bin_costs[b]), b))
    load = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for b in order:
        r = min(range(world_size), key=lambda i: (load[i], i))
        out[r].append(b)
        load[r] += int(bin_costs[b])
    return out


def reduce_stats(stats, group=None):
    """Sum (n_unique, n_cutoff_min, n_cutoff_max, n_total) over ranks; int64 is exact for KMC's counters."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([int(x) for x in stats], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tuple(int(x) for x in t.cpu())
