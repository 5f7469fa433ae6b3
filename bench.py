#!/usr/bin/env python
"""bench.py — stage-2 k-mers/s (k=31) of the B200 path, next to the reference's CPU stage 2.

A "step" = one bin of N_REC k-mers (BASELINE.json configs[1]: k=31, one bin of 2^26 packed 64-bit k-mers) taken
through the whole hot path Expand -> Sort -> Compact (kmc_core/kb_sorter.h:210-237) on every GPU (weak scaling: one
bin per GPU per step, bins are independent, no collective on the data path).

  value     : k-mers/s with the bin's super-k-mer bytes already resident in HBM, device-timed (CUDA events, max over ranks)
  e2e       : the same through the host-buffer C ABI (kmcb200_submit_bin / kmcb200_wait_bin, three slots), pinned host
              buffers, H2D of the bin and D2H of the database records + LUT + counters inside the timed region
  roofline  : the radix pass (dominant kernel): 2*N*W algorithmic bytes / CUDA-event duration of the pass launches
  cpu_baseline / --impl reference : the UNMODIFIED reference classes (oracle/_ref, CKmerBinSorter<1> + RADULS) on the host cores

Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n-rec N_REC]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K = 31
LUT_P = 7
CUTOFF_MIN, CUTOFF_MAX, COUNTER_MAX = 2, 10 ** 9, 255
METRIC = "stage-2 k-mers/s (k=31)"
UNIT = "k-mers/s"
REC_BYTES = 8
KEY_BYTES = 8
E2E_SLOTS = int(os.environ.get("KMCB200_E2E_SLOTS", "3"))          # bins in flight through submit_bin / wait_bin: the D2H of bin i-3 and the H2D of bin i overlap the kernels of bins i-2, i-1


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_config(n_rec, n_gpus):
    return {
        "workload": "k=31 canonical, one bin of %d k-mers per GPU per step (BASELINE configs[1]), ci=2 cx=1e9 cs=255 p=7" % n_rec,
        "n_rec_per_bin": n_rec, "record_bytes": REC_BYTES, "key_bytes": KEY_BYTES,
        "bin": "synthetic super-k-mers (kb_collector format), ~12 k-mers/super-k-mer, 30x duplicate-rich, 1% substitutions",
        "bins_per_step": n_gpus,
        "l2": "per-step working set ~%.1f GB (2 record buffers) >> 126 MB L2; two input bins alternate between steps" % (2 * n_rec * REC_BYTES / 1e9),
    }


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ reference arm
def reference_lib():
    from kmc_testlib import Reference, reference_available, ensure_reference_built
    if not (reference_available() or ensure_reference_built()):
        return None
    return Reference()


def to_testbin(sk):
    from kmc_testlib import Bin
    return Bin(data=sk.data, n_rec=sk.n_rec, n_super_kmers=sk.n_super_kmers, pack_bytes=sk.pack_bytes, pack_recs=sk.pack_bytes, k=K)


def cpu_sample_plan(R, prm, n_rec, cores, seconds_per_step):
    """Bounded sample of a reference step.  The reference parallelises over bins (n_sorters threads, kmc.h:1576-1584) and,
    for k=31, expands/compacts a bin on ONE thread (kb_sorter.h:299-362,1128-1281), so it is given what it is best at:
    cores/8 bins (capped at 16) of the workload's bin size side by side, ~8 threads each, shrunk only if the time budget
    of a step requires it."""
    import kmc_b200
    nb = int(min(max(cores // 8, 1), 16))
    probe = to_testbin(kmc_b200.synth_bin(777, K, 1 << 22))
    R.process_bins([probe] * nb, prm, n_sorters=cores)          # also faults the arena in
    t0 = time.perf_counter()
    R.process_bins([probe] * nb, prm, n_sorters=cores)
    rate = nb * probe.n_rec / (time.perf_counter() - t0)
    per = int(min(n_rec, max(1 << 22, rate * seconds_per_step / nb)))
    return nb, per


def run_reference_steps(R, prm, bins, cores, steps, warmup):
    total = sum(b.n_rec for b in bins)
    times, sort_times = [], []
    for i in range(warmup + steps):
        _, (wall, t_sort) = R.process_bins(bins, prm, n_sorters=cores)
        if i >= warmup:
            times.append(wall)
            sort_times.append(t_sort)
    return total, times, sort_times


def cpu_baseline_block(n_rec, seconds=12.0):
    """Reported beside the GPU number (rank 0, N=1): the unmodified reference on the box's host cores, bounded sample."""
    import kmc_b200
    from kmc_testlib import Params
    R = reference_lib()
    cores = os.cpu_count() or 1
    if R is None:
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "oracle/_ref not available on this box"}
    prm = Params(k=K, cutoff_min=CUTOFF_MIN, cutoff_max=CUTOFF_MAX, counter_max=COUNTER_MAX, lut_prefix_len=LUT_P)
    nb, per = cpu_sample_plan(R, prm, n_rec, cores, seconds / 2)
    b = to_testbin(kmc_b200.synth_bin(4242, K, per))
    total, times, st = run_reference_steps(R, prm, [b] * nb, cores, 1, 1)
    return {"value": total / times[0], "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "%d bin(s) x %d k-mers through the unmodified CKmerBinSorter<1>::ProcessBins + RADULS AVX2 (oracle/_ref), n_sorters=%d, wall %.2f s (sort_func %.2f thread-s)" % (nb, per, cores, times[0], st[0])}


def main_reference(args, rank, world):
    if rank != 0:
        return
    import kmc_b200
    from kmc_testlib import Params
    cores = os.cpu_count() or 1
    R = reference_lib()
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(args.n_rec, args.gpus)}
    if R is None:
        base["unavailable"] = "oracle/_ref/libkmc_ref.so was not built (needs /root/reference at build time)"
        print(json.dumps(base))
        return
    prm = Params(k=K, cutoff_min=CUTOFF_MIN, cutoff_max=CUTOFF_MAX, counter_max=COUNTER_MAX, lut_prefix_len=LUT_P)
    budget = 150.0 / max(args.steps + args.warmup, 1)
    nb, per = cpu_sample_plan(R, prm, args.n_rec, cores, budget)
    b0 = to_testbin(kmc_b200.synth_bin(4242, K, per))
    b1 = to_testbin(kmc_b200.synth_bin(4243, K, per)) if nb > 1 else b0
    bins = [b0 if i % 2 == 0 else b1 for i in range(nb)]
    total, times, st = run_reference_steps(R, prm, bins, cores, args.steps, args.warmup)
    t = sum(times)
    value = total * args.steps / t
    sample = "per step %d bin(s) x %d k-mers, unmodified CKmerBinSorter<1>::ProcessBins + RADULS AVX2 (oracle/_ref), n_sorters=%d host threads" % (nb, per, cores)
    base.update({"value": value, "ms_per_step": 1e3 * t / args.steps,
                 "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
                 "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(base))


# ------------------------------------------------------------------------------------------------ our arm
def main_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import kmc_b200
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n_rec = args.n_rec

    ctx = kmc_b200.Stage2Context(kmc_b200.Stage2Params(K, True, CUTOFF_MIN, CUTOFF_MAX, COUNTER_MAX, LUT_P), device=local_rank, n_slots=E2E_SLOTS)
    # two different bins per rank, alternating between steps
    host_bins = [kmc_b200.synth_bin(1000 + 17 * rank + j, K, n_rec) for j in range(2)]
    cap = ctx.out_capacity(n_rec) + 64

    # ---- value: inputs resident in HBM
    d_bins = []
    for hb in host_bins:
        t = torch.zeros(hb.size + 64, dtype=torch.uint8, device=dev)
        t[:hb.size] = torch.from_numpy(hb.data).to(dev)
        d_bins.append(t)
    d_out = torch.zeros(cap, dtype=torch.uint8, device=dev)
    d_lut = torch.zeros(ctx.lut_entries, dtype=torch.int64, device=dev)
    d_res = torch.zeros(8, dtype=torch.int64, device=dev)
    tstream = torch.cuda.Stream(device=dev)          # a real (non-default) stream: the library enqueues on it, torch events time it
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step_dev(i):
        hb = host_bins[i % 2]
        ctx.dev_process_bin(0, d_bins[i % 2].data_ptr(), hb.size, n_rec, hb.pack_bytes, d_out.data_ptr(), cap, d_lut.data_ptr(), d_res.data_ptr(), stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_dev(i)
    barrier()
    res = d_res.cpu().numpy()
    assert int(res[3]) == n_rec and int(res[5]) == 0 and int(res[6]) == 0, "warm-up step failed: %s" % res
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_dev(args.warmup + i)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = ctx.kernel_launches() - launches0
    st = ctx.stage_times(0)                      # CUDA events of the last timed step, recorded on the launching stream
    t_dev = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms = float(t_dev.item())
    value = world * n_rec * args.steps / (dev_ms * 1e-3)

    # ---- e2e: host buffers through submit/wait, two slots, pinned memory
    pin_bins = [torch.from_numpy(hb.data.copy()).pin_memory() for hb in host_bins]
    pin_out = [torch.zeros(cap, dtype=torch.uint8).pin_memory() for _ in range(E2E_SLOTS)]
    pin_lut = [torch.zeros(ctx.lut_entries, dtype=torch.int64).pin_memory() for _ in range(E2E_SLOTS)]

    def e2e_run(n_steps, first):
        moved_in = moved_out = 0
        for i in range(n_steps):
            s = (first + i) % E2E_SLOTS
            if i >= E2E_SLOTS:
                nb, stats = ctx.wait_bin(s)
                moved_out += nb
            hb = host_bins[i % 2]
            ctx.submit_bin(s, pin_bins[i % 2].data_ptr(), hb.size, n_rec, hb.pack_bytes, pin_out[s].data_ptr(), cap, pin_lut[s].data_ptr())
            moved_in += hb.size + 8 * (hb.pack_bytes.size + 1)
        for i in range(max(n_steps - E2E_SLOTS, 0), n_steps):
            nb, stats = ctx.wait_bin((first + i) % E2E_SLOTS)
            moved_out += nb
            assert stats[3] == n_rec
        return moved_in, moved_out + n_steps * (8 * ctx.lut_entries + 64)

    e2e_run(max(args.warmup, E2E_SLOTS), 0)
    barrier()
    t0 = time.perf_counter()
    h2d, d2h = e2e_run(args.steps, 0)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    t_e = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    t_e2e = float(t_e.item())
    e2e_value = world * n_rec * args.steps / t_e2e
    clocks = sampler.stop() if rank == 0 else None          # sampled across both timed regions (device-resident and end-to-end)

    if rank == 0:
        peak, peak_src = hbm_peak()
        names = st.get("pass_names") or ["radix_pass"] * len(st["pass_ms"])
        intervals = dict()
        for nm, x in zip(names, st["pass_ms"]):
            intervals.setdefault(nm, []).append(x)
        part = [x for nm, v in intervals.items() if nm.startswith("msd_partition") for x in v]
        if part:        # hybrid MSD sort: the two partition passes are the radix passes (1 read + 1 write of every record each)
            pass_ms, kernel_name = part, "msd_partition_kernel<1> (one 8-bit MSD partition pass over %d 8-byte records)" % n_rec
        else:
            pass_ms, kernel_name = [x for x in st["pass_ms"] if x > 0], "radix_pass_kernel<1> (one 8-bit LSD pass over %d 8-byte records)" % n_rec
        avg_pass = sum(pass_ms) / len(pass_ms)
        alg_bytes = 2.0 * n_rec * REC_BYTES                # one read + one write of every record
        achieved = alg_bytes / (avg_pass * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "radix_pass_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": workload_config(n_rec, world),
            "roofline": {"bound": "hbm", "kernel": kernel_name,
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_pass,
                         "pass_ms": pass_ms, "sort_intervals_ms": {k: v for k, v in intervals.items()},
                         "stage_ms": {"expand": st["expand_ms"], "sort": st["sort_ms"], "count": st["count_ms"]}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                    "ms_per_step": 1e3 * t_e2e / args.steps, "api": "kmcb200_submit_bin/kmcb200_wait_bin, %d slots, pinned host buffers" % E2E_SLOTS},
            "gpu_launches": launches, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline_block(n_rec)
        else:
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": "only measured at N=1"}
        print(json.dumps(out))
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-rec", type=int, default=1 << 26)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank, world)
    else:
        main_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
