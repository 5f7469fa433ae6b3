"""Turns gpurun_out/{launches_TAG.csv, prof_TAG.ncu-rep, bench_TAG*.json} into tracked summaries under profiles/."""
import collections, csv, json, os, subprocess, sys

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
out = ["# profile summary `%s`\n" % tag]

lp = os.path.join(G, "launches_%s.csv" % tag)
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        agg.setdefault(name, []).append(float(r[-1]) / 1e3)
    tot = sum(sum(v) for v in agg.values())
    out.append("## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, %d launches of `bench.py --steps 2 --warmup 3 --no-cpu --no-secondary`, taken inside the timed region; cold-cache, serialised: compare SHARES)\n" % len(rows))
    out.append("| kernel | launches | avg µs | min µs | max µs | share of listed time |\n|---|---|---|---|---|---|")
    for k, v in agg.items():
        out.append("| `%s` | %d | %.1f | %.1f | %.1f | %.1f %% |" % (k, len(v), sum(v) / len(v), min(v), max(v), 100 * sum(v) / tot))
    out.append("")
    open(os.path.join(P, "launches_%s.csv" % tag), "w").write(open(lp).read())

rp = os.path.join(G, "prof_%s.ncu-rep" % tag)
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
    cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs"),
            ("smsp__inst_executed.sum", "warp insts"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"), ("launch__grid_size", "grid")]
    out.append("## `ncu --set full` (one capture per kernel instance; %s)\n" % os.path.basename(rp))
    out.append("| # | kernel | " + " | ".join(c[1] for c in cols) + " |\n|---|---|" + "---|" * len(cols))
    traffic = {}          # stage of bench.py's roofline -> DRAM bytes of its main kernel (first capture of each)
    stage_of = [("leaf_hash_kernel", "leaf_count"), ("leaf_warp_kernel", "leaf_count"), ("expand_kernel", "expand")]
    def tobytes(x, unit):
        return float(x) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    n_part = 0
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        vals = []
        for c, _ in cols:
            v = r[idx[c]] if c in idx else ""
            u = rows[1][idx[c]] if c in idx else ""
            try: v = "%.4g" % float(v)
            except ValueError: pass
            vals.append(v + (" " + u if u and u not in ("%",) else ""))
        out.append("| %s | `%s` | " % (r[0], name) + " | ".join(vals) + " |")
        b = tobytes(r[idx["dram__bytes_read.sum"]], rows[1][idx["dram__bytes_read.sum"]]) + tobytes(r[idx["dram__bytes_write.sum"]], rows[1][idx["dram__bytes_write.sum"]])
        st = next((s for k, s in stage_of if k in name), None)
        if "msd_partition" in name:
            n_part += 1
            st = "msd_partition_L%d" % n_part if n_part <= 2 else None
        if "msd_count" in name: st = "msd_count_L2"
        if st and st not in traffic and float(r[idx["gpu__time_duration.sum"]]) > 20:          # (not the HEAVY instance that returns at once)
            traffic[st] = {"kernel": name, "dram_bytes_per_launch": b}
    out.append("")
    if traffic:
        src = "ncu --set full --clock-control none, %s: one bin of 117440512 k-mers (scripts/probe_bin.py), the workload's mean bin" % os.path.basename(rp)
        json.dump({"source": src, "stages": traffic}, open(os.path.join(P, "dominant_kernel_traffic.json"), "w"), indent=1)
        out.append("DRAM traffic per launch (bench.py reads the dominant stage's into roofline.traffic): " +
                   ", ".join("%s %.4g B" % (k, v["dram_bytes_per_launch"]) for k, v in traffic.items()) + "\n")

for suffix in ("", "_reference"):
    bp = os.path.join(G, "bench_%s%s.json" % (tag, suffix))
    if os.path.exists(bp):
        txt = open(bp).read().strip().split("\n")[-1]
        open(os.path.join(P, "bench_%s%s.json" % (tag, suffix)), "w").write(txt + "\n")
        d = json.loads(txt)
        out.append("## bench%s\n\n```json\n%s\n```\n" % (suffix, json.dumps({k: d[k] for k in d if k != "config"}, indent=1)))
open(os.path.join(P, "summary_%s.md" % tag), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
