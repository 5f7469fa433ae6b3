"""Per-source-line totals from `ncu --page source --csv --print-source cuda,sass`: instructions executed and stall samples by line."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
seen_fn = 0
fname = "?"
hdr = None
lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name":
        seen_fn += 1
        continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2 or not r[0].isdigit(): continue
    if seen_fn > 1 and False: break
    ix = {h: i for i, h in enumerate(hdr)}
    try:
        lines.append((fname, int(r[0]), r[1].strip(), int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), r))
    except (ValueError, KeyError): pass
# the report may hold several instances of the kernel: keep the first occurrence of each (file, line)
first = {}
for l in lines:
    first.setdefault((l[0], l[1]), l)
lines = list(first.values())
ti = sum(l[4] for l in lines); ts = sum(l[3] for l in lines)
print("instructions %d  samples %d" % (ti, ts))
for l in sorted(lines, key=lambda x: -x[4])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * l[4] / max(ti, 1), 100.0 * l[3] / max(ts, 1), l[0], l[1], l[2][:110]))
