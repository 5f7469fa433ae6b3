"""Summarise an `ncu --page source --csv` dump: stall-reason totals and the hottest SASS instructions."""
import csv, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: samples = int(r[ix["# Samples"]])
    except ValueError: continue
    for s in stalls:
        tot[s] += int(r[ix[s]] or 0)
    data.append((samples, r))
allsamp = sum(d[0] for d in data)
print("total samples", allsamp)
for s, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
    print("  %-28s %8d  %5.1f%%" % (s, v, 100.0 * v / max(allsamp, 1)))
print("hottest instructions:")
for samples, r in sorted(data, key=lambda x: -x[0])[:top]:
    best = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print("  %6d %5.1f%%  %-70s  exec=%s  %s" % (samples, 100.0 * samples / allsamp, r[ix["Source"]].strip()[:70], r[ix["Instructions Executed"]],
          ", ".join("%s=%d" % (s[6:], v) for v, s in best)))
