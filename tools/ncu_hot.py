"""Top stall sites from an ncu `--page source --csv --print-source cuda,sass` dump on stdin:
aggregates SASS rows under the CUDA source line that precedes them (handles several files)."""
import csv, sys, collections
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
hdr = None; cur = ("?", 0, ""); fname = "?"
agg = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    if r[0] != "":
        cur = (fname, int(r[0]), r[1].strip()); continue
    try:
        ins = int(float(d["Instructions Executed"])); smp = int(float(d["# Samples"]))
    except ValueError:
        continue
    a = agg.setdefault(cur, [0, 0]); a[0] += ins; a[1] += smp
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("instr", ti, "samples", ts)
for (f, ln, src), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*s/ts:5.1f}% smp {100*i/ti:5.1f}% inst  {f}:{ln}  {src[:100]}")
