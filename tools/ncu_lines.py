"""Aggregate an ncu source page (cuda,sass view) per source line: instructions executed + stall samples.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_lines.py [topN]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
hdr = None
out = []
for r in rows:
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    d = dict(zip(hdr, r))
    try:
        out.append((int(r[0]), r[1].strip(), int(float(d["Instructions Executed"])), int(float(d["# Samples"]))))
    except ValueError:
        pass
tot_i = sum(o[2] for o in out); tot_s = sum(o[3] for o in out)
print(f"total instructions {tot_i}  total samples {tot_s}")
print("---- by instructions")
for ln, src, ins, smp in sorted(out, key=lambda o: -o[2])[:top]:
    print(f"{ln:5d} {100*ins/tot_i:5.1f}% inst {100*smp/max(tot_s,1):5.1f}% smp  {src[:110]}")
print("---- by samples")
for ln, src, ins, smp in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{ln:5d} {100*ins/tot_i:5.1f}% inst {100*smp/max(tot_s,1):5.1f}% smp  {src[:110]}")
