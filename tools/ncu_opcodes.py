"""Opcode histogram (weighted by executed count) from an ncu SASS source page csv on stdin."""
import csv, sys, collections
rows = list(csv.reader(sys.stdin))
hdr = None; h = collections.Counter(); tot = 0
for r in rows:
    if len(r) > 5 and r[0] == "Address":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    try: n = int(float(d["Instructions Executed"]))
    except ValueError: continue
    src = d["Source"].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    h[op] += n; tot += n
print("total", tot)
for op, n in h.most_common(40): print(f"{op:10s} {n:10d} {100*n/tot:5.1f}%")
