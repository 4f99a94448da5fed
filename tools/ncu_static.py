"""Per source line of an ncu source page (cuda,sass view): STATIC SASS instruction count, executed instructions, samples and
no-instruction (fetch) stall samples; first launch in the file only.  Shows where the code size of a kernel sits.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_static.py [topN]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
hdr = None; fname = ""; seen = set(); first_fn = None; skip = False
agg = {}
cur = None
for r in csv.reader(sys.stdin):
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name":
        continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        key = fname
        skip = key in seen
        seen.add(key)
        ia = hdr.index("Address"); ii = hdr.index("Instructions Executed"); ism = hdr.index("# Samples"); ini = hdr.index("stall_no_inst")
        continue
    if hdr is None or skip or len(r) < len(hdr):
        continue
    if r[0] != "":
        cur = (fname, int(r[0]), r[1].strip())
        agg.setdefault(cur, [0, 0, 0, 0])
    elif r[ia] != "" and cur is not None:
        a = agg[cur]; a[0] += 1
        try:
            a[1] += int(float(r[ii] or 0)); a[2] += int(float(r[ism] or 0)); a[3] += int(float(r[ini] or 0))
        except ValueError:
            pass
tot = [sum(v[i] for v in agg.values()) for i in range(4)]
print(f"static SASS {tot[0]}  executed {tot[1]}  samples {tot[2]}  no_inst samples {tot[3]}")
print("---- by static SASS count")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0][:18]:18s}{k[1]:5d} static {v[0]:5d} ({100*v[0]/tot[0]:4.1f}%) exec {100*v[1]/max(tot[1],1):4.1f}% smp {100*v[2]/max(tot[2],1):4.1f}% noinst {v[3]:4d}  {k[2][:90]}")
