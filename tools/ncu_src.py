#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: instruction mix and stall reasons per region
(before first FFMA2 / FFMA2 body / after last FFMA2) for the first instance of a kernel."""
import csv
import sys
import collections

rows = list(csv.reader(open(sys.argv[1])))
H = rows[1]
col = {h: i for i, h in enumerate(H)}
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(H):
        continue
    try:
        data.append({"src": r[col["Source"]].strip(), "samp": int(r[col["Warp Stall Sampling (All Samples)"]] or 0),
                     "exec": int(r[col["Instructions Executed"]] or 0),
                     "st": {s: int(r[col[s]] or 0) for s in stalls}})
    except ValueError:
        pass
# first instance only: cut where addresses restart (source repeats): use first EXIT-terminated block
n = len(data)
for i, d in enumerate(data):
    if d["src"].startswith("BRA") and i + 1 < n and data[i + 1]["src"].startswith("NOP"):
        # end of function body padding
        j = i + 1
        while j < n and data[j]["src"].startswith("NOP"):
            j += 1
        data = data[:j]
        break
ff = [i for i, d in enumerate(data) if d["src"].startswith("FFMA2") or d["src"].startswith("FFMA ")]
a, b = (ff[0], ff[-1] + 1) if ff else (0, 0)
tot_s = sum(d["samp"] for d in data) or 1
tot_e = sum(d["exec"] for d in data) or 1
print("instructions", len(data), "samples", tot_s, "warp-instr executed", tot_e)
for name, lo, hi in (("prologue", 0, a), ("fma body", a, b), ("epilogue", b, len(data))):
    seg = data[lo:hi]
    s = sum(d["samp"] for d in seg)
    e = sum(d["exec"] for d in seg)
    agg = collections.Counter()
    for d in seg:
        agg.update(d["st"])
    top = ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100.0 * v / max(s, 1)) for k, v in agg.most_common(6))
    mix = collections.Counter(d["src"].split()[0].split(".")[0] if not d["src"].startswith("@") else d["src"].split()[1].split(".")[0] for d in seg)
    wmix = collections.Counter()
    for d in seg:
        op = d["src"].split()[1] if d["src"].startswith("@") else d["src"].split()[0]
        wmix[op.split(".")[0]] += d["exec"]
    print("%-9s samples %5.1f%%  exec %5.1f%%  | %s" % (name, 100.0 * s / tot_s, 100.0 * e / tot_e, top))
    print("          dynamic mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_e) for k, v in wmix.most_common(8)))
