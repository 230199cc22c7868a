#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall reasons overall and the hottest SASS instructions.

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:stem_kernel > src.csv
    python tools/ncu_src.py src.csv [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]]) for r in body)
print("instructions", len(body), "samples", tot, "warp-instr", sum(int(r[col["Instructions Executed"]]) for r in body))
agg = {s: sum(int(r[col[s]]) for r in body) for s in stalls}
print("stalls:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]]))[:top]
for i in sorted(order):
    r = body[i]
    rs = sorted(((int(r[col[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print("%5d %6.2f%% x%-8s %-70s %s" % (i, 100.0 * int(r[col["# Samples"]]) / max(tot, 1), r[col["Instructions Executed"]],
                                       r[col["Source"]].strip()[:70], " ".join("%s:%d" % (n, v) for v, n in rs if v)))
