#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass` dump: top SASS instructions by stall samples, global loads by
L2 sectors.  usage: ncu_lines.py sass.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]; col = {n: i for i, n in enumerate(hdr)}
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
data = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    def f(n):
        try: return float(r[col[n]])
        except Exception: return 0.0
    data.append(dict(addr=r[0], sass=r[1], samples=f("# Samples"), inst=f("Instructions Executed"), thr=f("Thread Instructions Executed"),
                     l2=f("L2 Theoretical Sectors Global"), l1req=f("L1 Tag Requests Global"), lsb=f("stall_long_sb"), lg=f("stall_lg"),
                     ssb=f("stall_short_sb"), mio=f("stall_mio"), wait=f("stall_wait"), local=f("L2 Theoretical Sectors Local")))
tot = sum(d["samples"] for d in data); ti = sum(d["inst"] for d in data)
print(f"instructions {len(data)}  samples {tot:.0f}  warp-inst {ti:.0f}")
print("--- top by stall samples (sample is attributed to the instruction WAITING)")
for i, d in enumerate(data): d["idx"] = i
for d in sorted(data, key=lambda d: -d["samples"])[:N]:
    print(f'{d["idx"]:5d} {d["samples"]/tot*100:5.1f}% inst {d["inst"]:10.0f} lsb {d["lsb"]:6.0f} lg {d["lg"]:5.0f} ssb {d["ssb"]:5.0f}  {d["sass"][:90]}')
print("--- global/local memory instructions by L2 sectors")
for d in sorted(data, key=lambda d: -(d["l2"] + d["local"]))[:N]:
    if d["l2"] + d["local"] == 0: break
    print(f'{d["idx"]:5d} l2sec {d["l2"]:11.0f} local {d["local"]:9.0f} req {d["l1req"]:10.0f} inst {d["inst"]:9.0f} thr/inst {d["thr"]/max(d["inst"],1):4.1f}  {d["sass"][:80]}')
