#!/usr/bin/env python
"""Per CUDA source line totals from `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`.  usage: ncu_srclines.py dump.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
def ci(name, start=0):
    try: return hdr.index(name, start)
    except ValueError: return None
c_s, c_i, c_l2, c_lsb, c_ssb, c_mio = ci("# Samples"), ci("Instructions Executed"), ci("L2 Theoretical Sectors Global"), ci("stall_long_sb"), ci("stall_short_sb"), ci("stall_mio")
c_loc = ci("L2 Theoretical Sectors Local")
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].isdigit(): continue
    def f(c):
        if c is None: return 0.0
        try: return float(r[c])
        except Exception: return 0.0
    lines.append((int(r[0]), r[1].strip()[:110], f(c_s), f(c_i), f(c_l2), f(c_lsb), f(c_ssb), f(c_mio), f(c_loc)))
ts = sum(l[2] for l in lines); ti = sum(l[3] for l in lines); tl2 = sum(l[4] for l in lines)
print(f"samples {ts:.0f} warp-inst {ti:.0f} l2 sectors {tl2:.0f} ({tl2*32/1e9:.2f} GB)")
print("--- by samples")
for l in sorted(lines, key=lambda l: -l[2])[:N]:
    print(f"{l[0]:5d} {l[2]/ts*100:5.1f}% inst {l[3]/ti*100:5.1f}% l2 {l[4]/max(tl2,1)*100:5.1f}% lsb {l[5]:6.0f} ssb {l[6]:6.0f} mio {l[7]:6.0f} | {l[1]}")
print("--- by L2 sectors")
for l in sorted(lines, key=lambda l: -l[4])[:12]:
    print(f"{l[0]:5d} l2 {l[4]/max(tl2,1)*100:5.1f}% ({l[4]*32/1e9:.2f} GB) local {l[8]:.0f} | {l[1]}")
