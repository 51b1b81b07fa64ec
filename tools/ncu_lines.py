"""Per-source-line dynamic instruction counts / stall samples of a kernel from an ncu capture.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_lines.py src.csv [env-substeps in the captured launch] [top N]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
unit = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
cur = None
key = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0, ""])   # static sass, executed, samples, thread instrs, text
hdr = None
for r in rows:
    if r and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] in ('Line No',):
        hdr = r
        continue
    if len(r) < 9 or r[0] == 'Function Name':
        continue
    if r[0] != '':
        key = (cur, int(r[0]))
        try:
            agg[key][1] += int(r[7]); agg[key][2] += int(r[6]); agg[key][3] += int(r[8])
        except ValueError:
            pass
        agg[key][4] = r[1].strip()[:110]
    else:
        agg[key][0] += 1
tot_e = sum(v[1] for v in agg.values()); tot_s = sum(v[2] for v in agg.values())
print(f"total executed {tot_e} = {tot_e / unit:.0f} per unit; samples {tot_s}; static sass {sum(v[0] for v in agg.values())}")
byfile = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), v in agg.items():
    byfile[f][0] += v[0]; byfile[f][1] += v[1]; byfile[f][2] += v[2]
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][1]):
    print(f"  {f:28s} static {v[0]:6d} exec/unit {v[1] / unit:8.0f} ({v[1] / tot_e:5.3f}) samples {v[2] / max(1, tot_s):6.3f}")
print("--- top lines by executed instructions")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{l:5d} static {v[0]:4d} exec/unit {v[1] / unit:7.1f} ({v[1] / tot_e:5.3f}) samp {v[2] / max(1, tot_s):5.3f} thr {v[3] / max(1, v[1]):4.1f} | {v[4]}")
print("--- top lines by samples")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top // 2]:
    print(f"{f}:{l:5d} static {v[0]:4d} exec/unit {v[1] / unit:7.1f} samp {v[2] / max(1, tot_s):5.3f} | {v[4]}")
