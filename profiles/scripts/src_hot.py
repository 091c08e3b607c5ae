"""Per-source-line instruction and stall-sample shares from an `ncu --page source --csv --print-source cuda,sass` export.
   python profiles/scripts/src_hot.py export.csv [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
agg = collections.defaultdict(lambda: [0, 0, ""])
f = ""; li = si = ii = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if r[0] == "Line No": li, si, ii, H = r.index("Line No"), r.index("# Samples"), r.index("Instructions Executed"), len(r); continue
    if li is None or len(r) < H or not r[li].isdigit(): continue
    x = len(r) - H                       # a source line with quotes in it splits into extra fields
    a = agg[(f, int(r[li]))]
    a[0] += int(r[si + x] or 0); a[1] += int(r[ii + x] or 0); a[2] = ",".join(r[1:2 + x])
ts = sum(v[0] for v in agg.values()) or 1; tn = sum(v[1] for v in agg.values()) or 1
print("stall samples", ts, "warp instructions", tn)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]}:{k[1]:4d}  inst {100 * v[1] / tn:5.1f}%  samples {100 * v[0] / ts:5.1f}%  {v[2].strip()[:110]}")
