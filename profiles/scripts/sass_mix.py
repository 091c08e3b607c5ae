"""Dynamic SASS opcode mix of one kernel from an `ncu --page source --csv --print-source cuda,sass` export.
   python profiles/scripts/sass_mix.py export.csv [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
H = ii = None
seen = {}
for r in rows:
    if not r: continue
    if r[0] == "Line No": H, ii = len(r), r.index("Instructions Executed"); continue
    if H and r[0] == "" and len(r) >= H:
        x = len(r) - H
        try: n = int(r[ii + x])
        except ValueError: continue
        seen[r[2 + x]] = (r[3 + x].strip(), n)
cnt = collections.Counter()
for sass, n in seen.values():
    t = sass.split()
    if not t: continue
    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    cnt[op.split(".")[0]] += n
tot = sum(cnt.values()) or 1
print("warp instructions", tot)
print(", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in cnt.most_common(top)))
