"""Per-source-line instruction / stall-sample shares from `ncu --page source --csv --print-source cuda,sass` output.
    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_lines.py src.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
REASONS = ['stall_barrier', 'stall_branch_resolving', 'stall_dispatch', 'stall_lg', 'stall_long_sb', 'stall_math',
           'stall_mio', 'stall_no_inst', 'stall_not_selected', 'stall_selected', 'stall_short_sb', 'stall_tex',
           'stall_wait', 'stall_misc', 'stall_sleep', 'stall_membar', 'stall_drain']
def num(v):
    try:
        return int(v.replace(',', ''))
    except ValueError:
        return 0


agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
hdr = cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r[0].isdigit():
        continue
    key = (cur, int(r[0]), r[1].strip()[:80])
    a = agg[key]
    a[0] += num(r[hdr.index("Instructions Executed")])
    a[1] += num(r[hdr.index("# Samples")])
    for x in REASONS:
        if x in hdr:
            a[2][x] += num(r[hdr.index(x)])
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print(f"total instructions {ti}, samples {ts}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    why = " ".join(f"{n[6:]}:{c * 100 // max(v[1], 1)}" for n, c in v[2].most_common(3))
    print(f"{v[1] / ts * 100:5.1f}% samp {v[0] / ti * 100:5.1f}% inst  {k[0]}:{k[1]:<4d} [{why}]  {k[2]}")
