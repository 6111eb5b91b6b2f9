"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
data = rows[hi + 1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in data:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    u = r[ui]
    if u == 'ns':
        v /= 1000
    elif u == 'ms':
        v *= 1000
    elif u == 's':
        v *= 1e6
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('<unnamed>::', '')
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
out = [f"# {title}", "# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and "
       "serialised: compare SHARES", "kernel,launches,total_us,share"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    out.append(f"{k},{n},{t:.1f},{t / tot:.4f}")
out.append(f"TOTAL,{sum(v[0] for v in agg.values())},{tot:.1f},1.0")
open(dst, 'w').write("\n".join(out) + "\n")
print("\n".join(out[:22]))
