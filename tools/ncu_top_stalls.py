#!/usr/bin/env python
"""Print the instructions with the most warp-stall samples from `ncu --page source --csv` output."""
import csv, sys
path = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
si = hdr.index('# Samples'); so = hdr.index('Source'); ie = hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
data = [r for r in rows[hi + 1:] if len(r) > si and r[si].strip().isdigit()]
tot = sum(int(r[si]) for r in data)
print('total samples', tot, 'instructions', len(data))
for r in sorted(data, key=lambda r: -int(r[si]))[:n]:
    st = {hdr[i][6:]: int(r[i]) for i in stalls if r[i].strip().isdigit() and int(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{int(r[si]):6d} {int(r[si]) / tot * 100:5.1f}% ie={r[ie]:>9s} {r[so][:64]:64s} {st}")
