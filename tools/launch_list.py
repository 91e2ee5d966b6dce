#!/usr/bin/env python3
"""Last bench step of an ncu launch list (--metrics gpu__time_duration.sum --csv): kernel, grid, block, ms, share."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
order = [(r[4], r[7], r[8], float(r[-1].replace(',', '')), r[-2]) for r in rows]
idx = [i for i, o in enumerate(order) if 'k_preprocess' in o[0]]
st = idx[-1]
ms = lambda o: o[3] / (1e6 if o[4] == 'ns' else 1e3 if o[4] == 'us' else 1)
tot = sum(ms(o) for o in order[st:])
print("kernel,grid,block,time_ms,share_pct")
for o in order[st:]:
    name = o[0].replace('b200at::', '').split('(')[0]
    print(f'"{name}","{o[2]}","{o[1]}",{ms(o):.4f},{100 * ms(o) / tot:.2f}')
print(f"total,,,{tot:.4f},100")
