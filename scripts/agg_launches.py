import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hi = next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value'); mu = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi+1:]:
    if len(r) <= mv: continue
    name = re.sub(r'\(.*', '', r[kn]).replace('lb::<unnamed>::', '').replace('void ', '')
    v = float(r[mv].replace(',', ''))
    unit = r[mu]
    if unit == 'ns': v /= 1e3
    elif unit == 'ms': v *= 1e3
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]/1e3/div:9.3f} ms  {100*v[1]/tot:5.1f}%  n={int(v[0]/div):4d}  avg {v[1]/v[0]:8.1f} us  {k[:80]}")
print(f"total {tot/1e3/div:.3f} ms per run")
