import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; agg = collections.OrderedDict(); tot = 0
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    v = float(d['Metric Value'].replace(',', '')); u = d['Metric Unit']
    ms = v / 1e6 if u.startswith('n') else (v / 1e3 if u.startswith('u') else v)
    k = d['Kernel Name'].split('(')[0].split('::')[-1]
    a = agg.setdefault(k, [0, 0.0, []]); a[0] += 1; a[1] += ms; a[2].append(ms); tot += ms
print(f"total {tot:.2f} ms")
for k, (c, ms, l) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}%  x{c:4d}  {k}" + ("  " + " ".join(f"{x:.2f}" for x in l) if 'bin_large' in k else ""))
