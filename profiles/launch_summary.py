"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, share."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, d = None, collections.defaultdict(list)
for r in rows:
    if r[0] == 'ID':
        hdr = r
        continue
    if hdr is None:
        continue
    rec = dict(zip(hdr, r))
    try:
        v = float(rec['Metric Value'])
        if rec.get('Metric Unit', 'ns') in ('us', 'usecond'):
            v *= 1e3
        d[rec['Kernel Name'].split('(')[0][:48]].append(v)
    except (KeyError, ValueError):
        pass
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:50s} n={len(v):4d} avg={sum(v)/len(v)/1e3:9.1f} us  share={sum(v)/tot*100:5.1f}%")
