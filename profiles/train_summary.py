"""Summarise an ncu launch list with time + DRAM bytes per kernel (training side)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; data = []
for r in rows:
    if r and r[0] == 'ID': hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for d in data:
    v = float(d['Metric Value'].replace(',', '')); u = d['Metric Unit']
    if u in ('ns', 'nsecond'): v /= 1e3
    if u == 'Mbyte': v *= 1e6
    if u == 'Kbyte': v *= 1e3
    if u == 'Gbyte': v *= 1e9
    agg[d['Kernel Name']][d['Metric Name']].append(v)
tot = sum(sum(v['gpu__time_duration.sum']) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]['gpu__time_duration.sum'])):
    t = v['gpu__time_duration.sum']
    print("%-60s n=%3d avg=%9.1f us share=%5.1f%% rd=%8.1f MB wr=%8.1f MB" % (k[:60], len(t), sum(t) / len(t), 100 * sum(t) / tot,
          sum(v['dram__bytes_read.sum']) / len(t) / 1e6, sum(v['dram__bytes_write.sum']) / len(t) / 1e6))
