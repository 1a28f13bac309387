"""Top source lines of one kernel from `ncu -i rep --page source --csv --print-source cuda,sass --kernel-name K`:
aggregated warp-stall samples and executed instructions per (file, line)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, hdr = None, None
agg = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[2] == "-" and r[0].isdigit():
        key = (fname, int(r[0]))
        s = int(r[hdr.index("# Samples")] or 0)
        n = int(r[hdr.index("Instructions Executed")] or 0)
        bar = int(r[hdr.index("stall_barrier")] or 0)
        a = agg.setdefault(key, [0, 0, 0, r[1]])
        a[0] += s; a[1] += n; a[2] += bar
ts = sum(v[0] for v in agg.values()); tn = sum(v[1] for v in agg.values())
print("total samples %d, warp instructions %d" % (ts, tn))
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s %4d  samp %5.1f%%  inst %5.1f%%  barrier %5d  | %s" % (f, l, 100.0 * v[0] / max(ts, 1), 100.0 * v[1] / max(tn, 1), v[2], v[3].strip()[:90]))
