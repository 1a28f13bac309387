#!/bin/bash
# headline step time for several pipeline depths and step counts
for K in 10 20 50; do for d in 2 3 4 6; do
  python bench.py --steps $K --warmup 3 --no-cpu --no-extra --depth $d 2>/dev/null | tail -1 > /tmp/sweep_line.json
  python - "$K" "$d" <<'PY'
import json, sys
d = json.load(open("/tmp/sweep_line.json"))
print("steps", sys.argv[1], "depth", sys.argv[2], round(d["ms_per_step"] * 1e3, 2), "us/step")
PY
done; done
