#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2; do
BG_DEBUG_SKIP=$d timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/exp_skip$d.log 2>&1
python - <<PY
import json
l=json.loads(open('gpurun_out/exp_skip$d.log').read().strip().splitlines()[-1])
print("skip=$d step %.1f us decode %.1f us" % (l['ms_per_step']*1e3, l['roofline']['kernel_ms']*1e3))
PY
done
