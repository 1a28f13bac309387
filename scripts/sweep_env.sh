#!/bin/bash
# usage: sweep_env.sh VAR v1 v2 ...   -- quick bench per value of an environment knob
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu --no-extra ${BENCH_ARGS} 2>/dev/null | tail -1 > /tmp/sweep_line.json
  python - "$v" <<'PY'
import json, sys
d = json.load(open("/tmp/sweep_line.json"))
print(sys.argv[1], round(d["ms_per_step"] * 1e3, 2), "us", d["detail"]["image_nms_kernel_stages_us"])
PY
done
