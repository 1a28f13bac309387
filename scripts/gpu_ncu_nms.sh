#!/bin/bash
# one full ncu capture of the per-image NMS kernel (source-level), after a plain run of the same command
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extra"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"image_nms" -s 4 -c 1 -o gpurun_out/prof_nms -f $CMD > gpurun_out/ncu_nms.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_nms.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/nms_source.csv 2>/dev/null
python profiles/source_hotspots.py gpurun_out/nms_source.csv 70 > gpurun_out/nms_hotspots.txt; head -75 gpurun_out/nms_hotspots.txt
