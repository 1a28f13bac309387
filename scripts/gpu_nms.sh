#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x -k "nms or detect or post_process or paths" > gpurun_out/pytest_nms.log 2>&1; echo "pytest rc=$?"
tail -n 15 gpurun_out/pytest_nms.log
timeout 600 python scripts/bench_nms.py > gpurun_out/bench_nms.log 2>&1; echo "bench_nms rc=$?"; tail -n 3 gpurun_out/bench_nms.log
