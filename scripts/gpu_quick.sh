#!/bin/bash
# quick GPU pass: parity suite, bench line, decode phase cycles, training timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?"
tail -n 1 gpurun_out/bench_quick.log
timeout 300 python scripts/prof_decode_cycles.py 2>&1 | tail -n 10
timeout 300 python scripts/prof_train.py > gpurun_out/train_plain.log 2>&1; tail -n 1 gpurun_out/train_plain.log
