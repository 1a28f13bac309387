#!/bin/bash
# full GPU pass: parity suite, bench, ncu launch list + full capture of the top kernels
mkdir -p gpurun_out
rm -f gpurun_out/stages.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/stages.txt
timeout 600 python bench.py --steps 20 --warmup 3 --extra > gpurun_out/bench_auto.log 2>&1; echo "bench rc=$?" >> gpurun_out/stages.txt
timeout 300 python bench.py --steps 20 --warmup 3 --variant 1 --no-cpu > gpurun_out/bench_v1.log 2>&1; echo "bench v1 rc=$?" >> gpurun_out/stages.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?" >> gpurun_out/stages.txt
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"decode_filter|nms_mask|nms_pairs|nms_reduce|seg_sort|detect_output|seg_tables" -s 42 -c 7 -o gpurun_out/prof_r1 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/stages.txt
cat gpurun_out/stages.txt
tail -n 3 gpurun_out/pytest_gpu.log
tail -n 2 gpurun_out/bench_auto.log
