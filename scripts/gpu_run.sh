#!/bin/bash
# full GPU pass: parity suite, bench, training timing; with "ncu": launch lists + full capture of the top kernels
mkdir -p gpurun_out
rm -f gpurun_out/stages.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/stages.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_auto.log 2>&1; echo "bench rc=$?" >> gpurun_out/stages.txt
timeout 300 python scripts/prof_train.py > gpurun_out/train_plain.log 2>&1; echo "train rc=$?" >> gpurun_out/stages.txt
if [ "$1" = "ncu" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extra"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?" >> gpurun_out/stages.txt
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"decode_filter|image_nms" -s 10 -c 4 -o gpurun_out/prof_r1 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/stages.txt
TCMD="python scripts/prof_train.py --iters 2 --warmup 2"
timeout 300 $TCMD > gpurun_out/train_plain2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/train_launches.csv $TCMD > gpurun_out/train_ncu.log 2>&1
echo "ncu train rc=$?" >> gpurun_out/stages.txt
fi
cat gpurun_out/stages.txt
tail -n 4 gpurun_out/pytest_gpu.log
tail -n 1 gpurun_out/bench_auto.log
tail -n 1 gpurun_out/train_plain.log
