#!/bin/bash
# launch list of the training side (assignment + loss fwd/bwd at config 3)
mkdir -p gpurun_out
CMD="python scripts/prof_train.py --iters 2 --warmup 2"
timeout 300 python scripts/prof_train.py > gpurun_out/train_plain.log 2>&1; echo "train plain rc=$?"
timeout 300 $CMD > gpurun_out/train_plain2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/train_ncu.log 2>&1
echo "ncu rc=$?"
cat gpurun_out/train_plain.log | tail -3
