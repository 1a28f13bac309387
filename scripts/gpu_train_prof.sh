#!/bin/bash
# launch list of the training side (assignment + loss fwd/bwd at config 3) + quick check of the inference bench
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?"
python -c "
import json; l=json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1]); print('step %.1f us, decode %.1f us, value %.0f' % (l['ms_per_step']*1e3, l['roofline']['kernel_ms']*1e3, l['value']))"
CMD="python scripts/prof_train.py --iters 2 --warmup 2"
timeout 300 python scripts/prof_train.py > gpurun_out/train_plain.log 2>&1; echo "train plain rc=$?"
timeout 300 $CMD > gpurun_out/train_plain2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/train_ncu.log 2>&1
echo "ncu rc=$?"
tail -n 1 gpurun_out/train_plain.log
