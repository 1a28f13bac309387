#!/bin/bash
# staged first GPU run: each stage has its own timeout and log
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print(os.cpu_count())" > gpurun_out/cores.txt
PT="python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 180 -p no:cacheprovider"
timeout 600 $PT -k "nms" > gpurun_out/t1_nms.log 2>&1; echo "nms rc=$?" >> gpurun_out/stages.txt
timeout 600 $PT -k "assign or ciou or ratio or decode_scale" > gpurun_out/t2_train_small.log 2>&1; echo "assign/ciou rc=$?" >> gpurun_out/stages.txt
timeout 600 $PT -k "loss" > gpurun_out/t3_loss.log 2>&1; echo "loss rc=$?" >> gpurun_out/stages.txt
timeout 600 $PT -k "detect and (golden or oracle) and not 2]" > gpurun_out/t4_detect_v1.log 2>&1; echo "detect v1 rc=$?" >> gpurun_out/stages.txt
timeout 600 $PT -k "detect and (golden or oracle) and 2]" > gpurun_out/t5_detect_v2.log 2>&1; echo "detect v2 rc=$?" >> gpurun_out/stages.txt
timeout 600 $PT -k "config2" > gpurun_out/t6_config2.log 2>&1; echo "config2 rc=$?" >> gpurun_out/stages.txt
timeout 600 python bench.py --steps 20 --warmup 3 --extra > gpurun_out/bench_auto.log 2>&1; echo "bench rc=$?" >> gpurun_out/stages.txt
timeout 300 python bench.py --steps 20 --warmup 3 --variant 1 --no-cpu > gpurun_out/bench_v1.log 2>&1; echo "bench v1 rc=$?" >> gpurun_out/stages.txt
cat gpurun_out/stages.txt
tail -3 gpurun_out/t*.log
