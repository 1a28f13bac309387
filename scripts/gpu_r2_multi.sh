#!/bin/bash
# round-2 multi-GPU pass (gpurun --gpus N): full parity suite incl. the NCCL and second-device tests, bench at 1 and N ranks
N=${1:-2}
mkdir -p gpurun_out
rm -f gpurun_out/stages.txt gpurun_out/train_forms.log
nvidia-smi --query-gpu=name --format=csv,noheader > gpurun_out/gpus.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/stages.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?" >> gpurun_out/stages.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?" >> gpurun_out/stages.txt
for form in decoded raw split; do
  for b in 256 32; do
    timeout 300 python scripts/prof_train.py --form $form --batch $b >> gpurun_out/train_forms.log 2>&1
  done
done
BG_MATCH_OCC=4 timeout 300 python scripts/prof_train.py --form raw --batch 256 >> gpurun_out/train_forms.log 2>&1
cat gpurun_out/stages.txt
tail -n 6 gpurun_out/pytest_gpu.log
grep "^train" gpurun_out/train_forms.log
tail -c 300 gpurun_out/bench_n$N.err
