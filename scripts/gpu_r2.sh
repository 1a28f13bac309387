#!/bin/bash
# round-2 GPU pass: parity suite (prints the measured keep-list parity), bench, training timings per input form;
# with "ncu": launch lists (time + DRAM bytes) of the training step and of the headline step
mkdir -p gpurun_out
rm -f gpurun_out/stages.txt
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/cores.txt
timeout 2400 python -m pytest tests -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/stages.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2.log 2> gpurun_out/bench_r2.err; echo "bench rc=$?" >> gpurun_out/stages.txt
for form in decoded raw split; do
  for b in 256 32; do
    timeout 300 python scripts/prof_train.py --form $form --batch $b >> gpurun_out/train_forms.log 2>&1
  done
done
echo "train forms rc=$?" >> gpurun_out/stages.txt
if [ "$1" = "ncu" ]; then
for form in raw split; do
TCMD="python scripts/prof_train.py --iters 2 --warmup 2 --form $form"
timeout 300 $TCMD > gpurun_out/train_plain_$form.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/train_launches_$form.csv $TCMD > gpurun_out/train_ncu_$form.log 2>&1
echo "ncu train $form rc=$?" >> gpurun_out/stages.txt
done
fi
cat gpurun_out/stages.txt
grep -c PASSED gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
tail -c 600 gpurun_out/bench_r2.log
cat gpurun_out/train_forms.log
