#!/bin/bash
# profile the training step: event timings per form, ncu launch list (time + DRAM bytes), full capture of the match kernel
FORM=${1:-raw}
mkdir -p gpurun_out
rm -f gpurun_out/train_forms.log
for form in decoded raw split; do
  for b in 256 32; do
    timeout 300 python scripts/prof_train.py --form $form --batch $b >> gpurun_out/train_forms.log 2>&1
  done
done
grep "^train" gpurun_out/train_forms.log
TCMD="python scripts/prof_train.py --iters 2 --warmup 2 --form $FORM"
timeout 300 $TCMD > gpurun_out/train_plain_$FORM.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/train_launches_$FORM.csv $TCMD > gpurun_out/train_ncu_$FORM.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"loss_match|loss_bwd_rows|loss_dense" -s 9 -c 3 -o gpurun_out/prof_train_$FORM -f $TCMD > gpurun_out/train_ncu_full_$FORM.log 2>&1
echo "ncu full rc=$?"
