#!/bin/bash
# round-2 full pass on one GPU: parity suite, bench (default flags of the driver), training timings per input form,
# ncu launch lists (headline step; training step with DRAM bytes) and full captures of the dominant kernels
mkdir -p gpurun_out
rm -f gpurun_out/stages.txt gpurun_out/train_forms.log
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/cores.txt
timeout 2400 python -m pytest tests -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/stages.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2.log 2> gpurun_out/bench_r2.err; echo "bench rc=$?" >> gpurun_out/stages.txt
for form in decoded raw split; do for b in 256 32; do timeout 300 python scripts/prof_train.py --form $form --batch $b >> gpurun_out/train_forms.log 2>&1; done; done
echo "train forms rc=$?" >> gpurun_out/stages.txt
if [ "$1" = "ncu" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extra"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?" >> gpurun_out/stages.txt
for form in raw split; do
TCMD="python scripts/prof_train.py --iters 2 --warmup 2 --form $form"
timeout 300 $TCMD > gpurun_out/train_plain_$form.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/train_launches_$form.csv $TCMD > gpurun_out/train_ncu_$form.log 2>&1
echo "ncu train $form rc=$?" >> gpurun_out/stages.txt
done
TCMD="python scripts/prof_train.py --iters 2 --warmup 2 --form raw"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"loss_bwd_stream|loss_bwd_rows|loss_match|loss_dense" -s 12 -c 5 -o gpurun_out/prof_train_r2 -f $TCMD > gpurun_out/train_ncu_full.log 2>&1
echo "ncu train full rc=$?" >> gpurun_out/stages.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"decode_filter|image_nms" -s 10 -c 8 -o gpurun_out/prof_detect_r2 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu detect full rc=$?" >> gpurun_out/stages.txt
SCMD="python scripts/bench_seg.py --quick"
timeout 300 $SCMD > gpurun_out/seg_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"seg_fwd|seg_bwd" -s 9 -c 3 -o gpurun_out/prof_seg_r2 -f $SCMD > gpurun_out/ncu_seg.log 2>&1
echo "ncu seg full rc=$?" >> gpurun_out/stages.txt
fi
MCMD="python scripts/bench_seg.py --masks-only"
if [ "$1" = "ncu" ]; then
timeout 300 $MCMD > gpurun_out/segmask_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"seg_lowres|seg_upsample" -s 8 -c 2 -o gpurun_out/prof_segmask_r2 -f $MCMD > gpurun_out/ncu_segmask.log 2>&1
echo "ncu segmask full rc=$?" >> gpurun_out/stages.txt
fi
timeout 600 python scripts/bench_configs.py > gpurun_out/bench_configs.json 2> gpurun_out/bench_configs.err; echo "bench_configs rc=$?" >> gpurun_out/stages.txt
timeout 300 python scripts/prof_c5.py > gpurun_out/prof_c5.log 2>&1; echo "prof_c5 rc=$?" >> gpurun_out/stages.txt
timeout 300 python scripts/exp_pipe.py --depths 3,4,6 > gpurun_out/exp_pipe.log 2>&1; echo "exp_pipe rc=$?" >> gpurun_out/stages.txt
timeout 600 python scripts/bench_seg.py > gpurun_out/bench_seg.log 2>&1; echo "bench_seg rc=$?" >> gpurun_out/stages.txt
cat gpurun_out/stages.txt
tail -n 4 gpurun_out/pytest_gpu.log
grep "^train" gpurun_out/train_forms.log
