#!/bin/bash
# scaling pass on one box (gpurun --gpus 8): the bench at N = 8, 4 (and optionally 2, 1) ranks, as the driver launches it
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/scale_n1.log 2> gpurun_out/scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_n$N.log 2> gpurun_out/scale_n$N.err
  fi
  echo "N=$N rc=$?"
  tail -c 300 gpurun_out/scale_n$N.err
done
