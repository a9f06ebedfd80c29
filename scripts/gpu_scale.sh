#!/bin/bash
# usage: gpu_scale.sh N  -- the driver's launch line for N GPUs of one box
N=$1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/scale_${N}_gpus.txt
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/scale_${N}.json 2> gpurun_out/scale_${N}.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_${N}.json 2> gpurun_out/scale_${N}.err
fi
tail -3 gpurun_out/scale_${N}.err; cut -c1-400 gpurun_out/scale_${N}.json
