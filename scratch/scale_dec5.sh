#!/bin/bash
# usage: scratch/scale_dec5.sh N  -> the default (dec5) bench line at N GPUs of one box
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 10 --warmup 4 > gpurun_out/r2_scale_dec5_n$N.json 2> gpurun_out/r2_scale_dec5_n$N.err
tail -1 gpurun_out/r2_scale_dec5_n$N.json | cut -c1-250
