#!/bin/bash
# usage: scratch/scale_run.sh N   -> dec5, joint, depthmix bench lines at N GPUs (one box), into gpurun_out/
N=$1
for c in dec5 joint depthmix; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --config $c --steps 10 --warmup 4 > gpurun_out/r2_scale_${c}_n$N.json 2> gpurun_out/r2_scale_${c}_n$N.err
  tail -1 gpurun_out/r2_scale_${c}_n$N.json | cut -c1-300
done
