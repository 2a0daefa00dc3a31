"""Two training steps of a bench configuration inside a cudaProfiler range (for `ncu --profile-from-start off`):
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/r2_dram2.csv python scratch/ncu_step.py dec5 2"""
import contextlib
import io
import os
import sys
import types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "dec5"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B, H, W = bench.DEFAULT_BATCH[name], 512, 1024
dev = torch.device("cuda")
args = types.SimpleNamespace(config=name, batch=B, height=H, width=W)
model, params, task_step = bench.build_task(args, dev, 1, 0)
inputs = {k: v.to(dev) for k, v in bench.synthetic(B, H, W, 1234, name in ("joint", "depthmix")).items()}
def step():
    with contextlib.redirect_stdout(io.StringIO()):
        task_step(inputs)
for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(steps):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled %d steps of %s" % (steps, name))
