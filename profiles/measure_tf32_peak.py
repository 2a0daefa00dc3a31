"""TF32 tensor-core peak of this B200, measured the way the driver measured MEASURED_PEAKS.json's bf16 figure:
cuBLAS matmul on 8192^3 fp32 operands with TF32 math allowed (2*N^3 flops), best of 10 launches (burst) and back to
back for 4 s (sustained), CUDA events.  The convolution family computes in TF32 on fp32 storage, so this is the tensor
ceiling it can be held to; bench.py reports `frac_of_tf32_peak` against the sustained figure when
profiles/r2_tf32_peak.json exists.

    python profiles/measure_tf32_peak.py > profiles/r2_tf32_peak.json
"""
import json
import time

import torch


def main():
    torch.backends.cuda.matmul.allow_tf32 = True
    n = 8192
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    c = torch.empty(n, n, device="cuda")
    fl = 2.0 * n ** 3
    for _ in range(5):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    k = 0
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            torch.matmul(a, b, out=c)
        k += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sus = fl * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
    # the same for bf16, as a cross-check against MEASURED_PEAKS.json on this very box
    ah, bh = a.bfloat16(), b.bfloat16()
    ch = torch.empty(n, n, device="cuda", dtype=torch.bfloat16)
    for _ in range(5):
        torch.matmul(ah, bh, out=ch)
    torch.cuda.synchronize()
    bb = 1e9
    for _ in range(10):
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        torch.matmul(ah, bh, out=ch)
        f1.record()
        torch.cuda.synchronize()
        bb = min(bb, f0.elapsed_time(f1))
    print(json.dumps({"tf32_tflops": fl / (best * 1e-3) / 1e12, "tf32_tflops_sustained": sus,
                      "bf16_tflops_same_box": fl / (bb * 1e-3) / 1e12, "gpu": torch.cuda.get_device_name(0),
                      "how": "torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS), best of 10 (burst) and back to back "
                             "for 4 s (sustained), CUDA events; bf16 burst on the same box as a cross-check"}))


if __name__ == "__main__":
    main()
