#!/usr/bin/env python
"""bench.py — images/sec of the 512x1024 3-frame monodepth training step (BASELINE.json configs[1]:
dec5 recipe, ResNet-50 OS16 encoder frozen, ASPP depth decoder, ResNet-18 pose net, batch 12 per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W]                  this repo (sm_100a kernels)
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  CPU arm: the reference algorithm
                                                                        (oracle port) on the host cores

One "step" = forward + photometric loss + backward + gradient all-reduce (N>1) + Adam.  One "image" = one
training sample (target + 2 source frames + intrinsics) — the reference's own definition (train.py:775-788).
Prints ONE JSON line on rank 0.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_GF_TRAIN_PER_SAMPLE = 1176.0   # config 2, fwd + dgrad + wgrad on trainable parts, 2*MACs


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=12, help="per-GPU batch (BASELINE.json: 12)")
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--ref-batch", type=int, default=2, help="bounded CPU sample: batch of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tc", action="store_true", help="force the generic CUDA-core convolution path")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port; /root/reference cannot travel to the GPU box)
# --------------------------------------------------------------------------------------------------
def cpu_threads():
    """Intra-op threads of the CPU arm: all host cores up to 32 (beyond that torch's CPU conv/BN kernels on this
    batch-2 workload get slower, not faster — 128 threads measured 10x slower than 8)."""
    return int(os.environ.get("SEGSDE_CPU_THREADS", min(os.cpu_count() or 1, 32)))


def cpu_step_factory(B, H, W):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import segsde_oracle as O
    import improving_segmentation_with_selfsupervised_depth_b200 as P
    from improving_segmentation_with_selfsupervised_depth_b200.synthetic import mono_config
    models, _ = P.install_dropin()
    with contextlib.redirect_stdout(io.StringIO()):
        template = models.get_model(mono_config("resnet50", H, W), 19).state_dict()   # shapes/keys only (CPU)
    sd = O.synthetic_state_dict(template, seed=0)
    trainable = [k for k in sd if sd[k].dtype.is_floating_point and "running" not in k
                 and not k.startswith("models.encoder.")]
    for k in trainable:
        sd[k].requires_grad_()
    opt = torch.optim.Adam([sd[k] for k in trainable], lr=1e-4)
    inputs = O.synthetic_inputs(B, H, W, seed=1234)
    cfg = {"num_layers": 50, "rswd": [False, False, True], "frame_ids": [0, -1, 1]}

    def step():
        opt.zero_grad()
        out = O.model_forward(sd, inputs, cfg, O.BNMode(True))
        noise = [torch.randn(B, 2, H, W) * 0.00001 for _ in range(4)]
        loss = O.monodepth_loss(inputs, [out[("disp", s)] for s in range(4)],
                                {f: out[("cam_T_cam", 0, f)] for f in (-1, 1)}, [0, -1, 1], H, W, noise=noise)["loss"]
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = cpu_threads()
    torch.set_num_threads(cores)
    B, H, W = args.ref_batch, args.height, args.width
    step = cpu_step_factory(B, H, W)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    sample = "fwd+loss+bwd+Adam, batch %d of the batch-12 workload, %dx%d, fp32, oracle port of the reference" % (B, H, W)
    print(json.dumps({
        "impl": "reference", "metric": "images/sec at 512x1024 3-frame monodepth", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "dec5 ResNet-50 monodepth train step, 512x1024 3-frame, batch 12/GPU (CPU sample: batch %d)" % B},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import improving_segmentation_with_selfsupervised_depth_b200 as P
    from improving_segmentation_with_selfsupervised_depth_b200 import _cabi as A
    from improving_segmentation_with_selfsupervised_depth_b200 import ops
    from improving_segmentation_with_selfsupervised_depth_b200.parallel import GradSync
    from improving_segmentation_with_selfsupervised_depth_b200.synthetic import MONO_LOSS_KW, mono_config, synthetic_inputs
    if args.no_tc:
        ops.USE_TC = False
    models, loss = P.install_dropin()
    B, H, W = args.batch, args.height, args.width
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = models.get_model(mono_config("resnet50", H, W, freeze_backbone=True), 19).to(dev).train()
    params = [p for p in model.parameters() if p.requires_grad]
    sync = GradSync(params) if world > 1 else None
    from improving_segmentation_with_selfsupervised_depth_b200 import optim as segsde_optim
    opt = segsde_optim.Adam(params, lr=1e-4)       # multi-tensor Adam of this library (no torch optimizer kernels)
    ml = loss.MonodepthLoss(height=H, width=W, batch_size=B, **MONO_LOSS_KW)

    host = {k: v.pin_memory() for k, v in synthetic_inputs(B, H, W, seed=1234 + rank).items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    # end-to-end arm: two device buffer sets filled from pinned host memory by a copy stream (double buffering)
    copy_stream = torch.cuda.Stream(device=dev)
    dev_bufs = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]

    def step(inputs):
        if sync is not None:
            sync.zero()
        else:
            opt.zero_grad(set_to_none=True)       # gradients come from the zero pools / fresh buffers
        out = model(inputs)
        ml.generate_images_pred(inputs, out)
        total = ml.compute_losses(inputs, out)["loss"]
        total.backward()
        if sync is not None:
            sync.reduce()
        opt.step()
        return total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, e2e):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        last = None
        if e2e:
            # every step's inputs travel pinned host -> device inside the timed region; the copy of step i+1 is
            # issued on a side stream into the other of two preallocated device buffer sets while step i computes
            # (no allocation inside the loop: a cudaMalloc there synchronises the device), the loss is read back
            # every step
            main = torch.cuda.current_stream()
            uploaded = [torch.cuda.Event(), torch.cuda.Event()]
            consumed = [torch.cuda.Event(), torch.cuda.Event()]

            def upload(slot, first_use):
                with torch.cuda.stream(copy_stream):
                    if not first_use:
                        copy_stream.wait_event(consumed[slot])      # the step that read this buffer set is done
                    for k, v in host.items():
                        dev_bufs[slot][k].copy_(v, non_blocking=True)
                    uploaded[slot].record(copy_stream)
            upload(0, True)
            for i in range(nsteps):
                slot = i & 1
                main.wait_event(uploaded[slot])
                if i + 1 < nsteps:
                    upload(slot ^ 1, i == 0)
                loss_t = step(dev_bufs[slot])
                consumed[slot].record(main)
                last = loss_t.item()                # D2H read of the step's loss
        else:
            for _ in range(nsteps):
                last = step(resident)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), last

    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(max(args.warmup, 3)):
            step(resident)
        timed(2, e2e=True)          # the host-buffer path is warmed up as well (untimed)
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # instrument every convolution launch with CUDA events during the timed region
    ops.PROFILE = []
    A.PROFILE, A.PROFILE_NAMES = [], {"segsde_reproj_fused"}
    n0 = A.launch_count()
    ms, last = timed(args.steps, e2e=False)
    launches = A.launch_count() - n0
    prof, ops.PROFILE = ops.PROFILE, None
    rprof, A.PROFILE, A.PROFILE_NAMES = A.PROFILE, None, None
    ms_e2e, last_e2e = timed(args.steps, e2e=True)
    clk = clocks.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family: implicit-GEMM convolutions (tensor bound) -------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured bf16_tflops_sustained (MEASURED_PEAKS.json)" if peaks else "fallback 1.4 PF sustained"
    fam = {}
    for kind, flops, e0, e1 in prof:
        d = fam.setdefault(kind, [0.0, 0.0, 0])
        d[0] += flops
        d[1] += e0.elapsed_time(e1)
        d[2] += 1
    tot_ms = sum(v[1] for v in fam.values()) or 1.0
    tot_fl = sum(v[0] for v in fam.values())
    roof = {"bound": "tensor", "achieved": tot_fl / (tot_ms * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": tot_fl / (tot_ms * 1e-3) / 1e12 / peak_tf, "traffic": None, "peak_source": peak_src,
            "kernel": "implicit-GEMM convolution family (fprop+dgrad+wgrad), all launches of the timed steps",
            "share_of_step": tot_ms / ms,
            "by_kind": {k: {"tflops": v[0] / (v[1] * 1e-3 + 1e-12) / 1e12, "ms_per_step": v[1] / args.steps,
                            "launches_per_step": v[2] / args.steps} for k, v in fam.items()}}

    # ---- the HBM-bound kernel of the path: fused reprojection + SSIM + L1 + auto-mask (+ gradients) ---------------
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    r_ms = sum(e0.elapsed_time(e1) for _, e0, e1 in rprof) or 1.0
    r_bytes = args.steps * sum(B * (4 * 9 * H * W + 4 * (H >> s) * (W >> s)) for s in range(4))
    roof_hbm = {"kernel": "reproj_kernel<GRAD> (4 launches per step, one per scale)", "bound": "hbm",
                "achieved": r_bytes / (r_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": r_bytes / (r_ms * 1e-3) / 1e9 / hbm_peak, "ms_per_step": r_ms / args.steps,
                "traffic": 344.5e6, "traffic_source": "ncu --set full, scale-0 launch, B=12: dram read 276.9 MB + write 67.6 MB "
                                                      "incl. the 50 MB identity-candidate cache written once per step "
                                                      "(profiles/r1_hot_kernels_final.md); instruction-bound, not HBM-bound"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import torch as _t
        cores = cpu_threads()
        _t.set_num_threads(cores)
        cstep = cpu_step_factory(args.ref_batch, H, W)
        t0 = time.perf_counter()
        cstep()
        dt = time.perf_counter() - t0
        cpu = {"value": args.ref_batch / dt, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "1 un-warmed step fwd+loss+bwd+Adam, batch %d, %dx%d fp32 (oracle port of the reference)"
                         % (args.ref_batch, H, W)}
    gb = B * world
    out = {
        "metric": "images/sec at 512x1024 3-frame monodepth", "value": gb * args.steps / (ms * 1e-3),
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32" if (ops.USE_TC and A.lib().segsde_tc_available()) else "f32", "data": "synthetic",
        "config": {"workload": "dec5 ResNet-50(OS16, frozen)+ASPP depth decoder+ResNet-18 pose, train step "
                               "(fwd+loss+bwd+allreduce+Adam), %dx%d 3-frame, batch %d/GPU" % (H, W, B),
                   "global_batch": gb, "parallelism": "dp%d" % world,
                   "l2": "inputs+activations per step >> 126 MB L2 (no flush needed)",
                   "train_gflop_per_image": FWD_GF_TRAIN_PER_SAMPLE},
        "e2e": {"value": gb * args.steps / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clk, "roofline": roof, "roofline_hbm_kernel": roof_hbm, "cpu_baseline": cpu,
        "loss": float(last.detach()) if hasattr(last, "detach") else float(last), "conv_roofline_frac_whole_step": FWD_GF_TRAIN_PER_SAMPLE * 1e9 * gb * args.steps
        / (ms * 1e-3) / 1e12 / peak_tf / world,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
