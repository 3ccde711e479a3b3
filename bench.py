#!/usr/bin/env python
"""bench.py -- the Conv2D hot path on the VGG16 shapes (BASELINE.json configs[2] / metric), on N B200s of one node.

A "step" is one pass of the hot path over one batch of synthetic input: for each of VGG16's 13 conv layers at
512x512x3 input resolution and `--batch` images per GPU
    forward  (Conv2DBiasActivation: 3x3 s1 p1 + bias + ReLU)      -> nb200_conv2d_forward
    input gradient  (Conv2DInputGradient)                        -> nb200_conv2d_input_gradient
    kernel gradient (Conv2DKernelsGradient)                      -> nb200_conv2d_kernels_gradient
followed by the data-parallel tail of ModelBase::Fit: sum-all-reduce of the 14.7 M kernel gradients over NCCL
(N > 1) and one fused Adam update (grad scale 1/N). Batch sharding = weak scaling: every rank holds `--batch` images.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for the meaning of every key.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# (C, K, H=W) of the 13 conv layers of VGG16 at 512x512 input: Neuro/src/Applications/VGG16.cpp:73-91
VGG16 = [(3, 64, 512), (64, 64, 512), (64, 128, 256), (128, 128, 256), (128, 256, 128), (256, 256, 128), (256, 256, 128),
         (256, 512, 64), (512, 512, 64), (512, 512, 64), (512, 512, 32), (512, 512, 32), (512, 512, 32)]
F, STRIDE, PAD = 3, 1, 1


def layer_flops(C, K, HW, batch=1):
    return 2.0 * batch * K * HW * HW * C * F * F


SAMPLE_FLOPS_PER_OP = sum(layer_flops(*l) for l in VGG16)   # 160.4 GFLOP per image for fwd (= dgrad = wgrad)


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe): one looping nvidia-smi
    process (-lms 50) started before and terminated after the region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        self.rows = []
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill(); out = ""
        self.rows = [[v.strip() for v in line.split(",")] for line in out.splitlines() if line.strip()]

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        smax = int(float(self.rows[0][1])) if self.rows else None
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm

def cpu_reference_run(steps, warmup, res):
    """Times the reference's own CPU implementation (oracle/_ref = TensorOpCpuMt compiled from the reference sources;
    else the oracle port) on a bounded sample: the same 13-layer stack, batch 1, at `res` x `res` input."""
    import numpy as np
    from neuro__b200 import synth
    from oracle import oracle as O

    use_ref = O.have_ref()
    if use_ref:
        O.ref_set_threads(0)
        cores = O.ref_threads()
    else:
        O.build(ref=False)
        cores = os.cpu_count()
    scale = 512 // res
    layers = [(C, K, HW // scale) for (C, K, HW) in VGG16]
    frac = sum(layer_flops(*l) for l in layers) / SAMPLE_FLOPS_PER_OP
    data = []
    for i, (C, K, HW) in enumerate(layers):
        x = synth.uniform(synth.SEED_X + i, (1, C, HW, HW))
        w = synth.glorot_uniform(synth.SEED_W + i, K, C, F, F)
        dy = synth.uniform(synth.SEED_DY + i, (1, K, HW, HW))
        data.append((x, w, dy, HW))

    def step():
        for (x, w, dy, HW) in data:
            if use_ref:
                O.ref_conv2d(x, w, STRIDE, PAD, PAD, mt=True)
                O.ref_conv2d_input_gradient(dy, w, STRIDE, PAD, PAD, (HW, HW), mt=True)
                O.ref_conv2d_kernels_gradient(x, dy, STRIDE, PAD, PAD, (F, F), mt=True)
            else:
                O.conv2d(x, w, STRIDE, PAD, PAD)
                O.conv2d_input_gradient(dy, w, STRIDE, PAD, PAD, (HW, HW))
                O.conv2d_kernels_gradient(x, dy, STRIDE, PAD, PAD, (F, F))

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    value = frac / dt    # equivalent full-resolution samples per second (FLOP-linear extrapolation of the sample)
    sample = ("VGG16 13-layer fwd+dgrad+wgrad, batch 1, %dx%dx3 input (%.5f of the 481.2 GFLOP per 512x512 sample), "
              "samples/s extrapolated linearly in FLOPs; %s" % (res, res, frac,
              "reference TensorOpCpuMt (PPL->OpenMP; dgrad parallel over N only, wgrad over K only)" if use_ref else "oracle port (OpenMP)"))
    return {"value": value, "unit": "samples/s", "cores": cores, "kind": "reference" if use_ref else "port", "sample": sample,
            "seconds_per_step": dt, "gflops": 3 * frac * SAMPLE_FLOPS_PER_OP / dt / 1e9}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: the largest input resolution whose estimated cost keeps the whole run within ~2.5 minutes
    # (seconds per step measured on 16 host cores: 64x64 ~ 9 s, 32x32 ~ 2.3 s, 16x16 ~ 0.7 s; dgrad is single-threaded at N=1)
    cores = os.cpu_count() or 8
    total_steps = args.steps + args.warmup
    res = 16
    for cand, est in ((64, 9.0), (32, 2.3)):
        if est * max(1.0, 16.0 / cores) * total_steps <= 150.0:
            res = cand
            break
    cb = cpu_reference_run(args.steps, args.warmup, res)
    line = {
        "impl": "reference", "metric": "vgg16_conv_fwd_dgrad_wgrad_samples_per_s", "value": cb["value"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch, args.gpus),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(batch, gpus):
    return {"workload": "VGG16 conv stack @512x512x3 (13 layers 3x3 s1 p1): fwd(bias+ReLU) + input-gradient + kernel-gradient "
                        "+ gradient all-reduce + Adam",
            "per_gpu_batch": batch, "global_batch": batch * gpus, "parallelism": "dp%d (batch sharded)" % gpus,
            "math": "tf32 tensor cores, fp32 accumulate (first layer C=3: HBM-bound; forward / input gradient on fp32 CUDA cores, kernel gradient on tensor cores)",
            "l2": "inputs larger than L2: each step streams every layer's tensors once (%.1f GB of distinct tensors per step)" % (batch * 0.95)}


# ------------------------------------------------------------------------------------------------ our arm

def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from neuro__b200 import lib, synth
    from neuro__b200.tensor_op import TensorOpB200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner / debug lines must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = lib.load()
    op = TensorOpB200(lib.MATH_TF32)
    B = args.batch
    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev); gen.manual_seed(synth.SEED_MODEL)   # same weights on every rank (replicas)
    dgen = torch.Generator(device=dev); dgen.manual_seed(1000 + rank)       # different data shard per rank

    # flat parameter / gradient / Adam-moment buckets with per-layer views (one all-reduce bucket per layer)
    sizes = [K * C * F * F for (C, K, HW) in VGG16]
    total = sum(sizes)
    params = torch.empty(total, device=dev); grads = torch.zeros(total, device=dev)
    m = torch.zeros(total, device=dev); v = torch.zeros(total, device=dev)
    layers = []
    off = 0
    for (C, K, HW), sz in zip(VGG16, sizes):
        limit = (6.0 / (C * F * F + K * F * F)) ** 0.5                       # GlorotUniform (VarianceScaling.cpp:59-65)
        w = params[off:off + sz].view(K, C, F, F)
        w.copy_((torch.rand(K, C, F, F, device=dev, generator=gen) * 2 - 1) * limit)
        dw = grads[off:off + sz].view(K, C, F, F)
        off += sz
        x = torch.rand(B, C, HW, HW, device=dev, generator=dgen) * 2 - 1   # U(-1,1) (Tensor::FillWithRand default)
        dy = torch.rand(B, K, HW, HW, device=dev, generator=dgen) * 2 - 1
        y = torch.empty(B, K, HW, HW, device=dev); dx = torch.empty(B, C, HW, HW, device=dev)
        bias = torch.zeros(K, device=dev)                                    # Conv2D bias init = zeros (Conv2D.h:46)
        desc = lib.ConvDesc(B, C, HW, HW, K, F, F, HW, HW, STRIDE, PAD, PAD, lib.NCHW, lib.MATH_TF32)
        layers.append(dict(x=x, w=w, dw=dw, dy=dy, y=y, dx=dx, bias=bias, desc=desc))
    names = [[op.kernel_name(o, l["desc"]) for o in (0, 1, 2)] for l in layers]

    lr, b1, b2, eps = 1e-5, 0.9, 0.999, 1e-8

    def step(events=None):
        def timed(tag, fn):
            if events is None:
                fn(); return
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); events.append((tag, e0, e1))
        for i, l in enumerate(layers):
            timed((0, i), lambda l=l: op.Conv2DBiasActivation(l["x"], l["w"], STRIDE, PAD, PAD, l["bias"], lib.ACT_RELU, 0.0, l["y"]))
        works = []
        for i in reversed(range(len(layers))):
            l = layers[i]
            timed((1, i), lambda l=l: op.Conv2DInputGradient(l["dy"], l["w"], STRIDE, PAD, PAD, lib.NCHW, l["dx"]))
            timed((2, i), lambda l=l: op.Conv2DKernelsGradient(l["x"], l["dy"], STRIDE, PAD, PAD, lib.NCHW, l["dw"]))
            if world > 1:   # exchange step: overlaps with the remaining layers' dgrad/wgrad
                works.append(dist.all_reduce(l["dw"], op=dist.ReduceOp.SUM, async_op=True))
        for wk in works:
            wk.wait()
        timed((3, 0), lambda: op.AdamStep(params, grads, m, v, lr, b1, b2, eps, gradScale=1.0 / world))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region (device-resident inputs) ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.15)   # let the first samples land inside the region
    launches0 = L.nb200_kernel_launches()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        step()
    t1.record()
    barrier()
    launches = L.nb200_kernel_launches() - launches0
    ms = t0.elapsed_time(t1) / args.steps
    if world > 1:
        tmax = torch.tensor([ms], device=dev); dist.all_reduce(tmax, op=dist.ReduceOp.MAX); ms = float(tmax.item())
    # second timed pass with a CUDA-event pair around every op call (feeds per_op and roofline; the ~80 extra event
    # records per step cost ~1 %, which is why the headline pass above runs without them)
    events = []
    op_steps = max(3, min(args.steps, 20))
    barrier()
    for _ in range(op_steps):
        step(events)
    barrier()
    if sampler:
        sampler.stop()
    fam_ms = {0: 0.0, 1: 0.0, 2: 0.0, 3: 0.0}
    tc_ms = {0: 0.0, 1: 0.0, 2: 0.0}; tc_fl = {0: 0.0, 1: 0.0, 2: 0.0}
    kern = {}   # CUDA kernel -> [flops, ms, launches] over the instrumented pass
    kern_bytes = {}
    KERNEL_OF = {"tcgen05_fprop": "tc_fprop_kernel (forward + input gradient, >64 filters)",
                 "tcgen05_dgrad": "tc_fprop_kernel (forward + input gradient, >64 filters)",
                 "tcgen05_rowtap_fprop": "tc_rowtap_kernel (forward + input gradient, <=64 filters)",
                 "tcgen05_rowtap_dgrad": "tc_rowtap_kernel (forward + input gradient, <=64 filters)",
                 "tcgen05_wgrad": "tc_wgrad_kernel (kernel gradient)",
                 "tcgen05_rowfold_wgrad": "tc_wgrad_rowfold_kernel (kernel gradient, <=64 channels and filters)",
                 "tcgen05_smallc_wgrad": "tc_smallc_wgrad_kernel (kernel gradient, <=4 channels, HBM-bound)"}
    for (fam, i), e0, e1 in events:
        dt = e0.elapsed_time(e1)
        fam_ms[fam] += dt
        if fam < 3 and names[i][fam].startswith("tcgen05") and "smallc" not in names[i][fam]:   # tensor-bound launches only
            tc_ms[fam] += dt; tc_fl[fam] += layers[i]["desc"].flops()
        if fam < 3:
            k = kern.setdefault(KERNEL_OF.get(names[i][fam], names[i][fam]), [0.0, 0.0, 0])
            k[0] += layers[i]["desc"].flops(); k[1] += dt; k[2] += 1
            l = layers[i]   # algorithmic bytes of the call: the two activation-sized tensors it streams + the filters
            kern_bytes[KERNEL_OF.get(names[i][fam], names[i][fam])] = kern_bytes.get(KERNEL_OF.get(names[i][fam], names[i][fam]), 0.0) + \
                4.0 * (l["x"].numel() + l["y"].numel() + l["w"].numel())

    # ---- e2e: the same step through the public API with HOST input and HOST result, copies inside the timed region ----
    # Every step's input batch comes from pinned host memory and its result (the image gradient, what style transfer reads
    # back) returns to host memory. As in any input pipeline (the reference has DataPreloader for this) the H2D copy of
    # step i+1 and the D2H copy of step i run on a copy stream underneath step i+1 / i's kernels; all of it is inside the
    # timed region and the region ends only when the last result has landed on the host.
    img_host = torch.from_numpy(synth.uniform(synth.SEED_X, (B, 3, 512, 512))).pin_memory()
    grad_host = [torch.empty((B, 3, 512, 512), dtype=torch.float32).pin_memory() for _ in range(2)]
    xbuf = [layers[0]["x"], torch.empty_like(layers[0]["x"])]
    dxbuf = [layers[0]["dx"], torch.empty_like(layers[0]["dx"])]
    copy_stream = torch.cuda.Stream()
    main = torch.cuda.current_stream()

    def e2e_run(n_steps):
        h2d = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]     # compute no longer reads xbuf[j] / has written dxbuf[j]
        d2h = [torch.cuda.Event(), torch.cuda.Event()]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_stream(main)
            xbuf[0].copy_(img_host, non_blocking=True); h2d[0].record(copy_stream)
        for i in range(n_steps):
            j = i & 1
            main.wait_event(h2d[j])
            if i >= 2:
                main.wait_event(d2h[j])                       # dxbuf[j] has been drained to the host
            layers[0]["x"], layers[0]["dx"] = xbuf[j], dxbuf[j]
            if i + 1 < n_steps:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(freed[j ^ 1])  # step i-1 is done with xbuf[j^1]
                    xbuf[j ^ 1].copy_(img_host, non_blocking=True); h2d[j ^ 1].record(copy_stream)
            step()
            freed[j].record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[j])
                grad_host[j].copy_(dxbuf[j], non_blocking=True); d2h[j].record(copy_stream)
        main.wait_stream(copy_stream)

    e2e_run(2); barrier()
    s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 20))
    s0.record()
    e2e_run(e2e_steps)
    s1.record()
    barrier()
    layers[0]["x"], layers[0]["dx"] = xbuf[0], dxbuf[0]
    e2e_ms = s0.elapsed_time(s1) / e2e_steps
    if world > 1:
        tmax = torch.tensor([e2e_ms], device=dev); dist.all_reduce(tmax, op=dist.ReduceOp.MAX); e2e_ms = float(tmax.item())

    if rank == 0:
        peaks = read_peaks()
        tf32_peak = peaks["bf16_sustained"] / 2.0       # TF32 dense = half the bf16 rate; sustained: kernels timed inside a long step
        # kernel level: the op-level event pairs are grouped by the CUDA kernel that op dispatched to (an op's pair also
        # covers its helper launches: the filter repack before tc_fprop/tc_rowtap, the split-K reduce after tc_wgrad; < 3 %)
        dom_name = max(kern, key=lambda k: kern[k][1])
        dom_fl, dom_ms, dom_n = kern[dom_name]
        achieved = dom_fl / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        # DRAM traffic of the dominant kernel: from the committed ncu --set full capture of the same workload (never measured
        # in this run: a profiler run is not a timing run); null when the capture has no row for the kernel
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_dram_traffic_per_launch.json")))
            ent = tj["kernels"].get(dom_name.split(" ")[0])
            if ent and B == 8:
                traffic, traffic_src = ent["dram_bytes_per_launch"], tj["source"]
        except (OSError, ValueError, KeyError):
            pass
        dom_bytes = kern_bytes.get(dom_name, 0.0)
        kernels = {k: {"ms_per_step": v[1] / op_steps, "launches_per_step": v[2] // op_steps,
                       "tflops": v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else None,
                       "share_of_step": v[1] / sum(fam_ms.values())} for k, v in kern.items()}
        per_op = {n: {"ms_per_step": fam_ms[f] / op_steps,
                      "tflops": (B * SAMPLE_FLOPS_PER_OP) / (fam_ms[f] / op_steps * 1e-3) / 1e12,
                      "tensor_core_tflops": (tc_fl[f] / (tc_ms[f] * 1e-3) / 1e12) if tc_ms[f] > 0 else None,
                      "frac_of_tf32_peak": ((tc_fl[f] / (tc_ms[f] * 1e-3) / 1e12) / tf32_peak) if tc_ms[f] > 0 else None}
                  for f, n in ((0, "forward"), (1, "input_gradient"), (2, "kernels_gradient"))}
        line = {
            "metric": "vgg16_conv_fwd_dgrad_wgrad_samples_per_s", "value": B * world / (ms * 1e-3), "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": workload_config(B, world),
            "tflops_total": 3 * B * world * SAMPLE_FLOPS_PER_OP / (ms * 1e-3) / 1e12,
            "per_op": per_op, "kernels": kernels, "adam_ms_per_step": fam_ms[3] / op_steps,
            "roofline": {"bound": "tensor", "kernel": dom_name, "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                         "frac": achieved / tf32_peak if tf32_peak else None, "traffic": traffic,
                         "traffic_unit": "DRAM bytes per launch (read + write)", "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": dom_bytes / dom_n if dom_n else None,
                         "avg_launch_ms": dom_ms / dom_n if dom_n else None,
                         "flops_per_launch": dom_fl / dom_n if dom_n else None,
                         "peak_source": "%s bf16_tflops_sustained / 2 (MEASURED_PEAKS.json has no TF32 entry; TF32 dense = 1/2 bf16)" % peaks["source"],
                         "share_of_step": dom_ms / sum(fam_ms.values()) if sum(fam_ms.values()) else None},
            "e2e": {"value": B * world / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": img_host.numel() * 4,
                    "d2h_bytes_per_step": grad_host[0].numel() * 4, "ms_per_step": e2e_ms},
            "gpu_launches": int(launches),
            "clocks": sampler.summary() if sampler else None,
        }
        if world == 1 and not args.no_extras:
            # extra measurements next to the headline; whatever happens here, the headline line above is still printed
            try:
                line["style_transfer_batch1"] = style_transfer_pass()
                line["style_transfer_batch1"]["frac_of_tf32_peak"] = line["style_transfer_batch1"]["tflops"] / tf32_peak
                line["neighbours"] = neighbour_pass(B)
                line["other_configs"] = other_configs_pass()
            except Exception as e:  # noqa: BLE001
                line["extras_error"] = str(e)[:300]
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_run(1, 1, 64)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def style_transfer_pass(iters=30):
    """BASELINE configs[2] as the reference runs it: ONE image, frozen VGG16 weights, per iteration the forward
    of the 13 conv layers (bias + ReLU fused) and then the input gradient of each back to the image -- no kernel gradient.
    Constant weights => filters prepared once (nb200_conv2d_prepare_filters); the 26-launch chain is launch-latency
    sensitive at batch 1 => also replayed as one CUDA graph. Returns ms per image for the three ways of issuing it."""
    import torch
    from neuro__b200 import lib, synth
    from neuro__b200.tensor_op import TensorOpB200
    dev = torch.device("cuda", torch.cuda.current_device())
    op = TensorOpB200(lib.MATH_TF32)
    gen = torch.Generator(device=dev); gen.manual_seed(synth.SEED_MODEL)
    L = []
    for (C, K, HW) in VGG16:
        limit = (6.0 / (C * F * F + K * F * F)) ** 0.5
        w = (torch.rand(K, C, F, F, device=dev, generator=gen) * 2 - 1) * limit
        x = torch.rand(1, C, HW, HW, device=dev, generator=gen) * 2 - 1
        dy = torch.rand(1, K, HW, HW, device=dev, generator=gen) * 2 - 1
        L.append(dict(w=w, x=x, dy=dy, y=torch.empty(1, K, HW, HW, device=dev), dx=torch.empty(1, C, HW, HW, device=dev),
                      bias=torch.zeros(K, device=dev)))
    for l in L:
        l["pf"] = op.PrepareKernels(lib.OP_FORWARD, l["x"], l["w"], l["y"], STRIDE, PAD, PAD)
        l["pg"] = op.PrepareKernels(lib.OP_INPUT_GRADIENT, l["dx"], l["w"], l["dy"], STRIDE, PAD, PAD)

    def chain(prepared):
        for l in L:
            op.Conv2DBiasActivation(l["x"], l["w"], STRIDE, PAD, PAD, l["bias"], lib.ACT_RELU, 0.0, l["y"], prepared=l["pf"] if prepared else None)
        for l in reversed(L):
            op.Conv2DInputGradient(l["dy"], l["w"], STRIDE, PAD, PAD, lib.NCHW, l["dx"], prepared=l["pg"] if prepared else None)

    def time_it(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    out = {"per_call_repack_ms": time_it(lambda: chain(False)), "prepared_filters_ms": time_it(lambda: chain(True))}
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            chain(True)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            chain(True)
        out["prepared_filters_cuda_graph_ms"] = time_it(g.replay)
    except Exception as e:  # noqa: BLE001 -- the graph is an extra; the two numbers above stand without it
        out["prepared_filters_cuda_graph_ms"] = None
        out["cuda_graph_error"] = str(e)[:200]
        torch.cuda.synchronize()
    best = min(v for k, v in out.items() if k.endswith("_ms") and v)
    out["gflop_per_image"] = 2 * SAMPLE_FLOPS_PER_OP / 1e9
    out["tflops"] = 2 * SAMPLE_FLOPS_PER_OP / (best * 1e-3) / 1e12
    out["workload"] = "VGG16 @512x512x3, batch 1: 13 x forward(bias+ReLU) + 13 x input gradient, frozen weights"
    return out


def other_configs_pass(iters=10):
    """The other BASELINE configs, per op over all their conv / transposed-conv layers (neuro__b200/shapes.py): DCGAN batch 128,
    pix2pix 256x256 batch 8, conv autoencoder batch 256. One layer's op is issued `iters` times back to back between a CUDA-event
    pair; a second measurement replays the same calls as one CUDA graph (GPU time without the host's per-call cost)."""
    import torch
    from neuro__b200 import lib
    from neuro__b200.shapes import CONFIGS
    from neuro__b200.tensor_op import TensorOpB200
    op = TensorOpB200(lib.MATH_TF32)
    peaks = read_peaks()
    tf32_peak = peaks["bf16_sustained"] / 2.0
    out = {}
    for cfg in ("dcgan", "pix2pix", "autoenc"):
        N, layers = CONFIGS[cfg]
        tot = {"issued": [0.0, 0.0, 0.0], "graph": [0.0, 0.0, 0.0]}
        flops = 0.0
        for (_, C, H, K, Fs, st, pd) in layers:
            Ho = (H + 2 * pd - Fs) // st + 1
            x = torch.randn(N, C, H, H, device="cuda"); w = torch.randn(K, C, Fs, Fs, device="cuda") * 0.05
            y = torch.empty(N, K, Ho, Ho, device="cuda"); dy = torch.randn_like(y); dx = torch.empty_like(x); dw = torch.empty_like(w)
            flops += 2.0 * N * K * Ho * Ho * C * Fs * Fs
            fns = [lambda: op.Conv2D(x, w, st, pd, pd, lib.NCHW, y), lambda: op.Conv2DInputGradient(dy, w, st, pd, pd, lib.NCHW, dx),
                   lambda: op.Conv2DKernelsGradient(x, dy, st, pd, pd, lib.NCHW, dw)]
            for i, fn in enumerate(fns):
                fn(); fn(); torch.cuda.synchronize()
                runs = {"issued": lambda fn=fn: [fn() for _ in range(iters)]}
                try:
                    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        fn()
                    torch.cuda.current_stream().wait_stream(side)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        for _ in range(iters):
                            fn()
                    g.replay(); torch.cuda.synchronize()
                    runs["graph"] = g.replay
                except Exception:  # noqa: BLE001
                    torch.cuda.synchronize()
                for how, run in runs.items():
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(); run(); e1.record(); torch.cuda.synchronize()
                    tot[how][i] += e0.elapsed_time(e1) / iters
            del x, w, y, dy, dx, dw
        names = ("forward", "input_gradient", "kernels_gradient")
        out[cfg] = {"batch": N, "layers": len(layers), "gflop_per_op": flops / 1e9,
                    "ms": {n: tot["issued"][i] for i, n in enumerate(names)},
                    "ms_cuda_graph": {n: tot["graph"][i] for i, n in enumerate(names)},
                    "tflops_cuda_graph": {n: (flops / (tot["graph"][i] * 1e-3) / 1e12 if tot["graph"][i] > 0 else None) for i, n in enumerate(names)},
                    "frac_of_tf32_peak_cuda_graph": {n: (flops / (tot["graph"][i] * 1e-3) / 1e12 / tf32_peak if tot["graph"][i] > 0 else None)
                                                     for i, n in enumerate(names)}}
    return out


def neighbour_pass(batch, iters=20):
    """The HBM-bound backward prologue next to the conv ops: dz = relu'(y)*dy and db = sum(dz) in one
    pass (nb200_conv2d_bias_activation_gradient) on VGG16 block1's activation (batch x 64 x 512 x 512). Algorithmic bytes =
    12 per element (read y, read dy, write dz)."""
    import torch
    from neuro__b200 import lib
    from neuro__b200.tensor_op import TensorOpB200
    op = TensorOpB200()
    y = torch.rand(batch, 64, 512, 512, device="cuda") - 0.5; dy = torch.rand_like(y); dz = torch.empty_like(y)
    db = torch.empty(64, device="cuda")
    for _ in range(3):
        op.Conv2DBiasActivationGradient(y, dy, lib.ACT_RELU, 0.0, dz, db)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        op.Conv2DBiasActivationGradient(y, dy, lib.ACT_RELU, 0.0, dz, db)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    peaks = read_peaks()
    gbs = 12.0 * y.numel() / (ms * 1e-3) / 1e9
    return {"kernel": "act_bias_gradient_kernel (relu'(y)*dy + bias gradient, one pass)", "ms": ms, "achieved_gbs": gbs,
            "peak_gbs": peaks["hbm_gbs"], "frac": gbs / peaks["hbm_gbs"], "bytes": 12.0 * y.numel()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="images per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the batch-1 style-transfer and neighbour-kernel passes")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
