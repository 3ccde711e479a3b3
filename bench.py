#!/usr/bin/env python
"""bench.py -- the Conv2D hot path on the VGG16 shapes (BASELINE.json configs[2] / metric), on N B200s of one node.

A "step" is one training step of VGG16's convolutional stack (Neuro/src/Applications/VGG16.cpp:73-91: 13 convolutions
3x3 s1 p1 + bias + ReLU, a 2x2 max pooling after each of the 5 blocks) at 512x512x3 input and `--batch` images per GPU,
run by the SAME engine the loss-curve tests check (neuro__b200/fit.py: ConvStackTrainer), through the C ABI:
    forward    13 x Conv2DBiasActivation, 5 x Pool2D                                  nb200_conv2d_forward, nb200_pool2d
    loss       MSE against a synthetic target at the block5_pool output (torch, not on the path)
    backward   per layer, last to first: activation gradient + bias gradient (one pass), kernel gradient, input
               gradient (down to the IMAGE: style transfer reads that), pooling gradient
    exchange   bucketed SUM-all-reduce of the 14.7 M kernel + bias gradients over NCCL (N > 1), under the backward kernels
    update     one fused Adam step over the flat parameter buffer
Batch sharding = weak scaling: every rank holds `--batch` images. FLOPs counted: the three convolution ops only
(3 x 160.4 GFLOP per image).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for the meaning of every key.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# (C, K, H=W) of the 13 conv layers of VGG16 at 512x512 input: Neuro/src/Applications/VGG16.cpp:73-91
VGG16 = [(3, 64, 512), (64, 64, 512), (64, 128, 256), (128, 128, 256), (128, 256, 128), (256, 256, 128), (256, 256, 128),
         (256, 512, 64), (512, 512, 64), (512, 512, 64), (512, 512, 32), (512, 512, 32), (512, 512, 32)]
VGG_BLOCKS = {"vgg16": [2, 2, 3, 3, 3], "vgg19": [2, 2, 4, 4, 4]}      # convolutions per block (VGG16.cpp:73-91, VGG19.cpp:22-43)
VGG_WIDTH = [64, 128, 256, 512, 512]
F, STRIDE, PAD = 3, 1, 1
MIN_TIMED_SECONDS = 2.0      # the timed region is looped until it is at least this long, whatever --steps says
CPU_SAMPLE_RES = 32          # ONE bounded CPU sample for both legs: the same stack at batch 1, 32x32x3 input


def layer_flops(C, K, HW, batch=1):
    return 2.0 * batch * K * HW * HW * C * F * F


SAMPLE_FLOPS_PER_OP = sum(layer_flops(*l) for l in VGG16)   # 160.4 GFLOP per image for fwd (= dgrad = wgrad)


def vgg_layers(name="vgg16"):
    from neuro__b200 import lib
    from neuro__b200.fit import ConvLayerSpec, PoolSpec
    out = []
    for n, k in zip(VGG_BLOCKS[name], VGG_WIDTH):
        out += [ConvLayerSpec(k, 3, 1, 1, lib.ACT_RELU) for _ in range(n)]
        out.append(PoolSpec(2, 2, 0, lib.POOL_MAX))
    return out


def vgg_flops_per_image(name, res):
    fl, C, HW = 0.0, 3, res
    for n, k in zip(VGG_BLOCKS[name], VGG_WIDTH):
        for _ in range(n):
            fl += 2.0 * k * HW * HW * C * 9; C = k
        HW //= 2
    return fl


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


def measure_tf32_peak():
    """TF32 tensor-core peak of THIS box, measured by tools/tf32_peak (back-to-back tcgen05.mma kind::tf32, burst and 4 s
    sustained; MEASURED_PEAKS.json has no TF32 entry). Falls back to half the measured bf16 rate if the tool is missing."""
    exe = os.path.join(ROOT, "tools", "tf32_peak")
    try:
        out = subprocess.run([exe, "3.0"], capture_output=True, text=True, timeout=60).stdout.strip().splitlines()[-1]
        d = json.loads(out)
        if "tf32_tflops_burst" in d:
            return {"burst": d["tf32_tflops_burst"], "sustained": d["tf32_tflops_settled"], "source": "tools/tf32_peak on this box: " + d["how"]}
    except Exception:  # noqa: BLE001
        pass
    p = read_peaks()
    return {"burst": p["bf16_burst"] / 2.0, "sustained": p["bf16_sustained"] / 2.0,
            "source": "%s MEASURED_PEAKS.json bf16 / 2 (tools/tf32_peak unavailable)" % p["source"]}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe): one looping nvidia-smi
    process (-lms 50) started before and terminated after the region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

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

def cpu_reference_run(steps, warmup, res=CPU_SAMPLE_RES):
    """Times the reference's own CPU implementation (oracle/_ref = TensorOpCpuMt and the pooling / activation-gradient loops
    compiled from the reference sources; else the oracle port) on a bounded sample: the same training step, batch 1, at
    `res` x `res` input, with every host thread the reference's parallel_for can use."""
    import numpy as np
    from neuro__b200 import synth
    from oracle import oracle as O

    use_ref = O.have_ref()
    cores = os.cpu_count() or 1
    if use_ref:
        O.ref_set_threads(cores)           # explicit: torchrun exports OMP_NUM_THREADS=1
        cores = O.ref_threads()
    else:
        O.build(ref=False)
    conv = (lambda x, w: O.ref_conv2d(x, w, 1, 1, 1, mt=True)) if use_ref else (lambda x, w: O.conv2d(x, w, 1, 1, 1))
    dgrad = (lambda dy, w, hw: O.ref_conv2d_input_gradient(dy, w, 1, 1, 1, hw, mt=True)) if use_ref else (lambda dy, w, hw: O.conv2d_input_gradient(dy, w, 1, 1, 1, hw))
    wgrad = (lambda x, dy: O.ref_conv2d_kernels_gradient(x, dy, 1, 1, 1, (3, 3), mt=True)) if use_ref else (lambda x, dy: O.conv2d_kernels_gradient(x, dy, 1, 1, 1, (3, 3)))
    pool = O.ref_pool2d if use_ref else O.pool2d
    poolg = O.ref_pool2d_gradient if use_ref else O.pool2d_gradient
    actg = O.ref_activation_gradient if use_ref else O.activation_gradient
    frac = vgg_flops_per_image("vgg16", res) / SAMPLE_FLOPS_PER_OP
    ws, C = [], 3
    for i, (n, k) in enumerate(zip(VGG_BLOCKS["vgg16"], VGG_WIDTH)):
        for j in range(n):
            ws.append(synth.glorot_uniform(synth.SEED_W + len(ws), k, C, 3, 3)); C = k
    x0 = synth.uniform(synth.SEED_X, (1, 3, res, res))

    def step():
        acts, pools, x, li = [], [], x0, 0
        for n, k in zip(VGG_BLOCKS["vgg16"], VGG_WIDTH):
            for _ in range(n):
                y = np.maximum(conv(x, ws[li]), 0); acts.append((x, y, li)); x = y; li += 1   # bias = 0 (Conv2D.h:46), ReLU
            p = pool(x, 2, 2, 0); pools.append((x, p)); acts.append(None); x = p
        g = (x * np.float32(2.0 / x.size)).astype(np.float32)
        for a in reversed(acts):
            if a is None:
                xin, p = pools.pop(); g = poolg(p, xin, g, 2, 2, 0)
                continue
            xin, y, li = a
            dz = actg(2, 0.0, y, g)            # EActivation _ReLU
            wgrad(xin, dz)
            g = dgrad(dz, ws[li], xin.shape[2:])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    value = frac / dt    # equivalent full-resolution samples per second (FLOP-linear extrapolation of the sample)
    sample = ("VGG16 conv-stack training step (13 conv fwd/dgrad/wgrad + ReLU gradient + 5 max-pool fwd/bwd), batch 1, %dx%dx3 input "
              "(%.5f of the 481.2 GFLOP per 512x512 sample), samples/s extrapolated linearly in FLOPs; %s, %d threads set explicitly"
              % (res, res, frac, "reference TensorOpCpuMt (PPL->OpenMP; dgrad parallel over N only, wgrad over K only)" if use_ref
                 else "oracle port (OpenMP)", cores))
    return {"value": value, "unit": "samples/s", "cores": cores, "kind": "reference" if use_ref else "port", "sample": sample,
            "seconds_per_step": dt, "gflops": 3 * frac * SAMPLE_FLOPS_PER_OP / dt / 1e9}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_run(args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "vgg16_conv_fwd_dgrad_wgrad_samples_per_s", "value": cb["value"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": cb["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch, args.gpus),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(batch, gpus):
    return {"workload": "VGG16 conv stack @512x512x3, training step: 13 x (conv 3x3 s1 p1 + bias + ReLU) and 5 x max-pool forward; "
                        "ReLU/bias gradient, kernel gradient, input gradient (to the image), pool gradient backward; bucketed gradient "
                        "all-reduce; Adam",
            "per_gpu_batch": batch, "global_batch": batch * gpus, "parallelism": "dp%d (batch sharded)" % gpus,
            "math": "tf32 tensor cores, fp32 accumulate (first layer C=3: HBM-bound; forward / input gradient on fp32 CUDA cores, kernel gradient on tensor cores)",
            "l2": "inputs larger than L2: each step streams every layer's tensors once (%.1f GB of distinct tensors per step)" % (batch * 1.2)}


# ------------------------------------------------------------------------------------------------ our arm

class TimedOp:
    """TensorOpB200 proxy: when `events` is a list, every op call is bracketed by a CUDA-event pair on the launching stream."""

    def __init__(self, op):
        self._op, self.events, self._tagv = op, None, None

    def set_tag(self, i):
        self._tagv = i

    def __getattr__(self, name):
        fn = getattr(self._op, name)
        if not callable(fn) or name.startswith("_") or name in ("kernel_name", "PrepareKernels"):
            return fn

        def call(*a, **k):
            if self.events is None:
                return fn(*a, **k)
            import torch
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); r = fn(*a, **k); e1.record()
            self.events.append((name, self._tagv, e0, e1))
            return r
        return call


KERNEL_OF = {"tcgen05_fprop": "tc_fprop_kernel (forward + input gradient, >64 filters)",
             "tcgen05_dgrad": "tc_fprop_kernel (forward + input gradient, >64 filters)",
             "tcgen05_rowtap_fprop": "tc_rowtap_kernel (forward + input gradient, <=64 filters)",
             "tcgen05_rowtap_dgrad": "tc_rowtap_kernel (forward + input gradient, <=64 filters)",
             "tcgen05_wgrad": "tc_wgrad_kernel (kernel gradient)",
             "tcgen05_rowfold_wgrad": "tc_wgrad_rowfold_kernel (kernel gradient, <=64 channels and filters)",
             "tcgen05_smallc_wgrad": "tc_smallc_wgrad_kernel (kernel gradient, <=4 channels, HBM-bound)"}
OP_OF = {"Conv2DBiasActivation": 0, "Conv2DInputGradient": 1, "Conv2DKernelsGradient": 2}


def init_dist():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner / debug lines must not share stdout with the JSON line
        # The exchange is 59 MB per ~10 ms step: bandwidth is irrelevant, but every NCCL CTA takes an SM away from the
        # one-CTA-per-SM persistent convolution grids it runs under (+12 % on the input gradient at N = 8).
        os.environ.setdefault("NCCL_MAX_CTAS", "4")
        import datetime
        # a rank that fails alone must not leave the others hanging for NCCL's default 10 minutes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    return rank, world, local


def run_ours(args):
    import torch
    import torch.distributed as dist
    from neuro__b200 import lib, synth
    from neuro__b200.fit import ConvLayerSpec, ConvStackTrainer
    from neuro__b200.tensor_op import TensorOpB200

    rank, world, local = init_dist()
    L = lib.load()
    dev = torch.device("cuda", local)
    B = args.batch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # TF32 peak of this box (rank 0 measures, before anything else heats the chip; the others wait at the barrier)
    tf32 = measure_tf32_peak() if rank == 0 else None
    barrier()

    op = TimedOp(TensorOpB200(lib.MATH_TF32))
    # --graph: forward + backward + exchange replayed as ONE CUDA graph (the trainer's public switch; bit-identical to the eager step,
    # tests/test_fit_gpu.py, tests/test_fit_nccl_gpu.py) for the headline and e2e passes; the instrumented pass issues call by call.
    # Off by default: at batch 8 the GPU is never starved by the host and the replayed step measured no faster (9.82 vs 9.64 ms,
    # within the +-2 % of the power cap); it is what the small-layer configs in other_configs gain from.
    tr = ConvStackTrainer(op, (3, 512, 512), vgg_layers("vgg16"), dev, optimizer="adam", lr=1e-5, seed=synth.SEED_MODEL,
                          world_size=world, rank=rank, input_gradient=True, use_graph=args.graph)
    gen = torch.Generator(device=dev); gen.manual_seed(1000 + rank)       # a different data shard per rank
    x_dev = torch.rand(B, 3, 512, 512, device=dev, generator=gen) * 2 - 1   # U(-1,1) (Tensor::FillWithRand default)
    t_dev = torch.rand(B, *tr.out_shape, device=dev, generator=gen)
    tr.load_batch(x_dev, t_dev)
    conv_idx = [i for i, l in enumerate(tr.layers) if isinstance(l, ConvLayerSpec)]
    descs = {}
    for i in conv_idx:
        C, H, W, K, Ho, Wo = tr.shapes[i]
        descs[i] = lib.ConvDesc(B, C, H, W, K, F, F, Ho, Wo, STRIDE, PAD, PAD, lib.NCHW, lib.MATH_TF32)
    names = {i: [tr.op.kernel_name(o, descs[i]) for o in (0, 1, 2)] for i in conv_idx}
    global_batch = B * world

    for _ in range(max(args.warmup, 3)):
        tr.run_step(global_batch)
    barrier()

    # ---- timed region (device-resident inputs): blocks of --steps steps, repeated until >= MIN_TIMED_SECONDS ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.15)   # let the first samples land inside the region
    launches0 = L.nb200_kernel_launches()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); tr.run_step(global_batch); e1.record(); torch.cuda.synchronize()
    est = max(e0.elapsed_time(e1) * 1e-3, 1e-4)
    rounds = max(1, int(math.ceil(MIN_TIMED_SECONDS / (est * args.steps))))
    if world > 1:   # every rank must run the same number of steps
        r = torch.tensor([rounds], device=dev); dist.all_reduce(r, op=dist.ReduceOp.MAX); rounds = int(r.item())
    timed_steps = rounds * args.steps
    launches0 = L.nb200_kernel_launches()
    barrier()
    t0.record()
    for _ in range(timed_steps):
        tr.run_step(global_batch)
    t1.record()
    barrier()
    # the library counts the kernels it launches; a replayed graph launches tr.graph_kernel_launches of them per step on top
    launches = (L.nb200_kernel_launches() - launches0 + (timed_steps * tr.graph_kernel_launches if tr.use_graph else 0)) // rounds   # per block of --steps steps
    ms = t0.elapsed_time(t1) / timed_steps
    if world > 1:
        tmax = torch.tensor([ms], device=dev); dist.all_reduce(tmax, op=dist.ReduceOp.MAX); ms = float(tmax.item())
    # second timed pass with a CUDA-event pair around every op call (feeds per_op and roofline; the ~100 extra event
    # records per step cost ~1 %, which is why the headline pass above runs without them)
    op.events = []
    op_steps = max(3, min(args.steps, 20))
    graph_mode, tr.use_graph = tr.use_graph, False
    barrier()
    for _ in range(op_steps):
        tr.run_step(global_batch)
    barrier()
    events, op.events = op.events, None
    tr.use_graph = graph_mode
    if sampler:
        sampler.stop()
    fam_ms = {0: 0.0, 1: 0.0, 2: 0.0}
    other_ms = {}
    tc_ms = {0: 0.0, 1: 0.0, 2: 0.0}; tc_fl = {0: 0.0, 1: 0.0, 2: 0.0}
    kern, kern_bytes = {}, {}   # CUDA kernel -> [flops, ms, launches] over the instrumented pass
    total_ms = 0.0
    for name, i, a, b in events:
        dt = a.elapsed_time(b)
        total_ms += dt
        fam = OP_OF.get(name)
        if fam is None:
            other_ms[name] = other_ms.get(name, 0.0) + dt
            continue
        fam_ms[fam] += dt
        kn = names[i][fam]
        if kn.startswith("tcgen05") and "smallc" not in kn:   # tensor-bound launches only
            tc_ms[fam] += dt; tc_fl[fam] += descs[i].flops()
        key = KERNEL_OF.get(kn, kn)
        k = kern.setdefault(key, [0.0, 0.0, 0])
        k[0] += descs[i].flops(); k[1] += dt; k[2] += 1
        kern_bytes[key] = kern_bytes.get(key, 0.0) + descs[i].bytes()

    # ---- e2e: the same step through the public API with HOST input and HOST result, copies inside the timed region ----
    # Every step's input batch comes from pinned host memory and its result (the image gradient, what style transfer reads
    # back) returns to host memory. As in any input pipeline (the reference has DataPreloader for this) the H2D copy of
    # step i+1 and the D2H copy of step i run on a copy stream underneath step i+1 / i's kernels; all of it is inside the
    # timed region and the region ends only when the last result has landed on the host.
    img_host = torch.from_numpy(synth.uniform(synth.SEED_X, (B, 3, 512, 512))).pin_memory()
    grad_host = [torch.empty((B, 3, 512, 512), dtype=torch.float32).pin_memory() for _ in range(2)]
    stage = [torch.empty_like(tr.x_in) for _ in range(2)]        # H2D landing buffers; the step's first kernel reads tr.x_in
    dstage = [torch.empty_like(tr.dx_in) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    main = torch.cuda.current_stream()

    def e2e_run(n_steps):
        h2d = [torch.cuda.Event(), torch.cuda.Event()]
        taken = [torch.cuda.Event(), torch.cuda.Event()]     # main has moved stage[j] into the step's input / filled dstage[j]
        d2h = [torch.cuda.Event(), torch.cuda.Event()]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_stream(main)
            stage[0].copy_(img_host, non_blocking=True); h2d[0].record(copy_stream)
        for i in range(n_steps):
            j = i & 1
            main.wait_event(h2d[j])
            tr.x_in.copy_(stage[j], non_blocking=True)        # device-to-device, 25 MB: the static input buffer of the step
            if i + 1 < n_steps:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(taken[j ^ 1])
                    stage[j ^ 1].copy_(img_host, non_blocking=True); h2d[j ^ 1].record(copy_stream)
            tr.run_step(global_batch)
            if i >= 2:
                main.wait_event(d2h[j])                       # dstage[j] has been drained to the host
            dstage[j].copy_(tr.dx_in, non_blocking=True)
            taken[j].record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(taken[j])
                grad_host[j].copy_(dstage[j], non_blocking=True); d2h[j].record(copy_stream)
        main.wait_stream(copy_stream)

    e2e_run(2); barrier()
    s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
    # (from the max-over-ranks step time: every rank must run the same number of steps -- they contain collectives)
    e2e_steps = max(3, min(timed_steps, int(math.ceil(MIN_TIMED_SECONDS / (ms * 1e-3)))))
    s0.record()
    e2e_run(e2e_steps)
    s1.record()
    barrier()
    e2e_ms = s0.elapsed_time(s1) / e2e_steps
    if world > 1:
        tmax = torch.tensor([e2e_ms], device=dev); dist.all_reduce(tmax, op=dist.ReduceOp.MAX); e2e_ms = float(tmax.item())

    line = None
    if rank == 0:
        peaks = read_peaks()
        clocks = sampler.summary() if sampler else None
        # burst or sustained denominator: by the SM clock sampled during the timed region
        burst_like = bool(clocks and clocks["sm_mhz"] and clocks["sm_max_mhz"] and clocks["sm_mhz"] >= 0.93 * clocks["sm_max_mhz"])
        tf32_peak = tf32["burst"] if burst_like else tf32["sustained"]
        # kernel level: the op-level event pairs are grouped by the CUDA kernel that op dispatched to (an op's pair also
        # covers its helper launches: the filter repack before tc_fprop/tc_rowtap, the split-K reduce after tc_wgrad; < 3 %)
        dom_name = max(kern, key=lambda k: kern[k][1])
        dom_fl, dom_ms, dom_n = kern[dom_name]
        achieved = dom_fl / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        # DRAM traffic of the dominant kernel: from the committed ncu --set full capture of the same workload (never measured
        # in this run: a profiler run is not a timing run); null when the capture has no row for the kernel
        traffic, traffic_src = None, None
        for fn in ("r2_dram_traffic_per_launch.json", "r1_dram_traffic_per_launch.json"):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", fn)))
                ent = tj["kernels"].get(dom_name.split(" ")[0])
                if ent and B == 8:
                    traffic, traffic_src = ent["dram_bytes_per_launch"], tj["source"]
                    break
            except (OSError, ValueError, KeyError):
                pass
        conv_ms = sum(fam_ms.values())
        kernels = {k: {"ms_per_step": v[1] / op_steps, "launches_per_step": v[2] // op_steps,
                       "tflops": v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else None,
                       "share_of_step": v[1] / total_ms} for k, v in kern.items()}
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
            "timed_steps": timed_steps, "timed_seconds": ms * 1e-3 * timed_steps,
            "tflops_total": 3 * B * world * SAMPLE_FLOPS_PER_OP / (ms * 1e-3) / 1e12,
            "per_op": per_op, "kernels": kernels,
            "conv_ops_ms_per_step": conv_ms / op_steps,
            "neighbour_ops_ms_per_step": {k: v / op_steps for k, v in sorted(other_ms.items())},
            "tf32_peak": {"burst_tflops": tf32["burst"], "sustained_tflops": tf32["sustained"], "used": "burst" if burst_like else "sustained",
                          "source": tf32["source"]},
            "roofline": {"bound": "tensor", "kernel": dom_name, "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                         "frac": achieved / tf32_peak if tf32_peak else None, "traffic": traffic,
                         "traffic_unit": "DRAM bytes per launch (read + write)", "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": kern_bytes.get(dom_name, 0.0) / dom_n if dom_n else None,
                         "avg_launch_ms": dom_ms / dom_n if dom_n else None,
                         "flops_per_launch": dom_fl / dom_n if dom_n else None,
                         "peak_source": "TF32 peak measured on this box by tools/tf32_peak (%s: SM clock sampled in the timed region %s MHz of %s)"
                                        % ("burst" if burst_like else "sustained/settled", clocks and clocks["sm_mhz"], clocks and clocks["sm_max_mhz"]),
                         "share_of_step": dom_ms / total_ms if total_ms else None},
            "e2e": {"value": B * world / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": img_host.numel() * 4,
                    "d2h_bytes_per_step": grad_host[0].numel() * 4, "ms_per_step": e2e_ms, "steps": e2e_steps},
            "gpu_launches": int(launches), "step_issue": "cuda_graph" if tr.use_graph else "eager",
            "exchange": {"buckets": len(tr.buckets), "bucket_bytes": [4 * (hi - lo) for lo, hi, _ in tr.buckets],
                         "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS")} if world > 1 else None,
            "clocks": clocks,
        }
    del tr, x_dev, t_dev, stage, dstage
    torch.cuda.empty_cache()
    if not args.no_extras:
        # the other BASELINE configs as model steps at THIS world size (every rank takes part; rank 0 reports)
        try:
            extra = other_configs_pass(rank, world, dev)
            if line is not None:
                line["other_configs"] = extra
                for k in ("vgg16_style_transfer", "vgg19_4k_tiles"):
                    if k in extra and extra[k].get("tflops_per_gpu"):
                        extra[k]["frac_of_tf32_peak"] = extra[k]["tflops_per_gpu"] / line["roofline"]["peak"]
        except Exception as e:  # noqa: BLE001 -- whatever happens here, the headline line is still printed
            if line is not None:
                line["extras_error"] = str(e)[:300]
    if rank == 0:
        if world == 1 and not args.no_extras:
            try:
                line["neighbours"] = neighbour_pass(B)
            except Exception as e:  # noqa: BLE001
                line["extras_error"] = (line.get("extras_error", "") + " | " + str(e))[:300]
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_run(2, 1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ the other BASELINE configs

def model_stacks():
    """Sequential conv stacks of BASELINE configs 1, 2 and 4 as ConvStackTrainer layer lists: (input shape, layers, per-GPU batch)."""
    from neuro__b200 import lib
    from neuro__b200.fit import ConvLayerSpec as Cv, PoolSpec, UpSampleSpec
    lrelu = dict(activation=lib.ACT_LEAKY_RELU, alpha=0.2)
    return {
        # conv autoencoder, batch 256 (Neuro.Examples/include/ConvAutoencoderNetwork.h:25-35)
        "autoencoder": ((1, 28, 28), [Cv(16, 3, 1, 1), PoolSpec(2, 2), Cv(8, 3, 1, 1), PoolSpec(2, 2), Cv(8, 3, 1, 1), UpSampleSpec(2),
                                      Cv(16, 3, 1, 1), UpSampleSpec(2), Cv(1, 3, 1, 1, lib.ACT_SIGMOID)], 256),
        # DCGAN on CIFAR shapes, batch 128 (Neuro.Examples/src/CifarGAN.cpp:11-36): discriminator convs; generator from the 256x4x4 seed
        "dcgan_discriminator": ((3, 32, 32), [Cv(64, 3, 2, 1, **lrelu), Cv(128, 3, 2, 1, **lrelu), Cv(128, 3, 2, 1, **lrelu), Cv(256, 3, 1, 1, **lrelu)], 128),
        "dcgan_generator": ((256, 4, 4), [Cv(128, 4, 2, 1, transposed=True, **lrelu), Cv(128, 4, 2, 1, transposed=True, **lrelu),
                                          Cv(128, 4, 2, 1, transposed=True, **lrelu), Cv(3, 3, 1, 1, lib.ACT_TANH)], 128),
        # DeepConvGAN on MNIST shapes: the batch-normalised variant (Neuro.Examples/src/DeepConvGAN.cpp:3-43), statistics over the global batch
        "dcgan_bn_discriminator": ((1, 28, 28), [Cv(32, 3, 2, 1, batch_norm=True, **lrelu), Cv(64, 3, 2, 1, batch_norm=True, **lrelu),
                                                 Cv(128, 3, 2, 1, batch_norm=True, **lrelu), Cv(256, 3, 1, 1, batch_norm=True, **lrelu)], 128),
        # pix2pix PatchGAN discriminator, 256x256 pairs, batch 8 (Neuro.Examples/src/Pix2Pix.cpp:75-109; ZeroPadding2D folded into the extents)
        # (the reference pads 3 / 2 pixels on one side before each 4x4 pad-0 conv; here pad 1 on both sides: same extents 128, 64, 32, 31, 30)
        "pix2pix_patchgan": ((6, 259, 259), [Cv(64, 4, 2, 0, **lrelu), Cv(128, 4, 2, 1, batch_norm=True, **lrelu), Cv(256, 4, 2, 1, batch_norm=True, **lrelu),
                                             Cv(512, 4, 1, 1, batch_norm=True, **lrelu), Cv(1, 4, 1, 1, lib.ACT_SIGMOID)], 8),
    }


def stack_flops(in_shape, layers):
    from neuro__b200.fit import ConvLayerSpec, PoolSpec
    C, H, W = in_shape
    fl = 0.0
    for l in layers:
        if isinstance(l, ConvLayerSpec):
            if l.transposed:
                Ho = (H - 1) * l.stride + l.filter_size - 2 * l.padding
                fl += 2.0 * C * H * W * l.filters * l.filter_size ** 2
            else:
                Ho = (H + 2 * l.padding - l.filter_size) // l.stride + 1
                fl += 2.0 * l.filters * Ho * Ho * C * l.filter_size ** 2
            C, H, W = l.filters, Ho, Ho
        elif isinstance(l, PoolSpec):
            H = W = (H + 2 * l.padding - l.filter_size) // l.stride + 1
        else:
            H = W = H * l.scale
    return fl


def time_steps(fn, dev, world, min_seconds=0.5, min_steps=10):
    import torch
    import torch.distributed as dist
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    n = max(min_steps, int(math.ceil(min_seconds / max(e0.elapsed_time(e1) * 1e-3, 1e-5))))
    if world > 1:
        t = torch.tensor([n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); n = int(t.item())
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    if world > 1:
        t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    return ms, n


def other_configs_pass(rank, world, dev):
    """Model throughput of the other BASELINE configs at this world size (weak scaling, batch sharded, max over ranks):
      * training steps of the sequential conv stacks of configs 1, 2, 4 (ConvStackTrainer, whole step replayed as a CUDA graph;
        gradient all-reduce + global-batch batch-norm statistics when world > 1);
      * pix2pix U-Net generator (skip connections: not a sequential stack) as the per-layer conv-op step over its layer table
        (forward, activation/bias gradient, kernel gradient, input gradient on resident tensors) + one flat gradient all-reduce;
      * configs[2] as the reference runs it: VGG16 @512, ONE image per GPU, frozen weights, forward + gradient back to the image;
      * configs[4]: VGG19 on a 4096x4096 image cut into 64 tiles of 512x512, 8 tiles per GPU per step, forward + image gradient
        (independent tiles: replicas, no collective on the data path)."""
    import torch
    import torch.distributed as dist
    from neuro__b200 import lib, synth
    from neuro__b200.fit import ConvStackTrainer
    from neuro__b200.shapes import CONFIGS
    from neuro__b200.tensor_op import TensorOpB200
    out = {}
    gen = torch.Generator(device=dev); gen.manual_seed(77 + rank)
    for name, (in_shape, layers, batch) in model_stacks().items():
        res = {}
        for how in ("cuda_graph", "issued"):
            tr = ConvStackTrainer(TensorOpB200(lib.MATH_TF32), in_shape, layers, dev, optimizer="adam", lr=1e-4, seed=synth.SEED_MODEL,
                                  world_size=world, rank=rank, use_graph=how == "cuda_graph", input_gradient=name.endswith("discriminator"))
            tr.load_batch(torch.rand(batch, *in_shape, device=dev, generator=gen) * 2 - 1, torch.rand(batch, *tr.out_shape, device=dev, generator=gen))
            try:
                ms, n = time_steps(lambda: tr.run_step(batch * world), dev, world)
                res[how] = ms
            except Exception as e:  # noqa: BLE001
                res[how] = None; res[how + "_error"] = str(e)[:160]
                torch.cuda.synchronize()
            del tr
        best = min(v for v in (res.get("cuda_graph"), res.get("issued")) if v)
        fl = 3 * stack_flops(in_shape, layers) * batch
        out[name] = {"per_gpu_batch": batch, "ms_per_step": res, "samples_per_s": batch * world / (best * 1e-3),
                     "conv_tflops_per_gpu": fl / (best * 1e-3) / 1e12}
    # pix2pix U-Net generator: per-layer conv-op step over the layer table
    N, table = CONFIGS["pix2pix"]
    op = TensorOpB200(lib.MATH_TF32)
    gl = [r for r in table if r[0].startswith("G ")]
    tens, nparam, flops = [], 0, 0.0
    for (_, C, H, K, Fs, st, pd) in gl:
        Ho = (H + 2 * pd - Fs) // st + 1
        nparam += K * C * Fs * Fs + K
        flops += 3 * 2.0 * N * K * Ho * Ho * C * Fs * Fs
    params = torch.zeros(nparam, device=dev); grads = torch.zeros(nparam, device=dev); m = torch.zeros(nparam, device=dev); v = torch.zeros(nparam, device=dev)
    off = 0
    for (_, C, H, K, Fs, st, pd) in gl:
        Ho = (H + 2 * pd - Fs) // st + 1
        nw = K * C * Fs * Fs
        w = params[off:off + nw].view(K, C, Fs, Fs); w.copy_(torch.randn(K, C, Fs, Fs, device=dev, generator=gen) * 0.05)
        tens.append(dict(x=torch.randn(N, C, H, H, device=dev, generator=gen), w=w, b=params[off + nw:off + nw + K], dw=grads[off:off + nw].view(K, C, Fs, Fs),
                         db=grads[off + nw:off + nw + K], y=torch.empty(N, K, Ho, Ho, device=dev), dy=torch.randn(N, K, Ho, Ho, device=dev, generator=gen),
                         dz=torch.empty(N, K, Ho, Ho, device=dev), dx=torch.empty(N, C, H, H, device=dev), st=st, pd=pd))
        off += nw + K

    def unet_ops():
        for t in tens:
            op.Conv2DBiasActivation(t["x"], t["w"], t["st"], t["pd"], t["pd"], t["b"], lib.ACT_LEAKY_RELU, 0.2, t["y"])
        for t in reversed(tens):
            op.Conv2DBiasActivationGradient(t["y"], t["dy"], lib.ACT_LEAKY_RELU, 0.2, t["dz"], t["db"])
            op.Conv2DKernelsGradient(t["x"], t["dz"], t["st"], t["pd"], t["pd"], lib.NCHW, t["dw"])
            op.Conv2DInputGradient(t["dz"], t["w"], t["st"], t["pd"], t["pd"], lib.NCHW, t["dx"])
        if world > 1:
            dist.all_reduce(grads, op=dist.ReduceOp.SUM)
    try:
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            unet_ops(); unet_ops()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            unet_ops()
        step = lambda: (g.replay(), op.AdamStep(params, grads, m, v, 1e-4, 0.9, 0.999, 1e-8, gradScale=1.0 / world))
        ms, n = time_steps(step, dev, world)
        out["pix2pix_unet_generator"] = {"per_gpu_batch": N, "ms_per_step": {"cuda_graph": ms}, "samples_per_s": N * world / (ms * 1e-3),
                                         "conv_tflops_per_gpu": flops / (ms * 1e-3) / 1e12, "note": "per-layer conv-op step over the layer table (skip connections are not a sequential stack)"}
    except Exception as e:  # noqa: BLE001
        out["pix2pix_unet_generator"] = {"error": str(e)[:200]}
        torch.cuda.synchronize()
    del tens, params, grads, m, v
    torch.cuda.empty_cache()
    # style transfer (configs[2]) and 4K tiles (configs[4]): frozen weights, forward + image gradient, independent images per GPU
    for key, model, batch, note in (("vgg16_style_transfer", "vgg16", 1, "VGG16 @512x512x3, 1 image per GPU per iteration: 13 x forward(bias+ReLU) + pools + input gradients to the image, frozen weights"),
                                    ("vgg19_4k_tiles", "vgg19", 8, "VGG19 on a 4096x4096x3 image as 64 tiles of 512x512 (VGG19.cpp:22-43), 8 tiles per GPU per step (8 GPUs = one whole image per step): forward + image gradient")):
        try:
            tr = ConvStackTrainer(TensorOpB200(lib.MATH_TF32), (3, 512, 512), vgg_layers(model), dev, seed=synth.SEED_MODEL, world_size=1, rank=0, input_gradient=True)
            tr.load_batch(torch.rand(batch, 3, 512, 512, device=dev, generator=gen) * 255 - 116, torch.rand(batch, *tr.out_shape, device=dev, generator=gen))
            res = {}
            ms, n = time_steps(lambda: tr.image_gradient_step(batch), dev, world)
            res["issued"] = ms
            try:
                side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    tr.image_gradient_step(batch)
                torch.cuda.current_stream().wait_stream(side)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    tr.image_gradient_step(batch)
                res["cuda_graph"], n = time_steps(g.replay, dev, world)
            except Exception as e:  # noqa: BLE001
                res["cuda_graph"] = None; res["cuda_graph_error"] = str(e)[:160]
                torch.cuda.synchronize()
            best = min(v for v in res.values() if isinstance(v, float))
            fl = 2 * vgg_flops_per_image(model, 512) * batch
            out[key] = {"workload": note, "images_per_gpu_per_step": batch, "ms_per_step": res, "ms_per_image": best / batch,
                        "images_per_s": batch * world / (best * 1e-3), "tflops_per_gpu": fl / (best * 1e-3) / 1e12,
                        "sharding": "independent images / tiles per GPU, no data-path collective (replicas)"}
            if key == "vgg19_4k_tiles":
                out[key]["seconds_per_4k_image_iteration"] = 64.0 / (batch * world) * best * 1e-3
            del tr
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out[key] = {"error": str(e)[:200]}
            torch.cuda.synchronize()
    return out


def neighbour_pass(batch, iters=20):
    """The HBM-bound backward prologue next to the conv ops: dz = relu'(y)*dy and db = sum(dz) in one
    pass (nb200_conv2d_bias_activation_gradient) on VGG16 block1's activation (batch x 64 x 512 x 512). Algorithmic bytes =
    12 per element (read y, read dy, write dz). Plus batch-norm training forward (12 B/element) and gradient (20 B/element)."""
    import torch
    from neuro__b200 import lib
    from neuro__b200.tensor_op import TensorOpB200
    op = TensorOpB200()
    peaks = read_peaks()
    y = torch.rand(batch, 64, 512, 512, device="cuda") - 0.5; dy = torch.rand_like(y); dz = torch.empty_like(y)
    db = torch.empty(64, device="cuda")

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    ms = timed(lambda: op.Conv2DBiasActivationGradient(y, dy, lib.ACT_RELU, 0.0, dz, db))
    gbs = 12.0 * y.numel() / (ms * 1e-3) / 1e9
    out = {"kernel": "act_bias_gradient_kernel (relu'(y)*dy + bias gradient, one pass)", "ms": ms, "achieved_gbs": gbs,
           "peak_gbs": peaks["hbm_gbs"], "frac": gbs / peaks["hbm_gbs"], "bytes": 12.0 * y.numel()}
    g = torch.ones(64, device="cuda"); b = torch.zeros(64, device="cuda"); sm = torch.empty(64, device="cuda"); sv = torch.empty(64, device="cuda")
    rm = torch.zeros(64, device="cuda"); rv = torch.ones(64, device="cuda")
    ms_f = timed(lambda: op.BatchNormalizationTrain(y, lib.BN_SPATIAL, g, b, 0.99, 1e-3, rm, rv, sm, sv, dz))
    ms_b = timed(lambda: op.BatchNormalizationGradient(y, lib.BN_SPATIAL, g, 1e-3, dy, sm, sv, rm, rv, True, dz))
    out["batch_norm"] = {"train_forward_ms": ms_f, "train_forward_gbs": 12.0 * y.numel() / (ms_f * 1e-3) / 1e9,
                         "gradient_ms": ms_b, "gradient_gbs": 20.0 * y.numel() / (ms_b * 1e-3) / 1e9,
                         "frac_forward": 12.0 * y.numel() / (ms_f * 1e-3) / 1e9 / peaks["hbm_gbs"],
                         "frac_gradient": 20.0 * y.numel() / (ms_b * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    return out


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else printed during the run (NCCL's version banner,
    library warnings) was redirected to stderr in main()."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n"); out.flush()


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)       # fd 1 -> stderr for native libraries and child processes
    sys.stdout = sys.stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="images per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs and the neighbour-kernel passes")
    ap.add_argument("--graph", action="store_true", help="replay the step as one CUDA graph instead of issuing it call by call")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
