"""Diagnostic: per-layer times of the three conv ops on the layer shapes of a BASELINE config.

  python tools/layer_probe.py dcgan|pix2pix|vgg1|autoenc [--prepared] [--graph]

Rows: conv y = conv(x[N,C,H,H], w[K,C,F,F], stride, pad). Conv2DTranspose layers are listed as the conv whose input
gradient is their forward. --prepared times forward / input gradient with filters prepared once (constant weights)."""
import sys
import torch
sys.path.insert(0, ".")
from neuro__b200 import lib  # noqa: E402
from neuro__b200.tensor_op import TensorOpB200  # noqa: E402

from neuro__b200.shapes import CONFIGS  # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "dcgan"
prepared = "--prepared" in sys.argv
graph = "--graph" in sys.argv
N, LAYERS = CONFIGS[name]
op = TensorOpB200(lib.MATH_TF32)
tot = [0.0, 0.0, 0.0]; totfl = 0.0
for (lname, C, H, K, F, st, p) in LAYERS:
    Ho = (H + 2 * p - F) // st + 1
    x = torch.randn(N, C, H, H, device="cuda"); w = torch.randn(K, C, F, F, device="cuda") * 0.05
    y = torch.empty(N, K, Ho, Ho, device="cuda"); dy = torch.randn_like(y); dx = torch.empty_like(x); dw = torch.empty_like(w)
    d = lib.ConvDesc(N, C, H, H, K, F, F, Ho, Ho, st, p, p, lib.NCHW, lib.MATH_TF32)
    pf = op.PrepareKernels(lib.OP_FORWARD, x, w, y, st, p, p) if prepared else None
    pg = op.PrepareKernels(lib.OP_INPUT_GRADIENT, dx, w, dy, st, p, p) if prepared else None
    fns = [lambda: op.Conv2D(x, w, st, p, p, lib.NCHW, y, prepared=pf), lambda: op.Conv2DInputGradient(dy, w, st, p, p, lib.NCHW, dx, prepared=pg),
           lambda: op.Conv2DKernelsGradient(x, dy, st, p, p, lib.NCHW, dw)]
    out = []
    for i, fn in enumerate(fns):
        fn(); fn(); torch.cuda.synchronize()
        run = lambda: [fn() for _ in range(10)]
        if graph:   # GPU time without the host's per-call cost (ctypes, tensor-map encode, launches): 10 calls replayed as one CUDA graph
            side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                run()
            run = gr.replay
            run(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tot[i] += ms
        out.append("%s %7.3f ms %6.1f TF/s %6.0f GB/s" % (op.kernel_name(i, d)[:20], ms, d.flops() / ms / 1e9, d.bytes() / ms / 1e6))
    totfl += d.flops()
    print("%-36s %6.2f GF | " % (lname, d.flops() / 1e9) + " | ".join(out), flush=True)
print("TOTAL %s N=%d prepared=%d graph=%d: fwd %.3f ms dgrad %.3f ms wgrad %.3f ms; %.1f GF per op -> %.1f / %.1f / %.1f TF/s" %
      (name, N, prepared, graph, tot[0], tot[1], tot[2], totfl / 1e9, totfl / tot[0] / 1e9, totfl / tot[1] / 1e9, totfl / tot[2] / 1e9))
