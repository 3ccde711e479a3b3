"""Diagnostic: time fwd/dgrad/wgrad per layer shape with CUDA events (L2 flushed between iterations)."""
import sys
import torch

sys.path.insert(0, ".")
from neuro__b200 import lib  # noqa: E402
from neuro__b200.tensor_op import TensorOpB200  # noqa: E402

VGG16 = [(3, 64, 512), (64, 64, 512), (64, 128, 256), (128, 128, 256), (128, 256, 128), (256, 256, 128), (256, 256, 128),
         (256, 512, 64), (512, 512, 64), (512, 512, 64), (512, 512, 32), (512, 512, 32), (512, 512, 32)]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
math = int(sys.argv[2]) if len(sys.argv) > 2 else lib.MATH_TF32
iters = 5
op = TensorOpB200(math)
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
tot = {0: 0.0, 1: 0.0, 2: 0.0}; totf = 0.0
for (C, K, HW) in VGG16:
    x = torch.randn(N, C, HW, HW, device="cuda"); w = torch.randn(K, C, 3, 3, device="cuda") * 0.05
    y = torch.empty(N, K, HW, HW, device="cuda"); dy = torch.randn_like(y); dx = torch.empty_like(x); dw = torch.empty_like(w)
    b = torch.zeros(K, device="cuda")
    d = lib.ConvDesc(N, C, HW, HW, K, 3, 3, HW, HW, 1, 1, 1, lib.NCHW, math)
    fns = [lambda: op.Conv2DBiasActivation(x, w, 1, 1, 1, b, lib.ACT_RELU, 0.0, y),
           lambda: op.Conv2DInputGradient(dy, w, 1, 1, 1, lib.NCHW, dx),
           lambda: op.Conv2DKernelsGradient(x, dy, 1, 1, 1, lib.NCHW, dw)]
    out = []
    for i, fn in enumerate(fns):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        tot[i] += ms
        out.append("%s %7.3f ms %6.1f TF/s" % (op.kernel_name(i, d)[:13], ms, d.flops() / ms / 1e9))
    totf += d.flops()
    print("C%-3d K%-3d %3d^2 %5.1f GF | " % (C, K, HW, d.flops() / 1e9) + " | ".join(out), flush=True)
print("TOTAL N=%d: fwd %.2f ms (%.1f TF/s) dgrad %.2f ms (%.1f TF/s) wgrad %.2f ms (%.1f TF/s)" % (
    N, tot[0], totf / tot[0] / 1e9, tot[1], totf / tot[1] / 1e9, tot[2], totf / tot[2] / 1e9))
