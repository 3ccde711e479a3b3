"""Diagnostic: time the DCGAN (CifarGAN, batch 128) conv / transposed-conv layer shapes (BASELINE configs[1])."""
import sys
import torch
sys.path.insert(0, ".")
from neuro__b200 import lib  # noqa: E402
from neuro__b200.tensor_op import TensorOpB200  # noqa: E402

N = 128
# (name, C, H, K, F, stride, pad): conv y = conv(x[N,C,H,H], w[K,C,F,F]); "T" rows are Conv2DTranspose layers expressed as the conv
# whose input gradient is their forward: input (N,K,Ho,Ho) -> output (N,C,H,H)
LAYERS = [("D conv1 3->64 s2", 3, 32, 64, 3, 2, 1), ("D conv2 64->128 s2", 64, 16, 128, 3, 2, 1), ("D conv3 128->128 s2", 128, 8, 128, 3, 2, 1),
          ("D conv4 128->256 s1", 128, 4, 256, 3, 1, 1), ("G deconv1 (256x4x4 -> 128x8x8)", 128, 8, 256, 4, 2, 1),
          ("G deconv2 (128x8x8 -> 128x16x16)", 128, 16, 128, 4, 2, 1), ("G deconv3 (128x16x16 -> 128x32x32)", 128, 32, 128, 4, 2, 1),
          ("G conv out 128->3 s1", 128, 32, 3, 3, 1, 1)]
op = TensorOpB200(lib.MATH_TF32)
for (name, C, H, K, F, st, p) in LAYERS:
    Ho = (H + 2 * p - F) // st + 1
    x = torch.randn(N, C, H, H, device="cuda"); w = torch.randn(K, C, F, F, device="cuda") * 0.05
    y = torch.empty(N, K, Ho, Ho, device="cuda"); dy = torch.randn_like(y); dx = torch.empty_like(x); dw = torch.empty_like(w)
    d = lib.ConvDesc(N, C, H, H, K, F, F, Ho, Ho, st, p, p, lib.NCHW, lib.MATH_TF32)
    fns = [lambda: op.Conv2D(x, w, st, p, p, lib.NCHW, y), lambda: op.Conv2DInputGradient(dy, w, st, p, p, lib.NCHW, dx),
           lambda: op.Conv2DKernelsGradient(x, dy, st, p, p, lib.NCHW, dw)]
    out = []
    for i, fn in enumerate(fns):
        fn(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out.append("%s %7.3f ms %6.1f TF/s" % (op.kernel_name(i, d)[:13], ms, d.flops() / ms / 1e9))
    print("%-36s %5.2f GF | " % (name, d.flops() / 1e9) + " | ".join(out), flush=True)
