"""One cold pass of the bench workload's ops (every VGG16 layer: forward, input gradient, kernel gradient), for ncu.
usage: profile_step.py [batch]"""
import sys
import torch
sys.path.insert(0, ".")
from neuro__b200 import lib  # noqa: E402
from neuro__b200.tensor_op import TensorOpB200  # noqa: E402

VGG16 = [(3, 64, 512), (64, 64, 512), (64, 128, 256), (128, 128, 256), (128, 256, 128), (256, 256, 128), (256, 256, 128),
         (256, 512, 64), (512, 512, 64), (512, 512, 64), (512, 512, 32), (512, 512, 32), (512, 512, 32)]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
op = TensorOpB200(lib.MATH_TF32)
for (C, K, HW) in VGG16:
    x = torch.randn(B, C, HW, HW, device="cuda"); w = torch.randn(K, C, 3, 3, device="cuda") * 0.05
    y = torch.empty(B, K, HW, HW, device="cuda"); dy = torch.randn_like(y); dx = torch.empty_like(x); dw = torch.empty_like(w)
    b = torch.zeros(K, device="cuda")
    op.Conv2DBiasActivation(x, w, 1, 1, 1, b, lib.ACT_RELU, 0.0, y)
    op.Conv2DInputGradient(dy, w, 1, 1, 1, lib.NCHW, dx)
    op.Conv2DKernelsGradient(x, dy, 1, 1, 1, lib.NCHW, dw)
    torch.cuda.synchronize()
    del x, w, y, dy, dx, dw
