"""One cold pass of the DCGAN (batch 128) conv / transposed-conv layers: forward, input gradient, kernel gradient -- for ncu.
usage: profile_gan.py [dcgan|pix2pix]"""
import sys
import torch
sys.path.insert(0, ".")
from neuro__b200 import lib  # noqa: E402
from neuro__b200.shapes import CONFIGS  # noqa: E402
from neuro__b200.tensor_op import TensorOpB200  # noqa: E402

N, LAYERS = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "dcgan"]
op = TensorOpB200(lib.MATH_TF32)
for (name, C, H, K, F, st, p) in LAYERS:
    Ho = (H + 2 * p - F) // st + 1
    x = torch.randn(N, C, H, H, device="cuda"); w = torch.randn(K, C, F, F, device="cuda") * 0.05
    y = torch.empty(N, K, Ho, Ho, device="cuda"); dy = torch.randn_like(y); dx = torch.empty_like(x); dw = torch.empty_like(w)
    op.Conv2D(x, w, st, p, p, lib.NCHW, y)
    op.Conv2DInputGradient(dy, w, st, p, p, lib.NCHW, dx)
    op.Conv2DKernelsGradient(x, dy, st, p, p, lib.NCHW, dw)
    torch.cuda.synchronize()
