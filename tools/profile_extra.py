"""One cold pass of the kernels added next to the VGG16 conv stack, for ncu: the fused activation/bias gradient, the
resamplers, a few-filter layer (roles exchanged), a weight-bound gathered layer with channel splits and a strided input
gradient (all parity classes in one launch). usage: profile_extra.py"""
import sys
import torch
sys.path.insert(0, ".")
from neuro__b200 import lib  # noqa: E402
from neuro__b200.tensor_op import TensorOpB200  # noqa: E402

op = TensorOpB200(lib.MATH_TF32)
B = 8
# backward prologue + pooling on VGG block1 activations (batch 8 x 64 x 512 x 512)
y = torch.rand(B, 64, 512, 512, device="cuda") - 0.5; dy = torch.rand_like(y); dz = torch.empty_like(y); db = torch.empty(64, device="cuda")
op.Conv2DBiasActivationGradient(y, dy, lib.ACT_RELU, 0.0, dz, db)
p = torch.empty(B, 64, 256, 256, device="cuda"); dp = torch.rand_like(p)
for mode in (lib.POOL_MAX, lib.POOL_AVG):
    op.Pool2D(y, 2, 2, mode, 0, 0, lib.NCHW, p)
    op.Pool2DGradient(p, y, dp, 2, 2, mode, 0, 0, lib.NCHW, dz)
op.UpSample2D(p, 2, dz)
op.UpSample2DGradient(dy, 2, p)
pad = torch.empty(B, 64, 514, 514, device="cuda")
op.ConstantPad2D(y, 1, 1, 1, 1, 0.0, pad)
torch.cuda.synchronize()
del y, dy, dz, p, dp, pad


def three(N, C, H, K, F, st, pd):
    Ho = (H + 2 * pd - F) // st + 1
    x = torch.randn(N, C, H, H, device="cuda"); w = torch.randn(K, C, F, F, device="cuda") * 0.05
    yy = torch.empty(N, K, Ho, Ho, device="cuda"); g = torch.randn_like(yy); dx = torch.empty_like(x); dw = torch.empty_like(w)
    op.Conv2D(x, w, st, pd, pd, lib.NCHW, yy)
    op.Conv2DInputGradient(g, w, st, pd, pd, lib.NCHW, dx)
    op.Conv2DKernelsGradient(x, g, st, pd, pd, lib.NCHW, dw)
    torch.cuda.synchronize()


three(8, 128, 256, 3, 3, 1, 1)      # pix2pix last conv 128 -> 3 @256: few-filter kernels
three(8, 1024, 4, 512, 3, 1, 1)     # pix2pix G dec2 1024 -> 512 @4: weight-bound, channel splits
three(128, 128, 32, 128, 4, 2, 1)   # DCGAN G deconv3 geometry: strided, input gradient = 4 parity classes in one launch
three(8, 256, 34, 512, 4, 1, 0)     # PatchGAN 256 -> 512 on 31x31 maps: kernel gradient through the pitched copy of dy
