"""Times the HBM-bound neighbour kernels on VGG16 block1-sized activations (batch 8 x 64 x 512 x 512) with CUDA events and prints
achieved GB/s against algorithmic bytes. usage: profile_resample.py"""
import json
import sys
import torch
sys.path.insert(0, ".")
from neuro__b200 import lib  # noqa: E402
from neuro__b200.tensor_op import TensorOpB200  # noqa: E402

op = TensorOpB200()
B = 8
big = torch.rand(B, 64, 512, 512, device="cuda") - 0.5; big2 = torch.rand_like(big); big3 = torch.empty_like(big)
small = torch.empty(B, 64, 256, 256, device="cuda"); small2 = torch.rand_like(small)
pad = torch.empty(B, 64, 514, 516, device="cuda"); db = torch.empty(64, device="cuda")
nb, ns = big.numel() * 4.0, small.numel() * 4.0


def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows = []
def row(name, fn, nbytes):
    ms = t(fn)
    rows.append({"kernel": name, "ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6})
    print("%-44s %7.3f ms %7.0f GB/s" % (name, ms, nbytes / ms / 1e6), flush=True)


row("act_bias_gradient (relu, +db)", lambda: op.Conv2DBiasActivationGradient(big, big2, lib.ACT_RELU, 0.0, big3, db), 3 * nb)
for mode, nm in ((lib.POOL_MAX, "max"), (lib.POOL_AVG, "avg")):
    row("pool2d %s 2x2 s2" % nm, lambda: op.Pool2D(big, 2, 2, mode, 0, 0, lib.NCHW, small), nb + ns)
    op.Pool2D(big, 2, 2, mode, 0, 0, lib.NCHW, small)
    row("pool2d_gradient %s 2x2 s2" % nm, lambda: op.Pool2DGradient(small, big, small2, 2, 2, mode, 0, 0, lib.NCHW, big3),
        (2 * nb + 2 * ns) if mode == lib.POOL_MAX else (nb + ns))
row("pool2d max 3x3 s2 p1 (general kernel)", lambda: op.Pool2D(big, 3, 2, lib.POOL_MAX, 1, 1, lib.NCHW, small), nb + ns)
row("upsample2d x2", lambda: op.UpSample2D(small, 2, big3), nb + ns)
row("upsample2d_gradient x2", lambda: op.UpSample2DGradient(big, 2, small), nb + ns)
row("constant_pad2d (1,3,1,1)", lambda: op.ConstantPad2D(big, 1, 3, 1, 1, 0.0, pad), nb + pad.numel() * 4.0)
json.dump(rows, open("gpurun_out/s6_resample.json", "w"), indent=1)
