"""Diagnostic: run fwd/dgrad/wgrad on a list of shapes in TF32 mode, print the kernel family picked and the
max-normalised error against the oracle. Not a test (never asserts): used to bring up new kernels on the GPU box."""
import ctypes
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from neuro__b200 import lib, synth  # noqa: E402
from neuro__b200.tensor_op import TensorOpB200  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.gpu_util import dev, make_inputs, max_norm_err  # noqa: E402

CASES = [
    # N, C, H, W, K, R, S, st, px, py
    (1, 32, 8, 32, 64, 1, 1, 1, 0, 0),
    (1, 32, 8, 32, 64, 3, 3, 1, 1, 1),
    (2, 64, 32, 32, 64, 3, 3, 1, 1, 1),
    (1, 128, 64, 64, 128, 3, 3, 1, 1, 1),
    (1, 64, 36, 40, 96, 3, 3, 1, 1, 1),
    (2, 40, 32, 32, 72, 3, 3, 1, 1, 1),
    (1, 16, 32, 32, 16, 3, 3, 1, 1, 1),
    (1, 256, 32, 32, 256, 3, 3, 1, 1, 1),
    (1, 32, 48, 64, 32, 3, 3, 1, 0, 0),
    (1, 32, 32, 32, 32, 5, 5, 1, 2, 2),
]
if len(sys.argv) > 1:
    CASES = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]

import os
op = TensorOpB200(int(os.environ.get("NB200_MATH", lib.MATH_TF32)))
for cfg in CASES:
    N, C, H, W, K, R, S, st, px, py = cfg
    x, w, dy = make_inputs(lib.NCHW, N, C, H, W, K, R, S, st, px, py, glorot=True)
    Ho, Wo = dy.shape[2], dy.shape[3]
    d = lib.ConvDesc(N, C, H, W, K, R, S, Ho, Wo, st, px, py, lib.NCHW, op.math)
    names = [op.kernel_name(o, d) for o in (0, 1, 2)]
    xd, wd, dyd = dev(x), dev(w), dev(dy)
    y = torch.zeros(dy.shape, device="cuda"); dx = torch.zeros(x.shape, device="cuda"); dw = torch.zeros(w.shape, device="cuda")
    res = []
    try:
        op.Conv2D(xd, wd, st, px, py, lib.NCHW, y); torch.cuda.synchronize()
        res.append("fwd[%s] %.2e" % (names[0], max_norm_err(y, O.conv2d(x, w, st, px, py))))
        op.Conv2DInputGradient(dyd, wd, st, px, py, lib.NCHW, dx); torch.cuda.synchronize()
        res.append("dgrad[%s] %.2e" % (names[1], max_norm_err(dx, O.conv2d_input_gradient(dy, w, st, px, py, (H, W)))))
        op.Conv2DKernelsGradient(xd, dyd, st, px, py, lib.NCHW, dw); torch.cuda.synchronize()
        res.append("wgrad[%s] %.2e" % (names[2], max_norm_err(dw, O.conv2d_kernels_gradient(x, dy, st, px, py, (R, S), f64=True))))
    except Exception as e:  # noqa: BLE001
        res.append("EXC %r" % (e,))
    print(cfg, " | ".join(res), flush=True)
