"""Helpers shared by the -m gpu parity tests (CUDA path vs oracle on identical seeded inputs)."""
import numpy as np
import torch

from neuro__b200 import lib, synth
from neuro__b200.tensor_op import TensorOpB200
from oracle import oracle as O

# max-normalised error bounds from BASELINE.json north_star; FP32 kernels must also meet the 3xTF32 bound
TOL = {lib.MATH_TF32: 2e-3, lib.MATH_3XTF32: 1e-5, lib.MATH_FP32: 1e-5}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def max_norm_err(got, ref):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    denom = float(np.abs(ref).max())
    return float(np.abs(got.astype(np.float64) - ref).max()) / (denom if denom > 0 else 1.0)


def make_inputs(fmt, N, C, H, W, K, R, S, st, px, py, glorot=False):
    x = synth.uniform(synth.SEED_X, (N, C, H, W))
    w = synth.glorot_uniform(synth.SEED_W, K, C, R, S) if glorot else synth.uniform(synth.SEED_W, (K, C, R, S))
    Ho, Wo = O.conv_out_size(H, R, st, py), O.conv_out_size(W, S, st, px)
    dy = synth.uniform(synth.SEED_DY, (N, K, Ho, Wo))
    if fmt == O.NHWC:
        x = np.ascontiguousarray(x.transpose(0, 2, 3, 1))
        dy = np.ascontiguousarray(dy.transpose(0, 2, 3, 1))
    return x, w, dy


def run_all_three(op, fmt, x, w, dy, st, px, py):
    """fwd, dgrad, wgrad through TensorOpB200 (C ABI underneath); returns numpy arrays."""
    xd, wd, dyd = dev(x), dev(w), dev(dy)
    y = torch.full(dy.shape, float("nan"), device="cuda")
    dx = torch.full(x.shape, float("nan"), device="cuda")
    dw = torch.full(w.shape, float("nan"), device="cuda")
    op.Conv2D(xd, wd, st, px, py, fmt, y)
    op.Conv2DInputGradient(dyd, wd, st, px, py, fmt, dx)
    op.Conv2DKernelsGradient(xd, dyd, st, px, py, fmt, dw)
    torch.cuda.synchronize()
    return y.cpu().numpy(), dx.cpu().numpy(), dw.cpu().numpy()
