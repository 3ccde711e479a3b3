"""GPU parity for the resamplers next to the convolutions (Pool2D / UpSample2D / ConstantPad2D and gradients): the CUDA
kernels add in the reference's order, so every result must be BIT-identical to the oracle (itself pinned to the
reference's literal vectors and compiled loops in tests/test_resample_oracle.py). Mirrors
Neuro.Tests/src/TensorOpGpuTests.cpp (Pool2D_*_CompareWithCpuResult, UpSample2D_*)."""
import numpy as np
import pytest
import torch

from neuro__b200 import lib, synth
from neuro__b200.tensor_op import TensorOpB200
from oracle import oracle as O
from tests.gpu_util import dev

pytestmark = pytest.mark.gpu

POOL_CASES = [  # (fmt, N, C, H, W, filter, stride, pad)
    (0, 2, 3, 8, 8, 2, 2, 0), (0, 2, 3, 9, 7, 3, 2, 1), (0, 1, 2, 7, 7, 3, 1, 1), (1, 2, 3, 8, 6, 2, 2, 0), (1, 1, 4, 9, 9, 3, 2, 1),
    (0, 1, 1, 5, 5, 5, 1, 2), (0, 3, 16, 28, 28, 2, 2, 0),       # conv autoencoder MaxPool2 (config 1)
    (0, 2, 64, 64, 64, 2, 2, 0), (1, 2, 64, 32, 32, 2, 2, 0),     # VGG block pooling geometry (2x2 s2), both formats
    (0, 1, 8, 33, 31, 3, 3, 0), (0, 2, 4, 12, 12, 4, 2, 1)]


@pytest.mark.parametrize("mode", [O.MAX_POOL, O.AVG_POOL], ids=["max", "avg"])
@pytest.mark.parametrize("cfg", POOL_CASES, ids=["-".join(map(str, c)) for c in POOL_CASES])
def test_pool2d_and_gradient(cfg, mode):
    fmt, N, C, H, W, f, st, p = cfg
    x = synth.uniform(21, (N, C, H, W))
    if mode == O.MAX_POOL:
        x = np.round(x * 4) / 4            # ties inside windows: only the first (row-major) maximum may receive the gradient
    if fmt == lib.NHWC:
        x = np.ascontiguousarray(x.transpose(0, 2, 3, 1))
    y_ref = O.pool2d(x, f, st, mode, p, p, fmt)
    dy = synth.uniform(22, y_ref.shape)
    dx_ref = O.pool2d_gradient(y_ref, x, dy, f, st, mode, p, p, fmt)
    op = TensorOpB200()
    xd = dev(x)
    y = torch.full(y_ref.shape, float("nan"), device="cuda"); dx = torch.full(x.shape, float("nan"), device="cuda")
    op.Pool2D(xd, f, st, mode, p, p, fmt, y)
    op.Pool2DGradient(y, xd, dev(dy), f, st, mode, p, p, fmt, dx)
    assert np.array_equal(y.cpu().numpy(), y_ref)
    assert np.array_equal(dx.cpu().numpy(), dx_ref)


def test_pool_reference_literal_vectors():
    """TensorTests.cpp:427-492 on the device."""
    op = TensorOpB200()
    t = torch.arange(72, dtype=torch.float32, device="cuda").view(2, 1, 6, 6)
    y = torch.empty(2, 1, 3, 3, device="cuda")
    op.Pool2D(t, 2, 2, lib.POOL_MAX, 0, 0, lib.NCHW, y)
    assert y.flatten().tolist() == [7, 9, 11, 19, 21, 23, 31, 33, 35, 43, 45, 47, 55, 57, 59, 67, 69, 71]
    g = torch.arange(1, 19, dtype=torch.float32, device="cuda").view(2, 1, 3, 3)
    dx = torch.empty_like(t)
    op.Pool2DGradient(y, t, g, 2, 2, lib.POOL_MAX, 0, 0, lib.NCHW, dx)
    want = torch.zeros_like(t); want[:, :, 1::2, 1::2] = g
    assert torch.equal(dx, want)
    op.Pool2D(t, 2, 2, lib.POOL_AVG, 0, 0, lib.NCHW, y)
    assert y.flatten().tolist() == [3.5, 5.5, 7.5, 15.5, 17.5, 19.5, 27.5, 29.5, 31.5, 39.5, 41.5, 43.5, 51.5, 53.5, 55.5, 63.5, 65.5, 67.5]
    op.Pool2DGradient(y, t, g, 2, 2, lib.POOL_AVG, 0, 0, lib.NCHW, dx)
    assert torch.equal(dx, (g / 4).repeat_interleave(2, 2).repeat_interleave(2, 3))


@pytest.mark.parametrize("shape,s", [((2, 1, 2, 2), 2), ((2, 3, 5, 4), 3), ((4, 8, 14, 14), 2), ((1, 2, 7, 9), 1), ((2, 16, 32, 32), 2)])
def test_upsample2d_and_gradient(shape, s):
    x = synth.uniform(23, shape)
    y_ref = O.upsample2d(x, s)
    dy = synth.uniform(24, y_ref.shape)
    op = TensorOpB200()
    y = torch.full(y_ref.shape, float("nan"), device="cuda"); dx = torch.full(shape, float("nan"), device="cuda")
    op.UpSample2D(dev(x), s, y)
    op.UpSample2DGradient(dev(dy), s, dx)
    assert np.array_equal(y.cpu().numpy(), y_ref)
    assert np.array_equal(dx.cpu().numpy(), O.upsample2d_gradient(dy, s))
    if shape == (2, 1, 2, 2):   # TensorTests.cpp:494-504
        t = torch.arange(8, dtype=torch.float32, device="cuda").view(2, 1, 2, 2)
        op.UpSample2D(t, 2, y)
        assert y.flatten().tolist() == [0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3, 4, 4, 5, 5, 4, 4, 5, 5, 6, 6, 7, 7, 6, 6, 7, 7]


@pytest.mark.parametrize("pads", [(1, 1, 1, 1, 0.0), (0, 3, 2, 0, -1.5), (2, 0, 0, 1, 7.0), (1, 2, 1, 2, 0.0)])
def test_constant_pad2d(pads):
    """PatchGAN's ZeroPadding2D in front of its 4x4 pad-0 convolutions (Pix2Pix.cpp)."""
    l, r, t, b, v = pads
    x = synth.uniform(25, (2, 6, 32, 31))
    ref = O.constant_pad2d(x, l, r, t, b, v)
    y = torch.full(ref.shape, float("nan"), device="cuda")
    TensorOpB200().ConstantPad2D(dev(x), l, r, t, b, v, y)
    assert np.array_equal(y.cpu().numpy(), ref)


def test_full_size_round_trips():
    """BASELINE-size activations (VGG16 block1 at batch 8: 8 x 64 x 512 x 512) through size-independent exact properties:
    pooling undoes nearest-neighbour up-sampling (max and average of four equal values), the up-sampling gradient of a constant
    block is 4x, padding then cropping is the identity, and the max-pool gradient routes every dy to exactly one input."""
    op = TensorOpB200()
    x = torch.rand(8, 64, 256, 256, device="cuda") - 0.5
    up = torch.empty(8, 64, 512, 512, device="cuda")
    op.UpSample2D(x, 2, up)
    back = torch.empty_like(x)
    for mode in (lib.POOL_MAX, lib.POOL_AVG):
        op.Pool2D(up, 2, 2, mode, 0, 0, lib.NCHW, back)
        assert torch.equal(back, x)
    g = torch.empty_like(x)
    op.UpSample2DGradient(up, 2, g)
    assert torch.equal(g, 4 * x)                      # ((a + a) + a) + a is exactly 4a
    padded = torch.empty(8, 64, 516, 516, device="cuda")
    op.ConstantPad2D(up, 1, 3, 2, 2, -7.0, padded)
    assert torch.equal(padded[:, :, 2:514, 1:513], up) and float(padded[:, :, :2].max()) == -7.0 and float(padded[:, :, :, 513:].min()) == -7.0
    r = torch.rand(8, 64, 512, 512, device="cuda")
    pooled = torch.empty_like(x); op.Pool2D(r, 2, 2, lib.POOL_MAX, 0, 0, lib.NCHW, pooled)
    dy = torch.rand_like(x) + 1.0
    dx = torch.empty_like(r); op.Pool2DGradient(pooled, r, dy, 2, 2, lib.POOL_MAX, 0, 0, lib.NCHW, dx)
    assert int((dx != 0).sum()) == dy.numel()                                    # one receiver per window
    assert torch.equal(dx.view(8, 64, 256, 2, 256, 2).sum(dim=(3, 5)), dy)      # and it receives dy itself
    assert torch.equal((dx != 0) * r, (dx != 0) * pooled.repeat_interleave(2, 2).repeat_interleave(2, 3))   # at the maximum


def test_against_committed_reference_outputs():
    """CUDA kernels vs the outputs of the reference's own compiled loops (tests/golden/ref_neighbour_cases.npz), bit for bit."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_neighbours as G
    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_neighbour_cases.npz"))
    op = TensorOpB200()
    for case in G.POOL_CASES:
        name, fmt, N, C, H, W, f, st, p = case
        for mode, tag in ((O.MAX_POOL, "max"), (O.AVG_POOL, "avg")):
            x = G.pool_input(case, mode)
            ref_y = golden["pool.%s.%s.y" % (name, tag)]
            y = torch.full(ref_y.shape, float("nan"), device="cuda"); dx = torch.full(x.shape, float("nan"), device="cuda")
            xd = dev(x)
            op.Pool2D(xd, f, st, mode, p, p, fmt, y)
            op.Pool2DGradient(y, xd, dev(synth.uniform(22, ref_y.shape)), f, st, mode, p, p, fmt, dx)
            assert np.array_equal(y.cpu().numpy(), ref_y), (name, tag)
            assert np.array_equal(dx.cpu().numpy(), golden["pool.%s.%s.dx" % (name, tag)]), (name, tag)
    for name, shape, s in G.UP_CASES:
        ref_y = golden["up.%s.y" % name]
        y = torch.full(ref_y.shape, float("nan"), device="cuda"); dx = torch.full(shape, float("nan"), device="cuda")
        op.UpSample2D(dev(synth.uniform(23, shape)), s, y)
        op.UpSample2DGradient(dev(synth.uniform(24, ref_y.shape)), s, dx)
        assert np.array_equal(y.cpu().numpy(), ref_y) and np.array_equal(dx.cpu().numpy(), golden["up.%s.dx" % name]), name
    for name, shape, l, r, t, b, v in G.PAD_CASES:
        ref_y = golden["pad.%s.y" % name]
        y = torch.full(ref_y.shape, float("nan"), device="cuda")
        op.ConstantPad2D(dev(synth.uniform(25, shape)), l, r, t, b, v, y)
        assert np.array_equal(y.cpu().numpy(), ref_y), name
    for act in (O.IDENTITY, O.SIGMOID, O.RELU, O.TANH, O.ELU, O.LEAKY_RELU):
        yv, dy = G.act_inputs(act)
        dz = torch.full(dy.shape, float("nan"), device="cuda")
        op.ActivationGradient(act, 0.2, dev(yv), dev(dy), dz)
        assert np.array_equal(dz.cpu().numpy(), golden["actgrad.%d.dz" % act]), act


@pytest.mark.parametrize("act", [lib.ACT_RELU, lib.ACT_LEAKY_RELU, lib.ACT_TANH, lib.ACT_SIGMOID, lib.ACT_ELU, lib.ACT_IDENTITY])
def test_fused_pool_activation_bias_gradient_is_the_two_reference_passes(act):
    """nb200_pool2d_gradient_activation = Pool2DGradient (TensorOpCpu.cpp:1249-1338) followed by ActivationGradient + Conv2DBiasGradient
    (Conv2dBiasActivationOp.cpp:47-60) in one pass: dz BIT-IDENTICAL to the oracle's two passes (ties inside windows included), db within
    fp32 summation order."""
    rng = np.random.RandomState(17 + act)
    op = TensorOpB200()
    for (N, C, H, W) in [(2, 5, 8, 16), (3, 4, 36, 40), (1, 3, 128, 256)]:
        z = rng.uniform(-1, 1, (N, C, H, W)).astype(np.float32)
        z[:, :, ::3, ::5] = np.float32(0.25); z[:, :, 1::3, 1::5] = np.float32(0.25)     # ties inside windows
        x = np.maximum(z, 0) if act == lib.ACT_RELU else (np.tanh(z) if act == lib.ACT_TANH else z)   # "activation outputs"
        x = np.ascontiguousarray(x, np.float32)
        y = O.pool2d(x, 2, 2, lib.POOL_MAX)
        dy = rng.uniform(-1, 1, y.shape).astype(np.float32)
        dx_ref = O.pool2d_gradient(y, x, dy, 2, 2, lib.POOL_MAX)
        dz_ref = O.activation_gradient(act, 0.2, x, dx_ref)
        db_ref = dz_ref.astype(np.float64).sum(axis=(0, 2, 3))
        xd, yd, dyd = dev(x), dev(y), dev(dy)
        assert op.Pool2DGradientActivationSupported(xd, 2, 2, lib.POOL_MAX, 0, 0, lib.NCHW, yd)
        dz = torch.full(x.shape, float("nan"), device="cuda"); db = torch.full((C,), float("nan"), device="cuda")
        op.Pool2DGradientActivation(yd, xd, dyd, 2, 2, lib.POOL_MAX, 0, 0, lib.NCHW, act, 0.2, dz, db)
        assert np.array_equal(dz.cpu().numpy(), dz_ref)
        assert np.abs(db.cpu().numpy() - db_ref).max() <= 1e-5 * max(1.0, np.abs(db_ref).max())
        # and it is what the two separate CUDA calls produce
        dx2 = torch.empty_like(dz); dz2 = torch.empty_like(dz)
        op.Pool2DGradient(yd, xd, dyd, 2, 2, lib.POOL_MAX, 0, 0, lib.NCHW, dx2)
        op.Conv2DBiasActivationGradient(xd, dx2, act, 0.2, dz2, None)
        assert torch.equal(dz, dz2)
    # shapes outside the fast path are refused, not silently mishandled
    xd = dev(np.zeros((1, 2, 9, 12), np.float32)); yd = dev(np.zeros((1, 2, 4, 6), np.float32))
    assert not op.Pool2DGradientActivationSupported(xd, 2, 2, lib.POOL_MAX, 0, 0, lib.NCHW, yd)
