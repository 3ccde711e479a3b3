"""GPU parity tests: the CUDA path (through the C ABI) against the oracle, the committed reference outputs and the
reference's literal known-answer vectors. Mirrors Neuro.Tests/src/TensorOpGpuTests.cpp:1196-1322 (CPU vs GPU on the
same inputs) and TensorTests.cpp:352-425."""
import os
import sys

import numpy as np
import pytest
import torch

from neuro__b200 import lib, synth
from neuro__b200.tensor_op import TensorOpB200, get_conv_transpose_output_shape
from oracle import oracle as O
from tests.gpu_util import TOL, dev, make_inputs, max_norm_err, run_all_three
from tests.reference_vectors import KNOWN_ANSWERS, known_answer_inputs

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_conv_cases.npz"))
MATHS = [lib.MATH_FP32, lib.MATH_TF32, lib.MATH_3XTF32]
MATH_IDS = ["fp32", "tf32", "3xtf32"]


@pytest.mark.parametrize("name,N,K,pad,expected", KNOWN_ANSWERS, ids=[k[0] for k in KNOWN_ANSWERS])
def test_known_answer_vectors(name, N, K, pad, expected):
    """Integer-valued inputs: every partial sum is exactly representable, so fp32 kernels must match exactly."""
    x, w = known_answer_inputs(N, K)
    op = TensorOpB200(lib.MATH_FP32)
    y = torch.empty((N, K, 4 + 2 * pad, 4 + 2 * pad), device="cuda")
    op.Conv2D(dev(x), dev(w), 1, pad, pad, lib.NCHW, y)
    assert y.cpu().numpy().ravel().tolist() == [float(v) for v in expected]
    y2 = torch.empty((N, 4 + 2 * pad, 4 + 2 * pad, K), device="cuda")
    op.Conv2D(dev(x.transpose(0, 2, 3, 1)), dev(w), 1, pad, pad, lib.NHWC, y2)
    assert np.array_equal(y2.cpu().numpy().transpose(0, 3, 1, 2), y.cpu().numpy())


@pytest.mark.parametrize("math", MATHS, ids=MATH_IDS)
@pytest.mark.parametrize("case", make_golden.CASES, ids=[c[0] for c in make_golden.CASES])
def test_against_committed_reference_outputs(case, math):
    name, fmt, N, C, H, W, K, R, S, st, px, py = case
    x, w, dy = make_golden.inputs(case)
    y, dx, dw = run_all_three(TensorOpB200(math), fmt, x, w, dy, st, px, py)
    assert max_norm_err(y, GOLDEN[name + ".y"]) <= TOL[math]
    assert max_norm_err(dx, GOLDEN[name + ".dx"]) <= TOL[math]
    assert max_norm_err(dw, GOLDEN[name + ".dw"]) <= TOL[math]


# (fmt, N, C, H, W, K, R, S, stride, padX, padY) -- shapes the tensor-core path is meant to take, at sizes the
# oracle finishes in seconds, plus awkward neighbours that must fall back cleanly
TC_CASES = [
    (0, 2, 64, 32, 32, 64, 3, 3, 1, 1, 1),     # VGG-like block, C=K=64
    (0, 1, 128, 64, 64, 128, 3, 3, 1, 1, 1),   # 128 channels, 4 row-tiles
    (0, 1, 64, 36, 40, 96, 3, 3, 1, 1, 1),     # H not multiple of 4, W not multiple of 32, K not multiple of 64
    (0, 2, 40, 32, 32, 72, 3, 3, 1, 1, 1),     # C not a multiple of 32, K not multiple of 8*... (ragged channel tiles)
    (0, 1, 256, 16, 16, 256, 3, 3, 1, 1, 1),   # small map, many channels
    (0, 1, 32, 48, 64, 32, 3, 3, 1, 0, 0),     # valid padding
    (0, 1, 32, 32, 32, 32, 5, 5, 1, 2, 2),     # 5x5
    (0, 1, 64, 32, 32, 64, 1, 1, 1, 0, 0),     # 1x1
    (0, 4, 64, 32, 32, 128, 3, 3, 2, 1, 1),    # DCGAN D stride 2
    (0, 4, 128, 16, 16, 64, 4, 4, 2, 1, 1),    # DCGAN G deconv geometry (4x4 s2 p1)
    (1, 2, 64, 32, 32, 64, 3, 3, 1, 1, 1),     # NHWC twin: the NCHW tensor-core kernels between two layout passes
    (1, 4, 64, 16, 16, 128, 3, 3, 2, 1, 1),    # NHWC, strided (gathered kernels underneath), channel / pixel counts not multiples of 32 in the layout pass tiles
    (1, 2, 40, 20, 24, 72, 3, 3, 1, 1, 1),     # NHWC, ragged channels and filters
    (0, 1, 64, 30, 30, 64, 3, 3, 1, 1, 1),     # W % 4 != 0 -> TMA stride rule fails, must fall back
    (0, 2, 3, 64, 64, 64, 3, 3, 1, 1, 1),      # first-layer (small-channel) kernels: VGG block1_conv1 geometry
    (0, 4, 1, 28, 28, 16, 3, 3, 1, 1, 1),      # conv autoencoder enc conv1 (config 1)
    (0, 2, 3, 512, 512, 16, 3, 3, 1, 1, 1),    # first layer at full resolution: input gradient with two dx rows per thread (>= 148 blocks)
    (0, 2, 3, 30, 27, 10, 3, 3, 1, 1, 1),      # small-channel, ragged width (scalar paths)
    (0, 2, 4, 20, 24, 12, 3, 3, 1, 2, 2),      # small-channel, full padding
    (0, 2, 2, 20, 24, 12, 3, 3, 1, 0, 0),      # small-channel, valid padding
    # gathered-A tensor-core kernel: strides, tiny maps (batch folds into M), odd widths, tap-less parity classes
    (0, 8, 128, 8, 8, 128, 3, 3, 2, 1, 1),     # DCGAN D conv3: 8x8 -> 4x4
    (0, 8, 128, 4, 4, 256, 3, 3, 1, 1, 1),     # DCGAN D conv4: 4x4 maps
    (0, 8, 128, 8, 8, 256, 4, 4, 2, 1, 1),     # DCGAN G deconv1 geometry (dgrad = 4x4 -> 8x8 transposed conv)
    (0, 2, 16, 11, 11, 16, 4, 4, 2, 0, 0),     # ragged stride: last input row/col unreachable
    (0, 1, 16, 35, 35, 32, 4, 4, 2, 0, 0),     # PatchGAN-like odd width
    (0, 2, 16, 9, 9, 24, 1, 1, 2, 0, 0),       # 1x1 stride 2: three of the four dx parity classes have no tap
    (0, 2, 24, 13, 10, 40, 3, 3, 3, 1, 1),     # stride 3
    (0, 4, 16, 4, 8, 16, 3, 3, 2, 1, 1),       # 2x4 output maps: 8-pixel reduction sub-tiles (SWIZZLE_32B rows)
    (0, 6, 200, 16, 16, 72, 4, 4, 2, 1, 1),    # two channel tiles, ragged filter tile, 16 taps in tap groups
    (0, 3, 16, 12, 12, 16, 3, 3, 2, 1, 1),     # 6x6 = 36-pixel output maps: last reduction chunk is partial
    # row-tap kernel (<= 64 filters, 3 columns, pad 1, width >= 64 and % 8 == 0): strips of 32-column tiles with a held sector
    (0, 2, 64, 48, 96, 64, 3, 3, 1, 1, 1),     # three full tiles per strip, fwd and dgrad both take the row-tap kernel
    (0, 3, 40, 36, 72, 24, 3, 3, 1, 1, 1),     # ragged channels, 24 filters (generic epilogue), last tile 8 columns wide
    (0, 3, 32, 38, 64, 64, 5, 3, 1, 1, 2),     # 5 filter rows x 3 columns, H not a multiple of 4
    (0, 1, 128, 96, 128, 48, 3, 3, 1, 1, 1),   # four channel blocks, one image
    # small-channel kernel gradient on the tensor cores (TF32 mode, >= 64K output pixels, W % 4 == 0, pad 1)
    (0, 4, 3, 128, 128, 64, 3, 3, 1, 1, 1),    # RGB first layer: 27 im2col rows in a 32-row tile, filters fill half the M tile
    (0, 2, 4, 128, 256, 160, 3, 3, 1, 1, 1),   # 36 rows in a 48-row tile, two filter tiles (the second ragged)
    (0, 1, 1, 256, 260, 24, 3, 3, 1, 1, 1),    # one channel, last 32-column segment 4 columns wide
    # row-fold kernel gradient (C, K <= 64, 3 columns, padX 1, R <= 3; also taken by several cases above)
    (0, 4, 16, 20, 32, 24, 2, 3, 1, 1, 0),     # two filter rows, no vertical padding
    (0, 4, 32, 16, 64, 16, 1, 3, 1, 1, 0),     # one filter row
    (0, 3, 64, 37, 40, 64, 3, 3, 1, 1, 1),     # rows split across images mid-run (3 x 37 rows over 111 CTAs), ragged segment
    (0, 1, 128, 24, 32, 32, 5, 5, 1, 2, 2),    # 5 columns with > 64 channels: 5 A tiles per step, kernel gradient must take the gathered kernel
    # few-filter layers (K <= 4): the small-channel kernels with x and y exchanged (GAN generator / autoencoder output convs)
    (0, 2, 128, 32, 32, 3, 3, 3, 1, 1, 1),     # DCGAN G conv out 128 -> 3
    (0, 2, 16, 28, 28, 1, 3, 3, 1, 1, 1),      # conv autoencoder dec conv3 16 -> 1
    (0, 1, 64, 256, 256, 3, 3, 3, 1, 1, 1),    # pix2pix last conv geometry: >= 64K pixels, kernel gradient on the tensor cores
    (0, 2, 24, 20, 24, 4, 3, 3, 1, 0, 0),      # valid padding (pad' = 2)
    (0, 2, 24, 20, 22, 2, 3, 3, 1, 2, 2),      # full padding (pad' = 0), ragged width
    (0, 3, 40, 18, 22, 3, 3, 3, 1, 1, 1),      # forward with the 40 channels split over the 8 warps of a block (5 each), ragged width
    # gathered kernel gradient on maps whose planes TMA cannot address directly (pitched copy of dy) and on tiny maps
    (0, 2, 32, 34, 34, 48, 4, 4, 1, 0, 0),     # PatchGAN 31x31 maps: Ho*Wo = 961, odd
    (0, 8, 64, 4, 4, 64, 3, 3, 2, 1, 1),       # U-Net bottleneck: 2x2 maps, 4 pixels in 8-pixel slots
    (0, 8, 64, 2, 2, 64, 3, 3, 2, 1, 1),       # 1x1 maps
    (0, 4, 32, 3, 3, 32, 3, 3, 1, 1, 1),       # 9 pixels in 16-pixel slots, planes pitched to 12
    (0, 2, 16, 5, 5, 16, 3, 3, 1, 1, 1),       # 25 pixels in one 32-pixel chunk, planes pitched to 28
    # stride-2 first layers with 1-6 channels: kernel gradient on the strided few-channel CUDA-core kernel
    (0, 2, 3, 32, 32, 64, 3, 3, 2, 1, 1),      # DCGAN D conv1 / pix2pix enc1 geometry (3 channels, 4 filters per warp)
    (0, 2, 6, 35, 35, 24, 4, 4, 2, 0, 0),      # PatchGAN d1 geometry: 6 channels, 4x4, odd extent, ragged stride (stays on the gathered kernel)
    (0, 2, 2, 19, 23, 24, 4, 4, 2, 0, 0),      # 2 channels, 4x4 (32 running sums, 2 filters per warp), odd extent, ragged stride
    (0, 3, 1, 28, 28, 20, 3, 3, 2, 1, 1),      # MNIST DCGAN first conv: one channel, filter count not a multiple of 32
    (0, 2, 4, 18, 22, 70, 3, 3, 2, 1, 1),      # 36 running sums per filter, two filters per warp, ragged filter blocks
    # gathered forward / input gradient with a narrowed filter tile and channel splits (weight-bound U-Net bottleneck layers)
    (0, 2, 256, 4, 4, 128, 3, 3, 1, 1, 1),     # one pixel tile: forward 2 filter tiles x 4 channel splits, input gradient 4 x 2
    (0, 4, 128, 8, 8, 128, 3, 3, 2, 1, 1),     # stride 2: the four parity classes of the input gradient share one split launch
    (0, 3, 200, 6, 6, 40, 4, 4, 2, 1, 1),      # ragged channel blocks (7 -> splits of 3 + 3 + 1), 16 taps, transposed-conv geometry
    # channel-split forward / input gradient (few tiles, many channels: batch-1 style transfer on the deep layers)
    (0, 1, 256, 32, 32, 128, 3, 3, 1, 1, 1),   # 8 tiles x 4 channel splits (forward), 16 tiles x 2 splits (input gradient)
    (0, 1, 160, 32, 64, 96, 3, 3, 1, 1, 1),    # 5 channel blocks: uneven split 3 + 2
    # few-channel kernel gradient as an SS-form tcgen05 GEMM with gathered im2col rows (TF32, >= 16K output pixels, C*R*S <= 96)
    (0, 8, 3, 128, 128, 24, 3, 3, 2, 1, 1),    # pix2pix enc1 / DCGAN D conv1 geometry: 3 channels, 3x3, stride 2 (27 of 32 rows live)
    (0, 4, 6, 131, 131, 40, 4, 4, 2, 0, 0),    # PatchGAN d1 geometry: 6 channels, 4x4, stride 2, odd input width (x not TMA-addressable), 96 rows
    (0, 2, 5, 100, 134, 136, 3, 3, 1, 0, 0),   # stride 1, 5 channels (48 rows, 45 live), two filter tiles (the second ragged), ragged last segment
    (0, 2, 2, 90, 100, 16, 5, 5, 3, 2, 2),     # stride 3, 5x5, padding 2: 50 rows in 64
    (0, 128, 3, 32, 32, 64, 3, 3, 2, 1, 1),    # DCGAN D conv1 at its BASELINE size: 16-column output rows, half of every 32-pixel dy box is TMA zero fill
    # odd filter / channel counts on the halo-tile path: the repacked filters are only a multiple of 128 bytes
    (0, 2, 16, 32, 32, 9, 3, 3, 1, 1, 1),      # 9 filters (forward repack 9*9*32*4 bytes), 9-row input-gradient tiles
    (0, 2, 32, 32, 32, 21, 1, 1, 1, 0, 0),     # 1x1 conv to 21 classes
    (0, 2, 9, 32, 32, 16, 3, 3, 1, 1, 1),      # 9 channels: the input gradient's repacked rows
]


@pytest.mark.parametrize("math", MATHS, ids=MATH_IDS)
@pytest.mark.parametrize("cfg", TC_CASES, ids=["-".join(map(str, c)) for c in TC_CASES])
def test_against_oracle(cfg, math):
    fmt, N, C, H, W, K, R, S, st, px, py = cfg
    x, w, dy = make_inputs(fmt, N, C, H, W, K, R, S, st, px, py, glorot=True)
    y, dx, dw = run_all_three(TensorOpB200(math), fmt, x, w, dy, st, px, py)
    assert max_norm_err(y, O.conv2d(x, w, st, px, py, fmt)) <= TOL[math]
    assert max_norm_err(dx, O.conv2d_input_gradient(dy, w, st, px, py, (H, W), fmt)) <= TOL[math]
    # long reductions: the reference's own fp32 running sum drifts, compare with its fp64 restatement too
    dw64 = O.conv2d_kernels_gradient(x, dy, st, px, py, (R, S), fmt, f64=True)
    assert max_norm_err(dw, dw64) <= TOL[math]
    assert max_norm_err(dw, O.conv2d_kernels_gradient(x, dy, st, px, py, (R, S), fmt)) <= max(TOL[math], 1e-4)


@pytest.mark.parametrize("math", MATHS, ids=MATH_IDS)
@pytest.mark.parametrize("act", [lib.ACT_IDENTITY, lib.ACT_SIGMOID, lib.ACT_RELU, lib.ACT_TANH, lib.ACT_ELU, lib.ACT_LEAKY_RELU])
def test_bias_activation(act, math):
    """Conv2DBiasActivation (TensorOpGpuTests.cpp:1238-1252 uses ReLU; all epilogues are covered here)."""
    for (N, C, H, W, K, F, st, p) in [(3, 3, 26, 26, 2, 3, 1, 0), (2, 64, 32, 32, 64, 3, 1, 1), (2, 32, 48, 96, 40, 3, 1, 1), (1, 256, 32, 32, 72, 3, 1, 1), (2, 24, 20, 20, 3, 3, 1, 1), (2, 64, 20, 20, 3, 3, 1, 1), (2, 256, 4, 4, 72, 3, 1, 1)]:
        x = synth.uniform(synth.SEED_X, (N, C, H, W)); w = synth.glorot_uniform(synth.SEED_W, K, C, F, F)
        b = synth.uniform(synth.SEED_BIAS, (K,))
        ref = O.conv2d_bias_activation(x, w, b, st, p, act, 0.2)
        y = torch.empty(ref.shape, device="cuda")
        TensorOpB200(math).Conv2DBiasActivation(dev(x), dev(w), st, p, p, dev(b), act, 0.2, y)
        assert max_norm_err(y, ref) <= TOL[math]


@pytest.mark.parametrize("fmt", [lib.NCHW, lib.NHWC])
def test_bias_gradient(fmt):
    """TensorOpGpuTests.cpp:1254-1268 (eps 1e-4 there)."""
    dy = synth.uniform(synth.SEED_DY, (3, 5, 24, 24))
    ref = O.conv2d_bias_gradient(dy)
    a = dy if fmt == lib.NCHW else np.ascontiguousarray(dy.transpose(0, 2, 3, 1))
    db = torch.empty(5, device="cuda")
    TensorOpB200().Conv2DBiasGradient(dev(a), db, fmt)
    assert np.abs(db.cpu().numpy() - ref).max() <= 1e-4
    # folded into the kernel-gradient call
    x = synth.uniform(synth.SEED_X, (3, 3, 26, 26))
    xa = x if fmt == lib.NCHW else np.ascontiguousarray(x.transpose(0, 2, 3, 1))
    dw = torch.empty((5, 3, 3, 3), device="cuda"); db2 = torch.empty(5, device="cuda")
    TensorOpB200(lib.MATH_FP32).Conv2DKernelsGradient(dev(xa), dev(a), 1, 0, 0, fmt, dw, db2)
    assert np.abs(db2.cpu().numpy() - ref).max() <= 1e-4
    assert max_norm_err(dw, O.conv2d_kernels_gradient(xa, a, 1, 0, 0, (3, 3), fmt)) <= 1e-5


@pytest.mark.parametrize("math", MATHS, ids=MATH_IDS)
@pytest.mark.parametrize("cfg", [(2, 8, 4, 4, 16, 4, 2, 1), (2, 16, 8, 8, 8, 3, 2, 0), (1, 64, 16, 16, 64, 4, 2, 1), (2, 6, 5, 5, 4, 3, 1, 1)])
def test_transposed_convolution_identities(cfg, math):
    """Conv2DTranspose: forward = input gradient, input gradient = forward, kernel gradient = kernel gradient with
    (gradient, input) swapped (Tensor.cpp:1806-1830; DeconvolutionLayerTests.cpp geometry: stride 1-2, pad 0-2)."""
    N, Cin, H, W, Cout, F, st, p = cfg
    op = TensorOpB200(math)
    x = synth.uniform(synth.SEED_X, (N, Cin, H, W))
    k = synth.glorot_uniform(synth.SEED_W, Cin, Cout, F, F)      # kernels shape (F,F,outDepth,inDepth)
    out_shape = get_conv_transpose_output_shape(x.shape, Cout, F, F, st, p, p)
    y = torch.empty(out_shape, device="cuda")
    op.Conv2DTransposed(dev(x), dev(k), st, p, lib.NCHW, y)
    ref = O.conv2d_input_gradient(x, k, st, p, p, out_shape[2:])
    assert max_norm_err(y, ref) <= TOL[math]
    g = synth.uniform(synth.SEED_DY, out_shape)
    dxin = torch.empty(x.shape, device="cuda")
    op.Conv2DTransposedInputsGradient(dev(g), dev(k), st, p, lib.NCHW, dxin)
    assert max_norm_err(dxin, O.conv2d(g, k, st, p)) <= TOL[math]
    dk = torch.empty(k.shape, device="cuda")
    op.Conv2DTransposedKernelsGradient(dev(x), dev(g), st, p, lib.NCHW, dk)
    assert max_norm_err(dk, O.conv2d_kernels_gradient(g, x, st, p, p, (F, F), f64=True)) <= TOL[math]


def test_ragged_stride_rows_are_zero_and_outputs_overwritten():
    N, C, H, W, K, F, st = 1, 2, 11, 11, 3, 4, 2
    w = synth.uniform(1, (K, C, F, F)); dy = synth.uniform(2, (N, K, 4, 4))
    dx = torch.full((N, C, H, W), 7.0, device="cuda")   # garbage must be overwritten, not accumulated into
    TensorOpB200(lib.MATH_FP32).Conv2DInputGradient(dev(dy), dev(w), st, 0, 0, lib.NCHW, dx)
    got = dx.cpu().numpy()
    assert np.all(got[:, :, 10, :] == 0) and np.all(got[:, :, :, 10] == 0)
    assert max_norm_err(got, O.conv2d_input_gradient(dy, w, st, 0, 0, (H, W))) <= 1e-5


def test_empty_inputs():
    op = TensorOpB200()
    x = torch.empty((0, 3, 8, 8), device="cuda"); w = torch.ones((4, 3, 3, 3), device="cuda")
    y = torch.empty((0, 4, 6, 6), device="cuda")
    op.Conv2D(x, w, 1, 0, 0, lib.NCHW, y)
    dw = torch.full((4, 3, 3, 3), 5.0, device="cuda")
    op.Conv2DKernelsGradient(x, y, 1, 0, 0, lib.NCHW, dw)
    torch.cuda.synchronize()
    assert float(dw.abs().max()) == 0.0   # sum over an empty batch


@pytest.mark.parametrize("math", [lib.MATH_TF32, lib.MATH_FP32], ids=["tf32", "fp32"])
def test_full_size_layer_adjoint_identities(math):
    """BASELINE-size layer (VGG16 block3: N1 C256 128x128 K256 3x3 s1 p1, 19.3 GFLOP) is too slow for the scalar
    oracle in a unit test, so it is checked through the size-independent adjoint identities
    <dy,conv(x,w)> = <dgrad(dy,w),x> = <wgrad(x,dy),w> (fp64 dots), plus oracle parity on one output row band."""
    N, C, H, W, K, F, st, p = 1, 256, 128, 128, 256, 3, 1, 1
    x = synth.uniform(synth.SEED_X, (N, C, H, W)); w = synth.glorot_uniform(synth.SEED_W, K, C, F, F)
    dy = synth.uniform(synth.SEED_DY, (N, K, H, W))
    y, dx, dw = run_all_three(TensorOpB200(math), lib.NCHW, x, w, dy, st, p, p)
    a = np.dot(dy.ravel().astype(np.float64), y.ravel())
    b = np.dot(dx.ravel().astype(np.float64), x.ravel())
    c = np.dot(dw.ravel().astype(np.float64), w.ravel())
    scale = np.linalg.norm(dy.ravel().astype(np.float64)) * np.linalg.norm(y.ravel().astype(np.float64))
    tol = 2e-3 if math == lib.MATH_TF32 else 1e-5
    assert abs(a - b) <= tol * scale and abs(a - c) <= tol * scale
    # a band of 6 input rows -> 4 output rows, exact same arithmetic as the full problem for those rows
    band = O.conv2d(np.ascontiguousarray(x[:, :, 40:46, :]), w, 1, 1, 0)     # padY=0 over rows 40..45 -> rows 41..44
    assert max_norm_err(y[:, :, 41:45, :], band) <= TOL[math]


def test_optimizer_steps_match_reference_formulas():
    """nb200_adam_step / nb200_sgd_step vs TensorOpCpu::AdamStep / SgdStep (TensorOpCpu.cpp:987-1009). The oracle restatement is
    pinned bit for bit to the compiled reference and to committed reference outputs (tests/test_optim_bn_oracle.py); the CUDA
    kernels round every step separately in the same order, so three consecutive updates are BIT-IDENTICAL."""
    n = 100003
    p = synth.uniform(1, (n,)); m = synth.uniform(3, (n,), 0, 0.1); v = synth.uniform(4, (n,), 0, 0.1)
    pr, mr, vr = p.copy(), m.copy(), v.copy()
    pd, md, vd = dev(p), dev(m), dev(v)
    op = TensorOpB200()
    for step in range(1, 4):
        g = synth.uniform(20 + step, (n,))
        lr_t = float(np.float32(1e-3 * np.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step)))   # Adam.cpp:90
        O.adam_step(pr, g, mr, vr, lr_t, 0.9, 0.999, 1e-8)
        op.AdamStep(pd, dev(g), md, vd, lr_t, 0.9, 0.999, 1e-8)
        assert np.array_equal(pd.cpu().numpy(), pr) and np.array_equal(md.cpu().numpy(), mr) and np.array_equal(vd.cpu().numpy(), vr)
    g = synth.uniform(2, (n,))
    qr = p.copy(); O.sgd_step(qr, g, 0.05)
    qd = dev(p); op.SgdStep(qd, dev(g), 0.05)
    assert np.array_equal(qd.cpu().numpy(), qr)
    # grad_scale folds the 1/replicas of an all-reduced sum (a power of two scales exactly)
    qd2 = dev(p); op.SgdStep(qd2, dev(g * 4), 0.05, gradScale=0.25)
    assert np.array_equal(qd2.cpu().numpy(), qr)
    pd2, md2, vd2 = dev(p), dev(m), dev(v)
    pr2, mr2, vr2 = p.copy(), m.copy(), v.copy()
    O.adam_step(pr2, g, mr2, vr2, 0.01, 0.9, 0.999, 1e-8)
    op.AdamStep(pd2, dev(g * 8), md2, vd2, 0.01, 0.9, 0.999, 1e-8, gradScale=0.125)
    assert np.array_equal(pd2.cpu().numpy(), pr2)


def test_optimizer_steps_match_committed_reference_outputs():
    """The same kernels against OUTPUTS OF THE REFERENCE ITSELF (tests/golden/ref_optim_bn_cases.npz, written by the reference's
    AdamStep / SgdStep compiled into oracle/_ref; generator tests/golden/make_golden_optim_bn.py) -- bit for bit."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_optim_bn as G
    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_optim_bn_cases.npz"))
    p, m, v, grads = G.adam_inputs()
    pd, md, vd = dev(p), dev(m), dev(v)
    op = TensorOpB200()
    for i, g in enumerate(grads):
        op.AdamStep(pd, dev(g), md, vd, G.ADAM_HYPER["lr"], G.ADAM_HYPER["beta1"], G.ADAM_HYPER["beta2"], G.ADAM_HYPER["eps"])
        assert np.array_equal(pd.cpu().numpy(), golden["adam.%d.p" % i]) and np.array_equal(md.cpu().numpy(), golden["adam.%d.m" % i])
        assert np.array_equal(vd.cpu().numpy(), golden["adam.%d.v" % i])
    qd = dev(synth.uniform(41, (G.ADAM_COUNT,)))
    op.SgdStep(qd, dev(grads[0]), G.SGD_LR)
    assert np.array_equal(qd.cpu().numpy(), golden["sgd.p"])


def test_host_buffer_entry_points():
    """e2e C-ABI calls with HOST buffers (copies inside)."""
    import ctypes
    L = lib.load()
    N, C, H, W, K, F = 2, 8, 12, 12, 6, 3
    x, w, dy = make_inputs(lib.NCHW, N, C, H, W, K, F, F, 1, 1, 1)
    b = synth.uniform(synth.SEED_BIAS, (K,))
    d = lib.ConvDesc(N, C, H, W, K, F, F, H, W, 1, 1, 1, lib.NCHW, lib.MATH_FP32)
    y = np.empty((N, K, H, W), np.float32); dx = np.empty_like(x); dw = np.empty_like(w); db = np.empty((K,), np.float32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.check(L.nb200_conv2d_forward_host(ctypes.byref(d), vp(x), vp(w), vp(b), lib.ACT_RELU, 0.0, vp(y), None))
    lib.check(L.nb200_conv2d_input_gradient_host(ctypes.byref(d), vp(dy), vp(w), vp(dx), None))
    lib.check(L.nb200_conv2d_kernels_gradient_host(ctypes.byref(d), vp(x), vp(dy), vp(dw), vp(db), None))
    assert max_norm_err(y, O.conv2d_bias_activation(x, w, b, 1, 1, O.RELU)) <= 1e-5
    assert max_norm_err(dx, O.conv2d_input_gradient(dy, w, 1, 1, 1, (H, W))) <= 1e-5
    assert max_norm_err(dw, O.conv2d_kernels_gradient(x, dy, 1, 1, 1, (F, F))) <= 1e-5
    assert np.abs(db - O.conv2d_bias_gradient(dy)).max() <= 1e-4


ACTS = [lib.ACT_IDENTITY, lib.ACT_SIGMOID, lib.ACT_RELU, lib.ACT_TANH, lib.ACT_ELU, lib.ACT_LEAKY_RELU]


@pytest.mark.parametrize("fmt", [lib.NCHW, lib.NHWC])
@pytest.mark.parametrize("act", ACTS)
def test_bias_activation_gradient(act, fmt):
    """Backward prologue of Conv2dBiasActivationOp (Conv2dBiasActivationOp.cpp:47-60): dz = act'(y)*dy is element-wise
    fp32 with the reference's operation order, so it must be BIT-exact; db = sum(dz) within the reference's own
    CPU-vs-GPU eps (1e-4, TensorOpGpuTests.cpp:1254-1268) relative to the largest sum."""
    op = TensorOpB200()
    # (N, K, Ho, Wo): vector path, ragged scalar path, several 8192-element segments per plane, tiny maps
    for shape in [(3, 5, 24, 24), (2, 64, 32, 32), (2, 7, 13, 9), (2, 4, 96, 100), (16, 24, 4, 4), (1, 1, 1, 1)]:
        y = O.conv2d_bias_activation(synth.uniform(synth.SEED_X, (shape[0], 2, shape[2], shape[3])),
                                     synth.uniform(synth.SEED_W, (shape[1], 2, 1, 1)), synth.uniform(synth.SEED_BIAS, (shape[1],)),
                                     1, 0, act, 0.2)                     # y in the activation's range
        dy = synth.uniform(synth.SEED_DY, shape)
        if fmt == lib.NHWC:
            y = np.ascontiguousarray(y.transpose(0, 2, 3, 1)); dy = np.ascontiguousarray(dy.transpose(0, 2, 3, 1))
        ref_dz = O.activation_gradient(act, 0.2, y, dy)
        ref_db = O.conv2d_bias_gradient(ref_dz, fmt)
        dz = torch.full(dy.shape, float("nan"), device="cuda"); db = torch.full((shape[1],), float("nan"), device="cuda")
        op.Conv2DBiasActivationGradient(dev(y), dev(dy), act, 0.2, dz, db, fmt)
        assert np.array_equal(dz.cpu().numpy(), ref_dz)
        assert np.abs(db.cpu().numpy() - ref_db).max() <= 1e-4 * max(1.0, float(np.abs(ref_db).max()))
        dz2 = torch.full(dy.shape, float("nan"), device="cuda")
        op.ActivationGradient(act, 0.2, dev(y), dev(dy), dz2, fmt)      # without the bias gradient
        assert np.array_equal(dz2.cpu().numpy(), ref_dz)


# (N, C, H, W, K, F, stride, pad): one per kernel family that keeps repacked filters in the workspace, plus one that does not
PREPARED_CASES = [
    (2, 64, 32, 32, 128, 3, 1, 1),    # tcgen05_fprop BN=128
    (1, 64, 48, 96, 64, 3, 1, 1),     # row-tap kernel
    (1, 256, 32, 32, 128, 3, 1, 1),   # channel split: partials live behind the filters in the same workspace
    (4, 64, 32, 32, 128, 3, 2, 1),    # gathered kernel (stride 2; the input gradient runs one launch per parity class)
    (8, 128, 8, 8, 256, 4, 2, 1),     # DCGAN deconv geometry
    (2, 3, 64, 64, 64, 3, 1, 1),      # first-layer kernels read w directly: prepare is a no-op
    (2, 128, 32, 32, 3, 3, 1, 1),     # few-filter layer: the transposed + rotated filters are what is prepared
    (2, 256, 4, 4, 128, 3, 1, 1),     # gathered kernel with channel splits: partials behind the prepared filters
]


@pytest.mark.parametrize("math", [lib.MATH_TF32, lib.MATH_3XTF32], ids=["tf32", "3xtf32"])
@pytest.mark.parametrize("cfg", PREPARED_CASES, ids=["-".join(map(str, c)) for c in PREPARED_CASES])
def test_prepared_filters_match_per_call_repack(cfg, math):
    """nb200_conv2d_prepare_filters + *_prepared run the same kernels on the same repacked filters as the plain calls,
    so results must be bit-identical -- also on reuse, and whatever happens to the op's shared workspace in between."""
    N, C, H, W, K, F, st, p = cfg
    x, w, dy = make_inputs(lib.NCHW, N, C, H, W, K, F, F, st, p, p, glorot=True)
    b = synth.uniform(synth.SEED_BIAS, (K,))
    op = TensorOpB200(math)
    xd, wd, bd, dyd = dev(x), dev(w), dev(b), dev(dy)
    y0 = torch.empty(dy.shape, device="cuda"); dx0 = torch.empty(x.shape, device="cuda")
    op.Conv2DBiasActivation(xd, wd, st, p, p, bd, lib.ACT_RELU, 0.0, y0)
    op.Conv2DInputGradient(dyd, wd, st, p, p, lib.NCHW, dx0)
    assert max_norm_err(y0, O.conv2d_bias_activation(x, w, b, st, p, O.RELU)) <= TOL[math]
    hf = op.PrepareKernels(lib.OP_FORWARD, xd, wd, y0, st, p, p)
    hg = op.PrepareKernels(lib.OP_INPUT_GRADIENT, dx0, wd, dyd, st, p, p)
    for _ in range(2):
        if op._ws is not None:
            op._ws.fill_(0xFF)      # the shared per-call workspace is not what the prepared calls read
        y1 = torch.full(dy.shape, float("nan"), device="cuda"); dx1 = torch.full(x.shape, float("nan"), device="cuda")
        op.Conv2DBiasActivation(xd, wd, st, p, p, bd, lib.ACT_RELU, 0.0, y1, prepared=hf)
        op.Conv2DInputGradient(dyd, wd, st, p, p, lib.NCHW, dx1, prepared=hg)
        assert torch.equal(y1, y0) and torch.equal(dx1, dx0)
    with pytest.raises(AssertionError):
        op.Conv2DInputGradient(dyd, wd, st, p, p, lib.NCHW, dx1, prepared=hf)   # a forward handle is not an input-gradient handle


FULL_SIZE = [  # (name, N, C, H, W, K, F, stride, pad): BASELINE-size layers of configs 2 and 4, too slow for the scalar oracle
    ("dcgan_g_deconv3", 128, 128, 32, 32, 128, 4, 2, 1),      # as a conv 32x32 -> 16x16; its input gradient is the 16 -> 32 transposed conv
    ("pix2pix_patchgan_d4", 8, 256, 34, 34, 512, 4, 1, 0),    # 31x31 output maps: kernel gradient through the pitched copy of dy
    ("pix2pix_unet_dec2", 8, 1024, 4, 4, 512, 3, 1, 1),       # weight-bound: narrowed filter tile + channel splits
    ("pix2pix_last", 8, 128, 256, 256, 3, 3, 1, 1),           # few-filter layer (roles exchanged)
    ("pix2pix_enc1", 8, 3, 256, 256, 64, 3, 2, 1),            # strided few-channel kernel gradient (SS-form tcgen05 GEMM, gathered im2col rows)
    ("pix2pix_patchgan_d1", 8, 6, 259, 259, 64, 4, 2, 0),     # the same with 96 im2col rows on the 259 x 259 padded pair
]


@pytest.mark.parametrize("cfg", FULL_SIZE, ids=[c[0] for c in FULL_SIZE])
def test_full_size_gan_layers_adjoint_identities(cfg):
    """<dy, conv(x,w)> = <dgrad(dy,w), x> = <wgrad(x,dy), w> in fp64 at BASELINE sizes (size-independent property), TF32 bound."""
    _, N, C, H, W, K, F, st, p = cfg
    x, w, dy = make_inputs(lib.NCHW, N, C, H, W, K, F, F, st, p, p, glorot=True)
    y, dx, dw = run_all_three(TensorOpB200(lib.MATH_TF32), lib.NCHW, x, w, dy, st, p, p)
    f64 = lambda a: a.ravel().astype(np.float64)
    a, b, c = np.dot(f64(dy), f64(y)), np.dot(f64(dx), f64(x)), np.dot(f64(dw), f64(w))
    scale = np.linalg.norm(f64(dy)) * np.linalg.norm(f64(y))
    assert abs(a - b) <= 2e-3 * scale and abs(a - c) <= 2e-3 * scale
    assert np.isfinite(dx).all() and np.isfinite(dw).all()
    # one output pixel per image against a direct fp64 evaluation of the definition
    oh, ow = y.shape[2] // 2, y.shape[3] // 3
    xp = np.pad(x, ((0, 0), (0, 0), (p, p), (p, p))).astype(np.float64)
    patch = xp[:, :, oh * st:oh * st + F, ow * st:ow * st + F]
    ref = np.einsum("ncrs,kcrs->nk", patch, w.astype(np.float64))
    assert np.abs(y[:, :, oh, ow] - ref).max() <= 2e-3 * np.abs(ref).max()


# ---- parity AT THE SIZES bench.py RUNS: the longest reductions of the benchmarked configuration ----
BENCH_SIZE = [  # (name, N, C, H, W, K): VGG16 batch 8, 3x3 s1 p1
    ("vgg_block1_conv2", 8, 64, 512, 512, 64),     # kernel-gradient reduction N*Ho*Wo = 2 097 152 (tc_wgrad_rowfold_kernel, tc_rowtap_kernel)
    ("vgg_block2_conv1", 8, 64, 256, 256, 128),    # 524 288 (rowfold kernel gradient, BN = 128 forward, row-tap input gradient)
    ("vgg_block1_conv1", 8, 3, 512, 512, 64),      # 2 097 152 (tc_smallc_wgrad_kernel, small-channel forward / input gradient)
    ("vgg_block3_conv1", 8, 128, 128, 128, 256),   # 131 072 (tc_wgrad_kernel)
]


@pytest.mark.parametrize("math", [lib.MATH_TF32, lib.MATH_3XTF32], ids=["tf32", "3xtf32"])
@pytest.mark.parametrize("cfg", BENCH_SIZE, ids=[c[0] for c in BENCH_SIZE])
def test_bench_size_layers_against_the_oracle(cfg, math):
    """The GPU runs the FULL layer exactly as bench.py does; the oracle (fp64 restatement of TensorOpCpu.cpp:1012-1184) checks
      * the kernel gradient on a filter x channel subset: dw[k, c] depends only on dy[:, k] and x[:, c], so the oracle is given
        those planes alone -- same 2.1 M-term reduction per element, a few hundred MFLOP instead of 155 GFLOP;
      * forward and input gradient on row bands that include the top and bottom padding rows and an image boundary.
    Bounds: TF32 2e-3, 3xTF32 1e-5 max-normalised (BASELINE.json north_star), normalised by the full tensor's maximum."""
    _, N, C, H, W, K = cfg
    F, st, p = 3, 1, 1
    x, w, dy = make_inputs(lib.NCHW, N, C, H, W, K, F, F, st, p, p, glorot=True)
    y, dx, dw = run_all_three(TensorOpB200(math), lib.NCHW, x, w, dy, st, p, p)
    tol = TOL[math]
    ks = [0, K // 2 + 1, K - 1]
    cs = sorted(set([0, C // 2, C - 1]))
    xs = np.ascontiguousarray(x[:, cs]); dys = np.ascontiguousarray(dy[:, ks])
    dw64 = O.conv2d_kernels_gradient(xs, dys, st, p, p, (F, F), f64=True)
    got = dw[np.ix_(ks, cs)]
    assert float(np.abs(got.astype(np.float64) - dw64).max()) <= tol * float(np.abs(dw64).max())
    # forward bands: rows [0, 4) (top padding), [H-4, H) (bottom padding) of the first and last image
    for n in (0, N - 1):
        for (r0, r1, pt, pb) in ((0, 4, 1, 0), (H - 4, H, 0, 1)):
            lo, hi = max(0, r0 - 1), min(H, r1 + 1)
            xb = np.pad(x[n:n + 1, :, lo:hi, :], ((0, 0), (0, 0), (pt, pb), (0, 0)))
            band = O.conv2d(np.ascontiguousarray(xb), w, 1, 1, 0, f64=True)
            assert band.shape[2] == r1 - r0
            assert float(np.abs(y[n:n + 1, :, r0:r1, :].astype(np.float64) - band).max()) <= tol * float(np.abs(band).max())
            # input gradient of the same rows: dx rows [r0, r1) see dy rows [r0-1, r1+1)
            dyb = np.pad(dy[n:n + 1, :, lo:hi, :], ((0, 0), (0, 0), (pt, pb), (0, 0)))
            full = O.conv2d_input_gradient(np.ascontiguousarray(dyb), w, 1, 1, 1, (dyb.shape[2], W), f64=True)
            ref = full[:, :, 1:1 + (r1 - r0), :]
            assert float(np.abs(dx[n:n + 1, :, r0:r1, :].astype(np.float64) - ref).max()) <= tol * float(np.abs(full).max())


def test_full_layer_element_for_element_with_the_compiled_reference():
    """One whole BASELINE layer (VGG16 block5: 512 -> 512 @ 32x32, 4.8 GFLOP per op) against the REFERENCE'S OWN
    TensorOpCpuMt ops (oracle/_ref, compiled unmodified from the reference sources) -- every element of y, dx and dw."""
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    N, C, H, W, K, F, st, p = 1, 512, 32, 32, 512, 3, 1, 1
    x, w, dy = make_inputs(lib.NCHW, N, C, H, W, K, F, F, st, p, p, glorot=True)
    O.ref_set_threads(0)
    yr = O.ref_conv2d(x, w, st, p, p, mt=True)
    dxr = O.ref_conv2d_input_gradient(dy, w, st, p, p, (H, W), mt=True)
    dwr = O.ref_conv2d_kernels_gradient(x, dy, st, p, p, (F, F), mt=True)
    for math in (lib.MATH_TF32, lib.MATH_3XTF32):
        y, dx, dw = run_all_three(TensorOpB200(math), lib.NCHW, x, w, dy, st, p, p)
        tol = TOL[math]
        assert max_norm_err(y, yr) <= tol and max_norm_err(dx, dxr) <= tol
        # the reference's fp32 running sum over 1024 pixels is itself ~1e-6 off; the 3xTF32 bound still holds against it
        assert max_norm_err(dw, dwr) <= tol


PLAN_CASES = [(8, 64, 32, 32, 64, 3, 1, 1), (4, 64, 32, 32, 128, 3, 2, 1), (8, 128, 8, 8, 256, 4, 2, 1), (2, 3, 64, 64, 64, 3, 1, 1), (2, 256, 4, 4, 128, 3, 1, 1)]


@pytest.mark.parametrize("cfg", PLAN_CASES, ids=["-".join(map(str, c)) for c in PLAN_CASES])
def test_plans_replay_the_plain_calls(cfg):
    """nb200_conv2d_plan_*: the op's launches recorded once (tensor maps, grids, workspace layout baked in) and replayed by one
    cudaGraphLaunch -- same kernels in the same order, so bit-identical to the reference-shaped calls; the plan reads the LIVE
    buffers (new inputs -> new outputs), and filters_constant skips the repack launch."""
    N, C, H, W, K, F, st, p = cfg
    x, w, dy = make_inputs(lib.NCHW, N, C, H, W, K, F, F, st, p, p, glorot=True)
    b = synth.uniform(synth.SEED_BIAS, (K,))
    op = TensorOpB200(lib.MATH_TF32)
    xd, wd, bd, dyd = dev(x), dev(w), dev(b), dev(dy)
    y0 = torch.empty(dy.shape, device="cuda"); dx0 = torch.empty(x.shape, device="cuda"); dw0 = torch.empty(w.shape, device="cuda"); db0 = torch.empty(K, device="cuda")
    op.Conv2DBiasActivation(xd, wd, st, p, p, bd, lib.ACT_RELU, 0.0, y0)
    op.Conv2DInputGradient(dyd, wd, st, p, p, lib.NCHW, dx0)
    op.Conv2DKernelsGradient(xd, dyd, st, p, p, lib.NCHW, dw0, db0)
    y1 = torch.full(dy.shape, float("nan"), device="cuda"); dx1 = torch.full(x.shape, float("nan"), device="cuda")
    dw1 = torch.full(w.shape, float("nan"), device="cuda"); db1 = torch.full((K,), float("nan"), device="cuda")
    pf = op.PlanConv2DBiasActivation(xd, wd, st, p, p, bd, lib.ACT_RELU, 0.0, y1, filtersConstant=True)
    pg = op.PlanConv2DInputGradient(dyd, wd, st, p, p, lib.NCHW, dx1)
    pw = op.PlanConv2DKernelsGradient(xd, dyd, st, p, p, lib.NCHW, dw1, db1)
    L = lib.load()
    for _ in range(2):
        for t in (y1, dx1, dw1, db1):
            t.fill_(float("nan"))
        n0 = L.nb200_kernel_launches()
        pf.run(); pg.run(); pw.run()
        assert L.nb200_kernel_launches() - n0 == pf.kernels + pg.kernels + pw.kernels
        assert torch.equal(y1, y0) and torch.equal(dx1, dx0) and torch.equal(dw1, dw0) and torch.equal(db1, db0)
    assert pf.kernels <= pg.kernels      # constant filters: no repack node in the forward plan
    # live buffers: scale the input, the replayed forward follows
    xd.mul_(0.5)
    op.Conv2DBiasActivation(xd, wd, st, p, p, bd, lib.ACT_RELU, 0.0, y0)
    pf.run()
    assert torch.equal(y1, y0)


def test_nhwc_prepared_filters_and_plan():
    """NHWC through the layout passes with constant filters: prepare once, run prepared / as a plan -- same results as the plain NHWC call."""
    N, C, H, W, K, F, st, p = 2, 64, 32, 32, 96, 3, 1, 1
    x, w, dy = make_inputs(lib.NHWC, N, C, H, W, K, F, F, st, p, p, glorot=True)
    op = TensorOpB200(lib.MATH_TF32)
    xd, wd, dyd = dev(x), dev(w), dev(dy)
    assert op.kernel_name(lib.OP_FORWARD, lib.ConvDesc(N, C, H, W, K, F, F, H, W, st, p, p, lib.NHWC, lib.MATH_TF32)).endswith("_nhwc")
    y0 = torch.empty(dy.shape, device="cuda"); dx0 = torch.empty(x.shape, device="cuda")
    op.Conv2D(xd, wd, st, p, p, lib.NHWC, y0); op.Conv2DInputGradient(dyd, wd, st, p, p, lib.NHWC, dx0)
    assert max_norm_err(y0, O.conv2d(x, w, st, p, p, lib.NHWC)) <= TOL[lib.MATH_TF32]
    hf = op.PrepareKernels(lib.OP_FORWARD, xd, wd, y0, st, p, p, lib.NHWC)
    hg = op.PrepareKernels(lib.OP_INPUT_GRADIENT, dx0, wd, dyd, st, p, p, lib.NHWC)
    y1 = torch.full(dy.shape, float("nan"), device="cuda"); dx1 = torch.full(x.shape, float("nan"), device="cuda")
    op.Conv2D(xd, wd, st, p, p, lib.NHWC, y1, prepared=hf); op.Conv2DInputGradient(dyd, wd, st, p, p, lib.NHWC, dx1, prepared=hg)
    assert torch.equal(y1, y0) and torch.equal(dx1, dx0)
    y2 = torch.full(dy.shape, float("nan"), device="cuda")
    plan = op.PlanConv2DBiasActivation(xd, wd, st, p, p, None, lib.ACT_IDENTITY, 0.0, y2, dataFormat=lib.NHWC, filtersConstant=True)
    plan.run()
    assert torch.equal(y2, y0)
