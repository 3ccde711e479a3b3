"""CPU tests that PIN the oracle: golden vectors, the compiled reference, committed fixtures, identities."""
import os
import sys

import numpy as np
import pytest

from neuro__b200 import synth
from oracle import oracle as O
from tests.reference_vectors import KNOWN_ANSWERS, known_answer_inputs

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_conv_cases.npz"))


@pytest.mark.parametrize("name,N,K,pad,expected", KNOWN_ANSWERS, ids=[k[0] for k in KNOWN_ANSWERS])
def test_known_answer_vectors(name, N, K, pad, expected):
    """TensorTests.cpp:352-425 -- integer-valued, must match exactly."""
    x, w = known_answer_inputs(N, K)
    y = O.conv2d(x, w, 1, pad)
    assert y.ravel().tolist() == [float(v) for v in expected]
    # NHWC restatement of the same problem gives the same numbers, permuted
    y2 = O.conv2d(np.ascontiguousarray(x.transpose(0, 2, 3, 1)), w, 1, pad, fmt=O.NHWC)
    assert np.array_equal(y2.transpose(0, 3, 1, 2), y)


@pytest.mark.parametrize("case", make_golden.CASES, ids=[c[0] for c in make_golden.CASES])
def test_oracle_matches_committed_reference_outputs(case):
    """Outputs of the reference's own sources (tests/golden/make_golden.py) -- bit for bit."""
    name, fmt, N, C, H, W, K, R, S, st, px, py = case
    x, w, dy = make_golden.inputs(case)
    assert np.array_equal(O.conv2d(x, w, st, px, py, fmt), GOLDEN[name + ".y"])
    assert np.array_equal(O.conv2d_input_gradient(dy, w, st, px, py, (H, W), fmt), GOLDEN[name + ".dx"])
    assert np.array_equal(O.conv2d_kernels_gradient(x, dy, st, px, py, (R, S), fmt), GOLDEN[name + ".dw"])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("fmt", [O.NCHW, O.NHWC])
@pytest.mark.parametrize("mt", [False, True])
def test_oracle_matches_live_reference(fmt, mt):
    """Random sweep against the compiled reference, incl. transposed-conv output sizes (dx larger than needed)."""
    rng = np.random.RandomState(7 + fmt + 2 * mt)
    for it in range(24):
        N, C, K = rng.randint(1, 4), rng.randint(1, 7), rng.randint(1, 7)
        R, S, st = rng.randint(1, 5), rng.randint(1, 5), rng.randint(1, 4)
        px, py = rng.randint(0, R + 1) % (S + 1), rng.randint(0, R + 1)
        H, W = rng.randint(R, R + 9), rng.randint(S, S + 9)
        x = synth.uniform(100 + it, (N, C, H, W))
        w = synth.uniform(200 + it, (K, C, R, S))
        Ho, Wo = O.conv_out_size(H, R, st, py), O.conv_out_size(W, S, st, px)
        dy = synth.uniform(300 + it, (N, K, Ho, Wo))
        if fmt == O.NHWC:
            x = np.ascontiguousarray(x.transpose(0, 2, 3, 1))
            dy = np.ascontiguousarray(dy.transpose(0, 2, 3, 1))
        assert np.array_equal(O.conv2d(x, w, st, px, py, fmt), O.ref_conv2d(x, w, st, px, py, fmt, mt))
        assert np.array_equal(O.conv2d_kernels_gradient(x, dy, st, px, py, (R, S), fmt),
                              O.ref_conv2d_kernels_gradient(x, dy, st, px, py, (R, S), fmt, mt))
        # input gradient at the forward input size, and at the transposed-conv size the reference layer uses
        for hw in ((H, W), (O.conv_transpose_out_size(Ho, R, st, py), O.conv_transpose_out_size(Wo, S, st, px))):
            if hw[0] <= 0 or hw[1] <= 0:
                continue
            assert np.array_equal(O.conv2d_input_gradient(dy, w, st, px, py, hw, fmt),
                                  O.ref_conv2d_input_gradient(dy, w, st, px, py, hw, fmt, mt))


def test_ragged_stride_leaves_untouched_rows_zero():
    """(H+2p-F) % s != 0: the forward ignores the trailing row/col, the input gradient keeps zeros there
    (SURVEY.md section 7 'semantic quirks'; PatchGAN 259/131/67 inputs)."""
    N, C, H, W, K, F, st = 1, 2, 11, 11, 3, 4, 2
    w = synth.uniform(1, (K, C, F, F))
    Ho = O.conv_out_size(H, F, st, 0)
    assert Ho == 4
    dy = synth.uniform(2, (N, K, Ho, Ho))
    dx = O.conv2d_input_gradient(dy, w, st, 0, 0, (H, W))
    assert np.all(dx[:, :, 10, :] == 0) and np.all(dx[:, :, :, 10] == 0)
    assert np.any(dx[:, :, 9, :] != 0)


@pytest.mark.parametrize("cfg", [(2, 3, 9, 5, 3, 1, 1), (2, 4, 8, 6, 4, 2, 1), (3, 3, 7, 2, 3, 2, 0), (2, 6, 11, 4, 4, 2, 0)])
def test_adjoint_identities(cfg):
    """<dy, conv(x,w)> = <dgrad(dy,w), x> = <wgrad(x,dy), w>  (SURVEY.md section 8c), dots in fp64."""
    N, C, H, K, F, st, p = cfg
    x = synth.uniform(11, (N, C, H, H))
    w = synth.uniform(12, (K, C, F, F))
    y = O.conv2d(x, w, st, p)
    dy = synth.uniform(13, y.shape)
    dx = O.conv2d_input_gradient(dy, w, st, p, p, (H, H))
    dw = O.conv2d_kernels_gradient(x, dy, st, p, p, (F, F))
    a = np.dot(dy.ravel().astype(np.float64), y.ravel())
    b = np.dot(dx.ravel().astype(np.float64), x.ravel())
    c = np.dot(dw.ravel().astype(np.float64), w.ravel())
    assert abs(a - b) <= 1e-5 * abs(a) and abs(a - c) <= 1e-5 * abs(a)


def test_f64_variants_agree_with_fp32_order():
    x = synth.uniform(11, (2, 5, 12, 10))
    w = synth.uniform(12, (4, 5, 3, 3))
    y = O.conv2d(x, w, 2, 1)
    dy = synth.uniform(13, y.shape)
    assert np.abs(O.conv2d(x, w, 2, 1, f64=True) - y).max() < 1e-5
    assert np.abs(O.conv2d_input_gradient(dy, w, 2, 1, 1, (12, 10), f64=True)
                  - O.conv2d_input_gradient(dy, w, 2, 1, 1, (12, 10))).max() < 1e-5
    assert np.abs(O.conv2d_kernels_gradient(x, dy, 2, 1, 1, (3, 3), f64=True)
                  - O.conv2d_kernels_gradient(x, dy, 2, 1, 1, (3, 3))).max() < 1e-4


def test_bias_activation_and_bias_gradient():
    """Conv2DBiasActivation = conv -> +bias -> act (TensorOpCpu.cpp:1055-1062); bias grad = sum over N,H,W (:1065)."""
    x = synth.uniform(11, (3, 3, 26, 26))
    w = synth.uniform(12, (2, 3, 3, 3))
    b = synth.uniform(14, (2,))
    y = O.conv2d(x, w, 1, 0)
    for act, fn in [(O.IDENTITY, lambda v: v), (O.RELU, lambda v: np.maximum(v, 0)),
                    (O.LEAKY_RELU, lambda v: np.where(v >= 0, v, np.float32(0.2) * v)),
                    (O.SIGMOID, lambda v: 1 / (1 + np.exp(-v))), (O.TANH, np.tanh),
                    (O.ELU, lambda v: np.where(v >= 0, v, np.float32(0.2) * (np.exp(v) - 1)))]:
        got = O.conv2d_bias_activation(x, w, b, 1, 0, act, 0.2)
        want = fn(y + b.reshape(1, 2, 1, 1))
        assert np.allclose(got, want, atol=2e-6), act
    dy = synth.uniform(13, (3, 5, 24, 24))
    db = O.conv2d_bias_gradient(dy)
    assert np.allclose(db, dy.astype(np.float64).sum(axis=(0, 2, 3)), atol=1e-4)


def test_optimizer_steps():
    p = synth.uniform(1, (1000,)); g = synth.uniform(2, (1000,))
    m = np.zeros_like(p); v = np.zeros_like(p)
    p0 = p.copy()
    O.adam_step(p, g, m, v, 0.01, 0.9, 0.999, 1e-8)
    assert np.allclose(m, 0.1 * g, atol=1e-7) and np.allclose(v, 0.001 * g * g, rtol=1e-4, atol=1e-9)
    assert np.allclose(p, p0 - 0.01 * m / (np.sqrt(v) + 1e-8), atol=1e-5)
    q = p.copy(); O.sgd_step(q, g, 0.1)
    assert np.allclose(q, p - 0.1 * g, atol=1e-7)


@pytest.mark.parametrize("act", [O.IDENTITY, O.SIGMOID, O.RELU, O.TANH, O.ELU, O.LEAKY_RELU])
def test_activation_gradient_restatement(act):
    """dz = act'(y)*dy through the output (TensorOpCpu.cpp:813-864): analytic formulas, and bit-for-bit equality with
    the reference's own ops where oracle/_ref is built."""
    y = O.conv2d_bias_activation(synth.uniform(11, (2, 3, 9, 7)), synth.uniform(12, (4, 3, 1, 1)), synth.uniform(14, (4,)), 1, 0, act, 0.2)
    dy = synth.uniform(13, y.shape)
    got = O.activation_gradient(act, 0.2, y, dy)
    y64 = y.astype(np.float64)
    want = {O.IDENTITY: np.ones_like(y64), O.SIGMOID: y64 * (1 - y64), O.RELU: (y64 > 0) * 1.0, O.TANH: 1 - y64 * y64,
            O.ELU: np.where(y64 > 0, 1.0, y64 + np.float32(0.2)), O.LEAKY_RELU: np.where(y64 > 0, 1.0, np.float32(0.2))}[act] * dy
    assert np.allclose(got, want, atol=1e-6)
    if O.have_ref():
        assert np.array_equal(got, O.ref_activation_gradient(act, 0.2, y, dy))
