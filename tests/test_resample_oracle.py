"""CPU pins for the resampler restatements in oracle/conv_oracle.c: the reference's literal vectors
(Neuro.Tests/src/TensorTests.cpp:427-504) and bit-for-bit equality with the reference's own loops (oracle/_ref)."""
import numpy as np
import pytest

from neuro__b200 import synth
from oracle import oracle as O


def _range(shape, start=0.0):
    """Tensor::FillWithRange(start, 1): flat index order."""
    return (np.arange(int(np.prod(shape)), dtype=np.float32) + np.float32(start)).reshape(shape)


def test_reference_literal_vectors():
    t1 = _range((1, 1, 6, 6))
    assert O.pool2d(t1, 2, 2, O.MAX_POOL).ravel().tolist() == [7, 9, 11, 19, 21, 23, 31, 33, 35]                      # :427-438
    t2 = _range((2, 1, 6, 6))
    assert O.pool2d(t2, 2, 2, O.MAX_POOL).ravel().tolist() == [7, 9, 11, 19, 21, 23, 31, 33, 35, 43, 45, 47, 55, 57, 59, 67, 69, 71]   # :440-450
    assert O.pool2d(t2, 2, 2, O.AVG_POOL).ravel().tolist() == [3.5, 5.5, 7.5, 15.5, 17.5, 19.5, 27.5, 29.5, 31.5, 39.5, 41.5, 43.5,
                                                               51.5, 53.5, 55.5, 63.5, 65.5, 67.5]                     # :452-462
    out = O.pool2d(t2, 2, 2, O.MAX_POOL)
    g = _range(out.shape, 1)
    want = np.zeros((2, 1, 6, 6), np.float32)
    want[:, :, 1::2, 1::2] = g                                                                                         # :464-477
    assert np.array_equal(O.pool2d_gradient(out, t2, g, 2, 2, O.MAX_POOL), want)
    outa = O.pool2d(t2, 2, 2, O.AVG_POOL)
    wanta = np.repeat(np.repeat(g / 4, 2, axis=2), 2, axis=3)                                                         # :479-492
    assert np.array_equal(O.pool2d_gradient(outa, t2, g, 2, 2, O.AVG_POOL), wanta)
    t3 = _range((2, 1, 2, 2))
    assert O.upsample2d(t3, 2).ravel().tolist() == [0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3, 4, 4, 5, 5, 4, 4, 5, 5,
                                                    6, 6, 7, 7, 6, 6, 7, 7]                                            # :494-504


POOL_CASES = [  # (fmt, N, C, H, W, filter, stride, pad)
    (0, 2, 3, 8, 8, 2, 2, 0), (0, 2, 3, 9, 7, 3, 2, 1), (0, 1, 2, 7, 7, 3, 1, 1), (1, 2, 3, 8, 6, 2, 2, 0), (1, 1, 4, 9, 9, 3, 2, 1), (0, 1, 1, 5, 5, 5, 1, 2)]


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("mode", [O.MAX_POOL, O.AVG_POOL])
@pytest.mark.parametrize("cfg", POOL_CASES)
def test_pool_restatement_equals_compiled_reference(cfg, mode):
    fmt, N, C, H, W, f, st, p = cfg
    x = synth.uniform(21, (N, C, H, W))
    x = np.round(x * 4) / 4 if mode == O.MAX_POOL else x          # ties inside windows: the first-match rule matters
    if fmt == O.NHWC:
        x = np.ascontiguousarray(x.transpose(0, 2, 3, 1))
    y = O.pool2d(x, f, st, mode, p, p, fmt)
    assert np.array_equal(y, O.ref_pool2d(x, f, st, mode, p, p, fmt))
    dy = synth.uniform(22, y.shape)
    assert np.array_equal(O.pool2d_gradient(y, x, dy, f, st, mode, p, p, fmt), O.ref_pool2d_gradient(y, x, dy, f, st, mode, p, p, fmt))


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_upsample_and_pad_restatements_equal_compiled_reference():
    x = synth.uniform(23, (2, 3, 5, 4))
    for s in (1, 2, 3):
        y = O.upsample2d(x, s)
        assert np.array_equal(y, O.ref_upsample2d(x, s))
        dy = synth.uniform(24, y.shape)
        assert np.array_equal(O.upsample2d_gradient(dy, s), O.ref_upsample2d_gradient(dy, s))
    for (l, r, t, b, v) in [(1, 1, 1, 1, 0.0), (0, 3, 2, 0, -1.5), (2, 0, 0, 1, 7.0)]:
        assert np.array_equal(O.constant_pad2d(x, l, r, t, b, v), O.ref_constant_pad2d(x, l, r, t, b, v))


# ---- committed outputs of the reference itself (tests/golden/make_golden_neighbours.py): pins the restatement on boxes without
# /root/reference, bit for bit
import os  # noqa: E402
import sys  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_neighbours as G  # noqa: E402

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_neighbour_cases.npz"))


def test_restatement_equals_committed_reference_outputs():
    for case in G.POOL_CASES:
        name, fmt, N, C, H, W, f, st, p = case
        for mode, tag in ((O.MAX_POOL, "max"), (O.AVG_POOL, "avg")):
            x = G.pool_input(case, mode)
            y = O.pool2d(x, f, st, mode, p, p, fmt)
            assert np.array_equal(y, GOLDEN["pool.%s.%s.y" % (name, tag)])
            dy = synth.uniform(22, y.shape)
            assert np.array_equal(O.pool2d_gradient(y, x, dy, f, st, mode, p, p, fmt), GOLDEN["pool.%s.%s.dx" % (name, tag)])
    for name, shape, s in G.UP_CASES:
        y = O.upsample2d(synth.uniform(23, shape), s)
        assert np.array_equal(y, GOLDEN["up.%s.y" % name])
        assert np.array_equal(O.upsample2d_gradient(synth.uniform(24, y.shape), s), GOLDEN["up.%s.dx" % name])
    for name, shape, l, r, t, b, v in G.PAD_CASES:
        assert np.array_equal(O.constant_pad2d(synth.uniform(25, shape), l, r, t, b, v), GOLDEN["pad.%s.y" % name])
    for act in (O.IDENTITY, O.SIGMOID, O.RELU, O.TANH, O.ELU, O.LEAKY_RELU):
        y, dy = G.act_inputs(act)
        assert np.array_equal(O.activation_gradient(act, 0.2, y, dy), GOLDEN["actgrad.%d.dz" % act])
