"""CPU tests that PIN the optimiser and batch-normalisation restatements of oracle/conv_oracle.c: bit for bit against the
compiled reference (oracle/_ref, where /root/reference was present at build time) and against committed outputs of the
reference (tests/golden/ref_optim_bn_cases.npz, generator tests/golden/make_golden_optim_bn.py)."""
import os
import sys

import numpy as np
import pytest

from neuro__b200 import synth
from oracle import oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_optim_bn as G  # noqa: E402

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_optim_bn_cases.npz"))


def test_adam_and_sgd_match_committed_reference_outputs():
    """TensorOpCpu::AdamStep / SgdStep (TensorOpCpu.cpp:987-1009), three consecutive updates -- bit for bit."""
    p, m, v, grads = G.adam_inputs()
    for i, g in enumerate(grads):
        O.adam_step(p, g, m, v, **G.ADAM_HYPER)
        assert np.array_equal(p, GOLDEN["adam.%d.p" % i]) and np.array_equal(m, GOLDEN["adam.%d.m" % i]) and np.array_equal(v, GOLDEN["adam.%d.v" % i])
    p2 = synth.uniform(41, (G.ADAM_COUNT,))
    O.sgd_step(p2, grads[0], G.SGD_LR)
    assert np.array_equal(p2, GOLDEN["sgd.p"])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_adam_and_sgd_match_live_reference():
    rng = np.random.RandomState(5)
    for n in (1, 31, 4096, 100003):
        p = rng.uniform(-1, 1, n).astype(np.float32); g = rng.uniform(-1, 1, n).astype(np.float32)
        m = rng.uniform(-.1, .1, n).astype(np.float32); v = rng.uniform(0, .1, n).astype(np.float32)
        a, b = [t.copy() for t in (p, m, v)], [t.copy() for t in (p, m, v)]
        for step in range(1, 4):
            lr_t = 1e-3 * np.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step)   # Adam.cpp:90
            O.adam_step(a[0], g, a[1], a[2], lr_t, 0.9, 0.999, 1e-8)
            O.ref_adam_step(b[0], g, b[1], b[2], lr_t, 0.9, 0.999, 1e-8)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
        a, b = p.copy(), p.copy()
        O.sgd_step(a, g, 0.05); O.ref_sgd_step(b, g, 0.05)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("case", G.BN_CASES, ids=[c[0] for c in G.BN_CASES])
def test_batch_norm_matches_committed_reference_outputs(case):
    """BatchNormalizationTrain / Gradient / inference (TensorOpCpu.cpp:1371-1480), all three modes -- bit for bit."""
    name, mode, shape = case
    x, dy, gamma, beta, rmean, rvar = G.bn_inputs(case)
    y, sm, sv = O.batch_norm_train(mode, x, gamma, beta, G.BN_MOMENTUM, G.BN_EPS, rmean, rvar)
    dx, dg, db = O.batch_norm_gradient(mode, x, gamma, dy, sm, sv)
    for k, a in (("y", y), ("save_mean", sm), ("save_inv_var", sv), ("running_mean", rmean), ("running_var", rvar), ("dx", dx),
                 ("dgamma", dg), ("dbeta", db)):
        assert np.array_equal(a, GOLDEN["bn.%s.%s" % (name, k)]), k
    if mode != O.INSTANCE:
        assert np.array_equal(O.batch_norm(mode, x, gamma, beta, G.BN_EPS, rmean, rvar), GOLDEN["bn.%s.y_inference" % name])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("mode", [O.PER_ACTIVATION, O.SPATIAL, O.INSTANCE])
def test_batch_norm_matches_live_reference(mode):
    rng = np.random.RandomState(11 + mode)
    for it in range(12):
        shape = (rng.randint(1, 9), rng.randint(1, 9), rng.randint(1, 8), rng.randint(1, 8))
        Gn = O.bn_layout(mode, shape)[1]
        x = (rng.uniform(-1, 1, shape) * 3 + 0.5).astype(np.float32); dy = rng.uniform(-1, 1, shape).astype(np.float32)
        gamma = rng.uniform(-1, 1, Gn).astype(np.float32); beta = rng.uniform(-1, 1, Gn).astype(np.float32)
        rm = rng.uniform(-1, 1, Gn).astype(np.float32); rv = rng.uniform(0, 1, Gn).astype(np.float32)
        rm2, rv2 = rm.copy(), rv.copy()
        mom, eps = float(rng.uniform(0.5, 0.99)), float(10 ** rng.uniform(-5, -2))
        a = O.batch_norm_train(mode, x, gamma, beta, mom, eps, rm, rv)
        b = O.ref_batch_norm_train(mode, x, gamma, beta, mom, eps, rm2, rv2)
        assert all(np.array_equal(p, q) for p, q in zip(a, b)) and np.array_equal(rm, rm2) and np.array_equal(rv, rv2)
        ga = O.batch_norm_gradient(mode, x, gamma, dy, a[1], a[2])
        gb = O.ref_batch_norm_gradient(mode, x, gamma, eps, dy, b[1], b[2])
        assert all(np.array_equal(p, q) for p, q in zip(ga, gb))


def test_batch_norm_statistics_are_what_they_claim():
    """Independent float64 check of the restated formulas (not bit-exact: the reference sums sequentially in fp32)."""
    x, dy, gamma, beta, rmean, rvar = G.bn_inputs(G.BN_CASES[3])
    y, sm, sv = O.batch_norm_train(O.SPATIAL, x, gamma, beta, 0.9, 1e-3)
    x64 = x.astype(np.float64)
    mean = x64.mean(axis=(0, 2, 3)); var = x64.var(axis=(0, 2, 3))
    assert np.allclose(sm, mean, atol=1e-5) and np.allclose(sv, 1 / np.sqrt(var + 1e-3), rtol=1e-5)
    yref = (x64 - mean[None, :, None, None]) / np.sqrt(var + 1e-3)[None, :, None, None] * gamma[None, :, None, None] + beta[None, :, None, None]
    assert np.abs(y - yref).max() < 1e-4
