"""-m gpu parity tests of the batch-normalisation kernels (neuro__b200/csrc/batchnorm.cu) through the C ABI.

Checker: the oracle restatement of TensorOpCpu::BatchNormalization{,Train,Gradient} (TensorOpCpu.cpp:1371-1480), itself pinned
bit for bit to the compiled reference and to committed reference outputs (tests/test_optim_bn_oracle.py), plus those
committed outputs directly and a float64 evaluation of the formulas. Floating point, reordered sums (the reference adds a
group's elements sequentially in fp32, the kernels use per-block two-pass moments merged by Chan's formula), so the
bound is a TOLERANCE, stated here: max-normalised error <= 2e-5 against the reference at the reference's test sizes,
<= 1e-5 against float64 at BASELINE sizes (where the reference's own sequential fp32 sums are the less accurate side)."""
import os
import sys

import numpy as np
import pytest
import torch

from neuro__b200 import lib
from neuro__b200.tensor_op import TensorOpB200
from oracle import oracle as O
from tests.gpu_util import dev, max_norm_err

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_optim_bn as G  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_optim_bn_cases.npz"))
TOL_REF, TOL_F64 = 2e-5, 1e-5


def run_train(op, mode, x, gamma, beta, momentum, eps, rmean, rvar):
    Gn = O.bn_layout(mode, x.shape)[1]
    xd = dev(x); y = torch.full(x.shape, float("nan"), device="cuda")
    sm = torch.full((Gn,), float("nan"), device="cuda"); sv = torch.full((Gn,), float("nan"), device="cuda")
    rm, rv = dev(rmean), dev(rvar)
    op.BatchNormalizationTrain(xd, mode, dev(gamma), dev(beta), momentum, eps, rm, rv, sm, sv, y)
    torch.cuda.synchronize()
    return y.cpu().numpy(), sm.cpu().numpy(), sv.cpu().numpy(), rm.cpu().numpy(), rv.cpu().numpy()


def run_gradient(op, mode, x, gamma, dy, sm, sv):
    Gn = O.bn_layout(mode, x.shape)[1]
    dx = torch.full(x.shape, float("nan"), device="cuda")
    dg = torch.full((Gn,), float("nan"), device="cuda"); db = torch.full((Gn,), float("nan"), device="cuda")
    op.BatchNormalizationGradient(dev(x), mode, dev(gamma), 0.0, dev(dy), dev(sm), dev(sv), dg, db, True, dx)
    torch.cuda.synchronize()
    return dx.cpu().numpy(), dg.cpu().numpy(), db.cpu().numpy()


@pytest.mark.parametrize("case", G.BN_CASES, ids=[c[0] for c in G.BN_CASES])
def test_batch_norm_matches_committed_reference_outputs(case):
    """All three EBatchNormMode values, the reference's own test shape (TensorOpGpuTests.cpp:1767-1873), GAN shapes, and the
    single-value case where the reference copies the input through."""
    name, mode, shape = case
    x, dy, gamma, beta, rmean, rvar = G.bn_inputs(case)
    op = TensorOpB200()
    y, sm, sv, rm, rv = run_train(op, mode, x, gamma, beta, G.BN_MOMENTUM, G.BN_EPS, rmean, rvar)
    single = O.bn_layout(mode, shape)[0] * O.bn_layout(mode, shape)[2] == 1
    assert max_norm_err(y, GOLDEN["bn.%s.y" % name]) <= TOL_REF
    if single:
        assert np.array_equal(y, x) and np.array_equal(rm, rmean) and np.array_equal(rv, rvar)   # nothing else is written
        dx, dg, db = run_gradient(op, mode, x, gamma, dy, np.zeros_like(gamma), np.ones_like(gamma))
        assert np.array_equal(dx, dy) and not dg.any() and not db.any()
        return
    for k, a in (("save_mean", sm), ("save_inv_var", sv), ("running_mean", rm), ("running_var", rv)):
        assert max_norm_err(a, GOLDEN["bn.%s.%s" % (name, k)]) <= TOL_REF, k
    # gradient from the REFERENCE's saved statistics, so only this op's arithmetic is compared
    dx, dg, db = run_gradient(op, mode, x, gamma, dy, GOLDEN["bn.%s.save_mean" % name], GOLDEN["bn.%s.save_inv_var" % name])
    Nn, Gn, S = O.bn_layout(mode, shape)
    term = np.abs(dy.reshape(Nn, Gn, S) * (gamma * GOLDEN["bn.%s.save_inv_var" % name])[None, :, None]).max()
    assert np.abs(dx.astype(np.float64) - GOLDEN["bn.%s.dx" % name]).max() <= TOL_REF * max(term, np.abs(GOLDEN["bn.%s.dx" % name]).max())
    for k, a in (("dgamma", dg), ("dbeta", db)):
        assert max_norm_err(a, GOLDEN["bn.%s.%s" % (name, k)]) <= TOL_REF, k
    if mode != O.INSTANCE:
        yi = torch.empty(shape, device="cuda")
        op.BatchNormalization(dev(x), mode, dev(gamma), dev(beta), G.BN_EPS, dev(GOLDEN["bn.%s.running_mean" % name]),
                              dev(GOLDEN["bn.%s.running_var" % name]), yi)
        # inference is elementwise with every step rounded like the reference's passes: bit-exact
        assert np.array_equal(yi.cpu().numpy(), GOLDEN["bn.%s.y_inference" % name])


@pytest.mark.parametrize("mode", [O.PER_ACTIVATION, O.SPATIAL, O.INSTANCE])
def test_batch_norm_random_shapes_vs_oracle(mode):
    """Ragged extents: odd H*W (scalar path), H*W a multiple of 4 (16-byte path), one chunk and several chunks per group."""
    rng = np.random.RandomState(100 + mode)
    op = TensorOpB200()
    shapes = [(3, 5, 7, 9), (4, 6, 8, 8), (2, 3, 1, 1), (9, 2, 33, 31), (5, 4, 32, 64), (1, 7, 12, 12), (16, 3, 40, 40)]
    for shape in shapes:
        Gn = O.bn_layout(mode, shape)[1]
        x = (rng.uniform(-1, 1, shape) * 3 + 0.5).astype(np.float32); dy = rng.uniform(-1, 1, shape).astype(np.float32)
        gamma = rng.uniform(-1, 1, Gn).astype(np.float32); beta = rng.uniform(-1, 1, Gn).astype(np.float32)
        rm = rng.uniform(-1, 1, Gn).astype(np.float32); rv = rng.uniform(0, 1, Gn).astype(np.float32)
        mom, eps = float(rng.uniform(0.5, 0.99)), float(10 ** rng.uniform(-5, -2))
        rm_ref, rv_ref = rm.copy(), rv.copy()
        ref = O.batch_norm_train(mode, x, gamma, beta, mom, eps, rm_ref, rv_ref)
        got = run_train(op, mode, x, gamma, beta, mom, eps, rm, rv)
        if O.bn_layout(mode, shape)[0] * O.bn_layout(mode, shape)[2] == 1:
            assert np.array_equal(got[0], x)
            continue
        for a, b in zip(got, ref + (rm_ref, rv_ref)):
            assert max_norm_err(a, b) <= TOL_REF, shape
        gref = O.batch_norm_gradient(mode, x, gamma, dy, ref[1], ref[2])
        ggot = run_gradient(op, mode, x, gamma, dy, ref[1], ref[2])
        # dx = dxNorm*inv + (two correction terms that cancel most of it when a group has few elements): the error scale is that of
        # the TERMS, max |dy*gamma*inv|, not of the (possibly ~0) result
        Nn, Gn2, S = O.bn_layout(mode, shape)
        term = np.abs(dy.reshape(Nn, Gn2, S) * (gamma * ref[2])[None, :, None]).max()
        assert np.abs(ggot[0].astype(np.float64) - gref[0]).max() <= TOL_REF * max(term, np.abs(gref[0]).max()), shape
        for a, b in zip(ggot[1:], gref[1:]):
            assert max_norm_err(a, b) <= TOL_REF, shape


def bn_float64(x, gamma, beta, dy, eps):
    x64, dy64 = x.astype(np.float64), dy.astype(np.float64)
    ax = (0, 2, 3)
    mean = x64.mean(axis=ax, keepdims=True); var = x64.var(axis=ax, keepdims=True)
    inv = 1 / np.sqrt(var + eps)
    g = gamma.astype(np.float64)[None, :, None, None]
    xmu = x64 - mean
    y = xmu * inv * g + beta.astype(np.float64)[None, :, None, None]
    m = x64.size / x64.shape[1]
    dxn = dy64 * g
    dvar = (dxn * xmu).sum(axis=ax, keepdims=True) * -0.5 * inv ** 3
    dmu = (dxn * -inv).sum(axis=ax, keepdims=True) + dvar * (xmu * -2).mean(axis=ax, keepdims=True)
    dx = dxn * inv + dvar * xmu * 2 / m + dmu / m
    return y, mean.ravel(), inv.ravel(), dx, (dy64 * xmu * inv).sum(axis=ax), dy64.sum(axis=ax)


def test_batch_norm_at_baseline_sizes_vs_float64():
    """pix2pix encoder (8 x 64 x 128 x 128: 131072 elements per channel, 32 blocks per group) and DCGAN (128 x 128 x 8 x 8)."""
    rng = np.random.RandomState(7)
    op = TensorOpB200()
    for shape in ((8, 64, 128, 128), (128, 128, 8, 8)):
        C = shape[1]
        x = (rng.standard_normal(shape) * 1.7 + 0.8).astype(np.float32); dy = rng.uniform(-1, 1, shape).astype(np.float32)
        gamma = rng.uniform(0.5, 1.5, C).astype(np.float32); beta = rng.uniform(-1, 1, C).astype(np.float32)
        y, sm, sv, _, _ = run_train(op, O.SPATIAL, x, gamma, beta, 0.9, 1e-3, np.zeros(C, np.float32), np.ones(C, np.float32))
        dx, dg, db = run_gradient(op, O.SPATIAL, x, gamma, dy, sm, sv)
        ref = bn_float64(x, gamma, beta, dy, 1e-3)
        for a, b in zip((y, sm, sv, dx, dg, db), ref):
            assert max_norm_err(a, b) <= TOL_F64


def test_batch_norm_over_replicas_equals_the_full_batch():
    """Batch-sharded replicas: per-shard moments gathered in rank order and gradient sums added give the statistics and the
    input gradient of the full batch (what the exchange in neuro__b200/fit.py does with torch.distributed)."""
    rng = np.random.RandomState(3)
    op = TensorOpB200()
    R, shape = 4, (16, 24, 12, 12)
    C = shape[1]
    x = (rng.standard_normal(shape) * 2 + np.arange(shape[0])[:, None, None, None] * 0.3).astype(np.float32)   # shards differ in mean
    dy = rng.uniform(-1, 1, shape).astype(np.float32)
    gamma = rng.uniform(0.5, 1.5, C).astype(np.float32); beta = rng.uniform(-1, 1, C).astype(np.float32)
    ref = bn_float64(x, gamma, beta, dy, 1e-3)
    per = shape[0] // R
    xs = [dev(x[r * per:(r + 1) * per]) for r in range(R)]; dys = [dev(dy[r * per:(r + 1) * per]) for r in range(R)]
    gd, bd = dev(gamma), dev(beta)
    allm = torch.empty(R, C, 2, device="cuda")
    for r in range(R):
        op.BatchNormalizationMoments(xs[r], O.SPATIAL, allm[r])
    ys, stats = [], []
    for r in range(R):
        y = torch.empty_like(xs[r]); sm = torch.empty(C, device="cuda"); sv = torch.empty(C, device="cuda")
        rm = torch.zeros(C, device="cuda"); rv = torch.ones(C, device="cuda")
        op.BatchNormalizationTrainFromMoments(allm, R, xs[r], O.SPATIAL, gd, bd, 0.9, 1e-3, rm, rv, sm, sv, y)
        ys.append(y); stats.append((sm, sv, rm, rv))
    for r in range(1, R):   # every replica derives bit-identical statistics
        assert all(torch.equal(a, b) for a, b in zip(stats[0], stats[r]))
    assert max_norm_err(torch.cat(ys), ref[0]) <= TOL_F64 and max_norm_err(stats[0][0], ref[1]) <= TOL_F64
    assert max_norm_err(stats[0][1], ref[2]) <= TOL_F64
    m = x.size / C
    assert max_norm_err(stats[0][3], 0.1 * 1 + 0.9 * x.astype(np.float64).var(axis=(0, 2, 3)) * m / (m - 1)) <= TOL_F64
    sums = torch.empty(R, C, 3, device="cuda")
    for r in range(R):
        op.BatchNormalizationGradientSums(xs[r], O.SPATIAL, dys[r], stats[0][0], sums[r])
    glob = sums.sum(dim=0).contiguous()   # the all-reduce
    dxs, dgs, dbs = [], [], []
    for r in range(R):
        dx = torch.empty_like(xs[r]); dg = torch.empty(C, device="cuda"); db = torch.empty(C, device="cuda")
        op.BatchNormalizationGradientFromSums(R, glob, sums[r].contiguous(), xs[r], O.SPATIAL, gd, dys[r], stats[0][0], stats[0][1], dg, db, dx)
        dxs.append(dx); dgs.append(dg); dbs.append(db)
    assert max_norm_err(torch.cat(dxs), ref[3]) <= TOL_F64
    assert max_norm_err(sum(dgs), ref[4]) <= TOL_F64 and max_norm_err(sum(dbs), ref[5]) <= TOL_F64
