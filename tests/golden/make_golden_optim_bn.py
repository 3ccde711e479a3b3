"""Generates tests/golden/ref_optim_bn_cases.npz from the REFERENCE ITSELF (same rule as make_golden.py): the reference's own
TensorOpCpu::AdamStep / SgdStep (TensorOpCpu.cpp:987-1009) and BatchNormalizationTrain / BatchNormalizationGradient /
BatchNormalization (:1371-1480), compiled unmodified into oracle/_ref/libneuro_ref.so (the shim forwards the Tensor
operators they are written with to the reference's own element loops), run on seeded synthetic inputs. Outputs only; inputs
are regenerated from the seeds at test time.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden_optim_bn.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from neuro__b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

ADAM_COUNT, ADAM_STEPS = 4099, 3
ADAM_HYPER = dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8)
SGD_LR = 0.02
# (name, mode, (N, C, H, W)); the first three are the shapes of the reference's own tests (TensorOpGpuTests.cpp:1767-1873:
# Shape(3,4,5,6), momentum 0.9, epsilon 0.001)
BN_MOMENTUM, BN_EPS = 0.9, 0.001
BN_CASES = [("per_activation_3x4x5x6", O.PER_ACTIVATION, (6, 5, 4, 3)), ("spatial_3x4x5x6", O.SPATIAL, (6, 5, 4, 3)),
            ("instance_3x4x5x6", O.INSTANCE, (6, 5, 4, 3)), ("spatial_dcgan_64x7x7", O.SPATIAL, (8, 64, 7, 7)),
            ("spatial_pix2pix_128x16x16", O.SPATIAL, (2, 128, 16, 16)), ("per_activation_dense_100", O.PER_ACTIVATION, (16, 100, 1, 1)),
            ("single_value_copy", O.PER_ACTIVATION, (1, 3, 4, 4))]


def adam_inputs():
    p = synth.uniform(41, (ADAM_COUNT,)); m = synth.uniform(42, (ADAM_COUNT,)) * np.float32(0.1)
    v = np.abs(synth.uniform(43, (ADAM_COUNT,))) * np.float32(0.1)
    grads = [synth.uniform(44 + i, (ADAM_COUNT,)) for i in range(ADAM_STEPS)]
    return p, m, v, grads


def bn_inputs(case):
    name, mode, shape = case
    G = O.bn_layout(mode, shape)[1]
    x = synth.uniform(51, shape) * np.float32(2.0) + np.float32(0.25)
    dy = synth.uniform(52, shape)
    gamma = synth.uniform(53, (G,)); beta = synth.uniform(54, (G,))
    rmean = synth.uniform(55, (G,)); rvar = np.abs(synth.uniform(56, (G,)))
    return x, dy, gamma, beta, rmean, rvar


def main():
    assert O.have_ref(), "build oracle/_ref first: make -C oracle ref"
    out = {}
    p, m, v, grads = adam_inputs()
    for i, g in enumerate(grads):
        O.ref_adam_step(p, g, m, v, **ADAM_HYPER)
        out["adam.%d.p" % i] = p.copy(); out["adam.%d.m" % i] = m.copy(); out["adam.%d.v" % i] = v.copy()
    p2 = synth.uniform(41, (ADAM_COUNT,))
    O.ref_sgd_step(p2, grads[0], SGD_LR)
    out["sgd.p"] = p2
    for case in BN_CASES:
        name, mode, shape = case
        x, dy, gamma, beta, rmean, rvar = bn_inputs(case)
        y, sm, sv = O.ref_batch_norm_train(mode, x, gamma, beta, BN_MOMENTUM, BN_EPS, rmean, rvar)
        dx, dg, db = O.ref_batch_norm_gradient(mode, x, gamma, BN_EPS, dy, sm, sv)
        for k, a in (("y", y), ("save_mean", sm), ("save_inv_var", sv), ("running_mean", rmean), ("running_var", rvar), ("dx", dx),
                     ("dgamma", dg), ("dbeta", db)):
            out["bn.%s.%s" % (name, k)] = a
        if mode != O.INSTANCE:
            out["bn.%s.y_inference" % name] = O.ref_batch_norm(mode, x, gamma, beta, BN_EPS, rmean, rvar)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_optim_bn_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes", len(out), "arrays")


if __name__ == "__main__":
    main()
