"""Generates tests/golden/ref_neighbour_cases.npz from the REFERENCE ITSELF (same rule as make_golden.py): the reference's own
TensorOpCpu loops for the neighbours of the convolution -- activation gradients (TensorOpCpu.cpp:813-864), Pool2D / Pool2DGradient
(:1187-1338), UpSample2D / UpSample2DGradient (:1340-1369), ConstantPad2D (:528-546) -- compiled unmodified into
oracle/_ref/libneuro_ref.so, run on seeded synthetic inputs. Outputs only; inputs are regenerated from the seeds at test time.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden_neighbours.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from neuro__b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

# (name, fmt, N, C, H, W, filter, stride, pad)
POOL_CASES = [("vgg_2x2", 0, 2, 8, 16, 16, 2, 2, 0), ("overlap_3x3_s2_p1", 0, 2, 3, 9, 7, 3, 2, 1), ("same_3x3_s1_p1", 0, 1, 2, 7, 7, 3, 1, 1),
              ("nhwc_2x2", 1, 2, 3, 8, 6, 2, 2, 0), ("nhwc_3x3_s2_p1", 1, 1, 4, 9, 9, 3, 2, 1), ("autoenc_2x2", 0, 3, 16, 28, 28, 2, 2, 0)]
UP_CASES = [("up2", (2, 3, 5, 4), 2), ("up3", (2, 3, 5, 4), 3), ("up2_autoenc", (3, 8, 14, 14), 2)]
PAD_CASES = [("zero_1111", (2, 6, 16, 15), 1, 1, 1, 1, 0.0), ("patchgan_0301", (2, 6, 16, 15), 0, 3, 0, 1, 0.0), ("value_2001", (2, 6, 16, 15), 2, 0, 0, 1, 7.0)]
ACT_SHAPE = (2, 5, 12, 10)


def pool_input(case, mode):
    name, fmt, N, C, H, W, f, st, p = case
    x = synth.uniform(21, (N, C, H, W))
    if mode == O.MAX_POOL:
        x = np.round(x * 4) / 4          # ties inside windows: the first-match rule of the max-pool gradient matters
    return np.ascontiguousarray(x.transpose(0, 2, 3, 1)) if fmt == O.NHWC else x


def act_inputs(act):
    """y in the activation's range (the gradient is taken through the OUTPUT), dy uniform."""
    y = O.conv2d_bias_activation(synth.uniform(11, (ACT_SHAPE[0], 2, ACT_SHAPE[2], ACT_SHAPE[3])), synth.uniform(12, (ACT_SHAPE[1], 2, 1, 1)),
                                 synth.uniform(14, (ACT_SHAPE[1],)), 1, 0, act, 0.2)
    return y, synth.uniform(13, ACT_SHAPE)


def main():
    assert O.have_ref(), "build oracle/_ref first: make -C oracle ref"
    out = {}
    for case in POOL_CASES:
        name, fmt, N, C, H, W, f, st, p = case
        for mode, tag in ((O.MAX_POOL, "max"), (O.AVG_POOL, "avg")):
            x = pool_input(case, mode)
            y = O.ref_pool2d(x, f, st, mode, p, p, fmt)
            dy = synth.uniform(22, y.shape)
            out["pool.%s.%s.y" % (name, tag)] = y
            out["pool.%s.%s.dx" % (name, tag)] = O.ref_pool2d_gradient(y, x, dy, f, st, mode, p, p, fmt)
    for name, shape, s in UP_CASES:
        x = synth.uniform(23, shape)
        y = O.ref_upsample2d(x, s)
        out["up.%s.y" % name] = y
        out["up.%s.dx" % name] = O.ref_upsample2d_gradient(synth.uniform(24, y.shape), s)
    for name, shape, l, r, t, b, v in PAD_CASES:
        out["pad.%s.y" % name] = O.ref_constant_pad2d(synth.uniform(25, shape), l, r, t, b, v)
    for act in (O.IDENTITY, O.SIGMOID, O.RELU, O.TANH, O.ELU, O.LEAKY_RELU):
        y, dy = act_inputs(act)
        out["actgrad.%d.dz" % act] = O.ref_activation_gradient(act, 0.2, y, dy)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_neighbour_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes", len(out), "arrays")


if __name__ == "__main__":
    main()
