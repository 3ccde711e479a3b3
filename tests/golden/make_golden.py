"""Generates tests/golden/ref_conv_cases.npz from the REFERENCE ITSELF.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
It drives oracle/_ref/libneuro_ref.so -- the reference's own TensorOpCpu.cpp / TensorOpCpuMt.cpp compiled
unmodified by `make -C oracle ref` -- on seeded synthetic inputs (neuro__b200/synth.py) and stores the
OUTPUTS only; inputs are regenerated from the seeds at test time. The fixture pins the oracle (and through
it the CUDA path) to the reference on boxes where /root/reference is absent.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from neuro__b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

# (name, fmt, N, C, H, W, K, R, S, stride, padX, padY)
CASES = [
    # reference equivalence-test shapes: TensorOpCpuMtTests.cpp:218-298, TensorOpGpuTests.cpp:1196-1322
    ("mt_valid", 0, 3, 3, 26, 26, 2, 3, 3, 1, 0, 0),
    ("gpu_same", 0, 3, 3, 26, 26, 2, 3, 3, 1, 1, 1),
    # finite-difference op test shape: OperationsTests.cpp:40-46
    ("op_fd", 0, 2, 3, 9, 9, 5, 3, 3, 1, 1, 1),
    # layer fit tests: stride {1,2}, pad {0,1,2}: ConvolutionLayerTests.cpp:88-109
    ("s2p1_f4", 0, 2, 4, 8, 8, 6, 4, 4, 2, 1, 1),
    ("s2p0_ragged", 0, 3, 3, 7, 7, 2, 3, 3, 2, 0, 0),
    ("s2p0_f4_ragged", 0, 2, 6, 11, 11, 4, 4, 4, 2, 0, 0),
    ("s1p2_full", 0, 1, 2, 6, 6, 3, 3, 3, 1, 2, 2),
    ("s3_padxy", 0, 1, 2, 10, 7, 3, 3, 3, 3, 2, 1),
    ("rect_filter", 0, 2, 3, 9, 12, 4, 2, 5, 1, 1, 0),
    ("one_by_one", 0, 2, 8, 5, 5, 4, 1, 1, 1, 0, 0),
    # NHWC twins (accepted by the CPU ops; commented out in the reference's GPU tests)
    ("nhwc_valid", 1, 3, 3, 26, 26, 2, 3, 3, 1, 0, 0),
    ("nhwc_s2p1_f4", 1, 2, 4, 8, 8, 6, 4, 4, 2, 1, 1),
    ("nhwc_s2p0_ragged", 1, 3, 3, 7, 7, 2, 3, 3, 2, 0, 0),
    # small versions of BASELINE configs: VGG 3x3 s1 p1, DCGAN D 3x3 s2 p1, DCGAN G deconv 4x4 s2 p1
    ("vgg_like", 0, 1, 16, 32, 32, 16, 3, 3, 1, 1, 1),
    ("dcgan_d", 0, 4, 8, 16, 16, 16, 3, 3, 2, 1, 1),
    ("dcgan_g", 0, 4, 8, 16, 16, 16, 4, 4, 2, 1, 1),
]


def inputs(case):
    name, fmt, N, C, H, W, K, R, S, st, px, py = case
    x = synth.uniform(synth.SEED_X, (N, C, H, W))
    w = synth.uniform(synth.SEED_W, (K, C, R, S))
    Ho, Wo = O.conv_out_size(H, R, st, py), O.conv_out_size(W, S, st, px)
    dy = synth.uniform(synth.SEED_DY, (N, K, Ho, Wo))
    if fmt == O.NHWC:
        x = np.ascontiguousarray(x.transpose(0, 2, 3, 1))
        dy = np.ascontiguousarray(dy.transpose(0, 2, 3, 1))
    return x, w, dy


def main():
    assert O.have_ref(), "build oracle/_ref first: make -C oracle ref"
    out = {}
    for case in CASES:
        name, fmt, N, C, H, W, K, R, S, st, px, py = case
        x, w, dy = inputs(case)
        out[name + ".y"] = O.ref_conv2d(x, w, st, px, py, fmt)
        out[name + ".dx"] = O.ref_conv2d_input_gradient(dy, w, st, px, py, (H, W), fmt)
        out[name + ".dw"] = O.ref_conv2d_kernels_gradient(x, dy, st, px, py, (R, S), fmt)
        # multi-threaded class must agree bit for bit with the single-threaded one (TensorOpCpuMtTests.cpp)
        assert np.array_equal(out[name + ".y"], O.ref_conv2d(x, w, st, px, py, fmt, mt=True))
        assert np.array_equal(out[name + ".dx"], O.ref_conv2d_input_gradient(dy, w, st, px, py, (H, W), fmt, mt=True))
        assert np.array_equal(out[name + ".dw"], O.ref_conv2d_kernels_gradient(x, dy, st, px, py, (R, S), fmt, mt=True))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_conv_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
