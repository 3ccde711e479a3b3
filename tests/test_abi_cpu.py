"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports what include/neuro_b200.h declares,
validates descriptors, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

import __graft_entry__ as graft
from neuro__b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    graft.build()
    return lib.load()


def test_header_and_exports_agree(L):
    header = open(os.path.join(ROOT, "include", "neuro_b200.h")).read()
    declared = sorted(set(re.findall(r"NB200_API[^;(]*?\b(nb200_\w+)\s*\(", header)))
    assert declared == sorted(lib.EXPORTS)
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.SO_PATH]).decode()
    exported = sorted(s for s in re.findall(r" T (\w+)", out) if s.startswith("nb200_"))
    assert exported == declared
    for name in declared:
        assert getattr(L, name) is not None


def test_shape_helpers_match_reference_formulas(L):
    # Tensor::GetPadding (Tensor.cpp:1966-1985)
    assert [L.nb200_padding(m, 3) for m in (0, 1, 2)] == [0, 1, 2]
    assert [L.nb200_padding(m, 4) for m in (0, 1, 2)] == [0, 2, 3]
    # Tensor::GetConvOutputShape (:2010-2029): floor division
    assert L.nb200_conv_out_size(6, 3, 1, 0) == 4 and L.nb200_conv_out_size(6, 3, 1, 1) == 6
    assert L.nb200_conv_out_size(259, 4, 2, 0) == 128 and L.nb200_conv_out_size(11, 4, 2, 0) == 4
    # Tensor::GetConvTransposeOutputShape (:2032-2051): smaller than the conv input when the stride is ragged
    assert L.nb200_conv_transpose_out_size(4, 4, 2, 1) == 8
    assert L.nb200_conv_transpose_out_size(4, 4, 2, 0) == 10


def test_descriptor_validation(L):
    d = lib.ConvDesc(1, 2, 6, 6, 1, 3, 3, 4, 4, 1, 0, 0, lib.NCHW, lib.MATH_FP32)
    bad = lib.ConvDesc(1, 2, 6, 6, 1, 3, 3, 5, 4, 1, 0, 0, lib.NCHW, lib.MATH_FP32)
    p = ctypes.c_void_p(16)
    assert L.nb200_conv2d_forward(ctypes.byref(bad), p, p, None, 0, 0.0, p, None, 0, None) == -1
    assert b"GetConvOutputShape" in L.nb200_last_error()
    bad2 = lib.ConvDesc(1, 2, 6, 6, 1, 3, 3, 4, 4, 0, 0, 0, lib.NCHW, lib.MATH_FP32)
    assert L.nb200_conv2d_forward(ctypes.byref(bad2), p, p, None, 0, 0.0, p, None, 0, None) == -1
    assert L.nb200_conv2d_forward(ctypes.byref(d), None, p, None, 0, 0.0, p, None, 0, None) == -1
    assert L.nb200_conv2d_forward(ctypes.byref(d), p, p, None, 6, 0.0, p, None, 0, None) == -1  # _Softmax is no epilogue
    # empty batch is a no-op, not an error (and never touches the device)
    empty = lib.ConvDesc(0, 2, 6, 6, 1, 3, 3, 4, 4, 1, 0, 0, lib.NCHW, lib.MATH_FP32)
    assert L.nb200_conv2d_forward(ctypes.byref(empty), None, None, None, 0, 0.0, None, None, 0, None) == 0


def test_neighbour_entry_points_validate_before_touching_a_device(L):
    """Resamplers, the fused activation/bias gradient and filter preparation reject bad arguments with NB200_E_INVALID (host logic),
    accept empty tensors, and -- like every compute call -- refuse to run without a device rather than falling back."""
    good = lib.PoolDesc(2, 3, 8, 8, 4, 4, 2, 2, 0, 0, lib.POOL_MAX, lib.NCHW)
    bad_out = lib.PoolDesc(2, 3, 8, 8, 5, 4, 2, 2, 0, 0, lib.POOL_MAX, lib.NCHW)        # GetPooling2DOutputShape gives 4x4
    bad_mode = lib.PoolDesc(2, 3, 8, 8, 4, 4, 2, 2, 0, 0, 7, lib.NCHW)
    too_big = lib.PoolDesc(2, 3, 2, 2, 1, 1, 5, 1, 0, 0, lib.POOL_AVG, lib.NCHW)         # window larger than the padded input
    for d in (bad_out, bad_mode, too_big):
        assert L.nb200_pool2d(ctypes.byref(d), None, None, None) == -1
        assert L.nb200_pool2d_gradient(ctypes.byref(d), None, None, None, None, None) == -1
    huge = lib.PoolDesc(64, 64, 2048, 2048, 1024, 1024, 2, 2, 0, 0, lib.POOL_MAX, lib.NCHW)  # 2^34 elements: beyond Shape::Length (uint32)
    assert L.nb200_pool2d(ctypes.byref(huge), None, None, None) == -1
    assert L.nb200_pool2d(ctypes.byref(good), None, None, None) == -1                    # null tensors
    empty = lib.PoolDesc(0, 3, 8, 8, 4, 4, 2, 2, 0, 0, lib.POOL_MAX, lib.NCHW)
    assert L.nb200_pool2d(ctypes.byref(empty), None, None, None) == 0
    assert L.nb200_upsample2d(0, 3, 4, 4, 2, None, None, None) == 0 and L.nb200_upsample2d(1, 3, 4, 4, 0, None, None, None) == -1
    assert L.nb200_upsample2d_gradient(1, 1, 40000, 40000, 2, None, None, None) == -1    # exceeds Shape::Length (uint32)
    assert L.nb200_constant_pad2d(1, 1, 4, 4, -1, 0, 0, 0, 0.0, None, None, None) == -1
    d = _desc(2, 8, 16, 16, 8, 3, 3, 1, 1, 1)
    assert L.nb200_conv2d_bias_activation_gradient(ctypes.byref(d), 6, 0.0, None, None, None, None, None, 0, None) == -1   # _Softmax has no gradient here
    assert L.nb200_conv2d_prepare_filters(lib.OP_KERNELS_GRADIENT, ctypes.byref(d), None, None, 0, None) == -1
    assert L.nb200_conv2d_prepare_filters(lib.OP_FORWARD, ctypes.byref(d), None, None, 0, None) == -1                      # null filters
    import torch
    if not torch.cuda.is_available():
        one = ctypes.c_float(0.0)
        p = ctypes.cast(ctypes.pointer(one), ctypes.c_void_p)
        assert L.nb200_pool2d(ctypes.byref(good), p, p, None) == -2                       # NB200_E_NO_DEVICE, never a CPU loop
        assert L.nb200_upsample2d(1, 1, 2, 2, 2, p, p, None) == -2
        assert L.nb200_conv2d_bias_activation_gradient(ctypes.byref(d), 2, 0.0, p, p, p, None, None, 0, None) == -2


def test_no_cpu_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d = lib.ConvDesc(1, 2, 6, 6, 1, 3, 3, 4, 4, 1, 0, 0, lib.NCHW, lib.MATH_FP32)
    p = ctypes.c_void_p(16)
    assert L.nb200_conv2d_forward(ctypes.byref(d), p, p, None, 0, 0.0, p, None, 0, None) == -2
    assert L.nb200_conv2d_input_gradient(ctypes.byref(d), p, p, p, None, 0, None) == -2
    assert L.nb200_conv2d_kernels_gradient(ctypes.byref(d), p, p, p, None, None, 0, None) == -2
    assert L.nb200_adam_step(p, p, p, p, 4, 1.0, 0.1, 0.9, 0.999, 1e-8, None) == -2
    with pytest.raises(lib.NeuroB200Error):
        lib.check(-2)


def test_product_does_not_touch_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use oracle/."""
    pkg = os.path.join(ROOT, "neuro__b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower().replace("# oracle-free", ""), os.path.join(dirpath, f)
    assert "oracle" not in open(os.path.join(ROOT, "include", "neuro_b200.h")).read().lower()


def _desc(N, C, H, W, K, R, S, stride, padX, padY, fmt=lib.NCHW, math=lib.MATH_TF32):
    Ho = (H + 2 * padY - R) // stride + 1
    Wo = (W + 2 * padX - S) // stride + 1
    return lib.ConvDesc(N, C, H, W, K, R, S, Ho, Wo, stride, padX, padY, fmt, math)


def test_kernel_dispatch_is_host_logic(L):
    """Which kernel family serves a shape is decided on the host (no GPU needed): pin it for the shapes BASELINE.json names,
    so that a heuristic change that silently drops a layer off the tensor cores shows up in the CPU suite."""
    L.nb200_conv2d_kernel_name.restype = ctypes.c_char_p

    def names(d):
        return [L.nb200_conv2d_kernel_name(op, ctypes.byref(d)).decode() for op in (lib.OP_FORWARD, lib.OP_INPUT_GRADIENT, lib.OP_KERNELS_GRADIENT)]

    # VGG16 @ 512x512, batch 8 (bench workload)
    assert names(_desc(8, 3, 512, 512, 64, 3, 3, 1, 1, 1)) == ["smallc_fprop", "smallc_dgrad", "tcgen05_smallc_wgrad"]
    assert names(_desc(8, 64, 512, 512, 64, 3, 3, 1, 1, 1)) == ["tcgen05_rowtap_fprop", "tcgen05_rowtap_dgrad", "tcgen05_rowfold_wgrad"]
    assert names(_desc(8, 64, 256, 256, 128, 3, 3, 1, 1, 1)) == ["tcgen05_fprop", "tcgen05_rowtap_dgrad", "tcgen05_rowfold_wgrad"]
    for (C, K, HW) in [(128, 128, 256), (128, 256, 128), (256, 256, 128), (256, 512, 64), (512, 512, 64), (512, 512, 32)]:
        assert names(_desc(8, C, HW, HW, K, 3, 3, 1, 1, 1)) == ["tcgen05_fprop", "tcgen05_dgrad", "tcgen05_wgrad"]
    # DCGAN / pix2pix geometry: strided, tiny maps, transposed convolutions -> gathered tensor-core kernels
    assert names(_desc(128, 64, 32, 32, 128, 3, 3, 2, 1, 1)) == ["tcgen05_gather_fprop", "tcgen05_gather_dgrad", "tcgen05_gather_wgrad"]
    assert names(_desc(128, 128, 8, 8, 256, 4, 4, 2, 1, 1))[0].startswith("tcgen05_gather")
    # pix2pix shapes that used to fall off the tensor cores in the kernel gradient: 31x31 PatchGAN maps (odd plane size), U-Net bottleneck
    assert names(_desc(8, 256, 34, 34, 512, 4, 4, 1, 0, 0))[2] == "tcgen05_gather_wgrad"
    assert names(_desc(8, 512, 4, 4, 512, 3, 3, 2, 1, 1))[2] == "tcgen05_gather_wgrad"
    assert names(_desc(8, 512, 2, 2, 512, 3, 3, 2, 1, 1))[2] == "tcgen05_gather_wgrad"
    # stride-2 first layers with 3 / 6 channels (pix2pix enc1, PatchGAN d1, DCGAN D conv1): HBM-bound kernel gradient as an SS-form
    # tcgen05 GEMM over the dy stream with the im2col rows gathered from global memory (C*R*S <= 96; >= 16K output pixels) ...
    assert names(_desc(8, 3, 256, 256, 64, 3, 3, 2, 1, 1)) == ["tcgen05_gather_fprop", "tcgen05_gather_dgrad", "tcgen05_smallc_gather_wgrad"]
    assert names(_desc(8, 6, 259, 259, 64, 4, 4, 2, 0, 0))[2] == "tcgen05_smallc_gather_wgrad"   # x width 259: no TMA on x needed
    assert names(_desc(128, 3, 32, 32, 64, 3, 3, 2, 1, 1))[2] == "tcgen05_smallc_gather_wgrad"
    # ... and on small problems (or other math modes) the CUDA-core strided kernel
    assert names(_desc(2, 3, 32, 32, 64, 3, 3, 2, 1, 1))[2] == "strided_smallc_wgrad"
    assert names(_desc(8, 3, 256, 256, 64, 3, 3, 2, 1, 1, math=lib.MATH_FP32))[2] == "strided_smallc_wgrad"
    # few-filter output convolutions (DCGAN G out, pix2pix last, autoencoder dec3): small-channel kernels with x and y exchanged
    assert names(_desc(128, 128, 32, 32, 3, 3, 3, 1, 1, 1)) == ["smallk_fprop", "smallk_dgrad", "tcgen05_smallk_wgrad"]
    assert names(_desc(8, 128, 256, 256, 3, 3, 3, 1, 1, 1)) == ["smallk_fprop", "smallk_dgrad", "tcgen05_smallk_wgrad"]
    assert names(_desc(256, 16, 28, 28, 1, 3, 3, 1, 1, 1))[:2] == ["smallk_fprop", "smallk_dgrad"]
    # 3xTF32: all three ops on tensor cores -- forward / input gradient on the halo-tile kernel (no row-tap), the kernel gradient on
    # the gathered kernel (one tap per CTA, accumulation chain cut every 8 steps); strided / small-map layers on the gathered kernels
    n3 = names(_desc(8, 64, 512, 512, 64, 3, 3, 1, 1, 1, math=lib.MATH_3XTF32))
    assert n3 == ["tcgen05_fprop", "tcgen05_dgrad", "tcgen05_gather_wgrad"]
    for shape in [(128, 64, 32, 32, 128, 3, 3, 2, 1, 1), (128, 128, 8, 8, 256, 4, 4, 2, 1, 1), (8, 256, 34, 34, 512, 4, 4, 1, 0, 0), (8, 512, 512 // 16, 512 // 16, 512, 3, 3, 1, 1, 1)]:
        n3 = names(_desc(*shape, math=lib.MATH_3XTF32))
        assert all(n.startswith("tcgen05_") for n in n3), (shape, n3)
    # fp32 math never touches the tensor-core families; NHWC runs the NCHW tensor-core kernels between two layout passes
    assert all(n.startswith("direct_") for n in names(_desc(8, 64, 64, 64, 64, 3, 3, 1, 1, 1, math=lib.MATH_FP32)))
    assert names(_desc(8, 64, 64, 64, 64, 3, 3, 1, 1, 1, fmt=lib.NHWC)) == ["tcgen05_rowtap_fprop_nhwc", "tcgen05_rowtap_dgrad_nhwc", "tcgen05_rowfold_wgrad_nhwc"]
    assert names(_desc(8, 64, 32, 32, 128, 3, 3, 2, 1, 1, fmt=lib.NHWC, math=lib.MATH_3XTF32)) == ["tcgen05_gather_fprop_nhwc", "tcgen05_gather_dgrad_nhwc", "tcgen05_gather_wgrad_nhwc"]
    assert all(n.startswith("direct_") for n in names(_desc(8, 64, 64, 64, 64, 3, 3, 1, 1, 1, fmt=lib.NHWC, math=lib.MATH_FP32)))
    assert all(n.startswith("direct_") for n in names(_desc(2, 3, 64, 64, 16, 3, 3, 1, 1, 1, fmt=lib.NHWC)))       # first layers stay on the fp32 path
    # ... and its workspace holds the NCHW copies of both activation tensors behind the twin's own workspace
    nh = _desc(8, 64, 64, 64, 64, 3, 3, 1, 1, 1, fmt=lib.NHWC); nc = _desc(8, 64, 64, 64, 64, 3, 3, 1, 1, 1)
    L.nb200_conv2d_workspace_bytes.restype = ctypes.c_size_t
    assert L.nb200_conv2d_workspace_bytes(lib.OP_FORWARD, ctypes.byref(nh)) >= L.nb200_conv2d_workspace_bytes(lib.OP_FORWARD, ctypes.byref(nc)) + 2 * 8 * 64 * 64 * 64 * 4


def test_workspace_sizes_cover_every_kernel_of_an_op(L):
    """nb200_conv2d_workspace_bytes must cover whichever kernel the dispatcher picks (filter repack for forward / input
    gradient, split-K partials for the kernel gradient)."""
    L.nb200_conv2d_workspace_bytes.restype = ctypes.c_size_t
    d = _desc(8, 64, 512, 512, 64, 3, 3, 1, 1, 1)
    fwd = L.nb200_conv2d_workspace_bytes(lib.OP_FORWARD, ctypes.byref(d))
    assert fwd >= 3 * 192 * 64 * 4 and fwd >= 9 * 64 * 64 * 4        # row-tap layout and the generic [tap][k][c] layout
    wg = L.nb200_conv2d_workspace_bytes(lib.OP_KERNELS_GRADIENT, ctypes.byref(d))
    assert wg >= 147 * 9 * 64 * 64 * 4                                # one partial per CTA of the row-fold kernel (4096 rows / 28)
    d0 = _desc(8, 3, 512, 512, 64, 3, 3, 1, 1, 1)
    assert L.nb200_conv2d_workspace_bytes(lib.OP_KERNELS_GRADIENT, ctypes.byref(d0)) >= 148 * 128 * 32 * 4
    assert L.nb200_conv2d_workspace_bytes(lib.OP_FORWARD, ctypes.byref(d0)) == 0


def test_workspace_sizes_of_the_gathered_and_small_families(L):
    """Host-side plans added for the GAN configs: channel splits keep their partials behind the repacked filters, the kernel
    gradient of odd-sized maps carries a pitched copy of dy, few-filter layers hold the transposed + rotated filters."""
    L.nb200_conv2d_workspace_bytes.restype = ctypes.c_size_t
    ws = lambda op, d: L.nb200_conv2d_workspace_bytes(op, ctypes.byref(d))
    # pix2pix G dec2: 1024 -> 512 on 4x4 maps, batch 8: one pixel tile -> BN 64, 8 channel splits
    d = _desc(8, 1024, 4, 4, 512, 3, 3, 1, 1, 1)
    repack = 9 * 512 * 1024 * 4
    out = 8 * 512 * 4 * 4 * 4
    assert ws(lib.OP_FORWARD, d) >= repack + 8 * out
    assert ws(lib.OP_INPUT_GRADIENT, d) >= repack + 2 * (8 * 1024 * 4 * 4 * 4)
    # a grid that already fills the chip keeps the plain layout (no partials)
    big = _desc(128, 128, 32, 32, 128, 4, 4, 2, 1, 1)
    assert ws(lib.OP_FORWARD, big) == (16 * 128 * 128 * 4 + 255) // 256 * 256
    # PatchGAN 256 -> 512 on 31x31 maps: split-K partials + dy pitched from 961 to 964 floats per plane
    pg = _desc(8, 256, 34, 34, 512, 4, 4, 1, 0, 0)
    assert ws(lib.OP_KERNELS_GRADIENT, pg) >= 16 * 512 * 256 * 4 + 8 * 512 * 964 * 4
    # 1x1 maps (U-Net bottleneck): planes pitched to 4 floats
    bn = _desc(8, 512, 2, 2, 512, 3, 3, 2, 1, 1)
    assert ws(lib.OP_KERNELS_GRADIENT, bn) >= 9 * 512 * 512 * 4 + 8 * 512 * 4 * 4
    # few-filter layer: forward / input gradient hold w' (C x K x 9); the kernel gradient adds dw' and the small-channel partials
    fk = _desc(8, 128, 256, 256, 3, 3, 3, 1, 1, 1)
    assert ws(lib.OP_FORWARD, fk) == ws(lib.OP_INPUT_GRADIENT, fk) == (128 * 3 * 9 * 4 + 255) // 256 * 256
    assert ws(lib.OP_KERNELS_GRADIENT, fk) > ws(lib.OP_FORWARD, fk)
    # strided few-channel kernel gradient: CUDA-core kernel (fp32 math) one partial per slice; tensor-core kernel one [128][32] partial per CTA
    sg = _desc(8, 3, 256, 256, 64, 3, 3, 2, 1, 1, math=lib.MATH_FP32)
    assert ws(lib.OP_KERNELS_GRADIENT, sg) % (64 * 27 * 4) == 0 and ws(lib.OP_KERNELS_GRADIENT, sg) >= 64 * 27 * 4
    sg = _desc(8, 3, 256, 256, 64, 3, 3, 2, 1, 1)
    assert ws(lib.OP_KERNELS_GRADIENT, sg) == 148 * 128 * 32 * 4
    d1 = _desc(8, 6, 259, 259, 64, 4, 4, 2, 0, 0)
    assert ws(lib.OP_KERNELS_GRADIENT, d1) == 148 * 128 * 96 * 4


def test_workspace_bytes_is_what_the_launcher_demands_for_odd_filter_counts(L):
    """R*S*Kout*Cp*4 is only a multiple of 128 for odd K (or odd C in the input gradient); the launcher rounds to 256.
    Both now go through ONE rounded number: a 1x1 / 3x3 conv to 9 or 21 classes reports a 256-byte multiple."""
    L.nb200_conv2d_workspace_bytes.restype = ctypes.c_size_t
    for (C, K, F, p) in [(16, 9, 3, 1), (16, 21, 1, 0), (32, 33, 3, 1), (96, 9, 3, 1)]:
        d = _desc(8, C, 32, 32, K, F, F, 1, p, p)
        assert L.nb200_conv2d_kernel_name(lib.OP_FORWARD, ctypes.byref(d)) == b"tcgen05_fprop"
        need = F * F * K * ((C + 31) // 32 * 32) * 4
        got = L.nb200_conv2d_workspace_bytes(lib.OP_FORWARD, ctypes.byref(d))
        assert got % 256 == 0 and got >= (need + 255) // 256 * 256
    d = _desc(8, 9, 32, 32, 16, 3, 3, 1, 1, 1)       # input gradient: the repacked rows are the C = 9 channels
    assert L.nb200_conv2d_kernel_name(lib.OP_INPUT_GRADIENT, ctypes.byref(d)) == b"tcgen05_dgrad"
    assert L.nb200_conv2d_workspace_bytes(lib.OP_INPUT_GRADIENT, ctypes.byref(d)) % 256 == 0


def test_gradient_ops_validate_their_extents(L):
    """(Ho, Wo) of the gradient ops used to flow unchecked into tensor maps. A gradient larger than the forward op
    can produce for this input is rejected; transposed-conv and ragged-stride extents (smaller or equal) stay valid; products of
    large extents are checked without overflowing."""
    p = ctypes.c_void_p(16)
    ok = lib.ConvDesc(0, 8, 10, 10, 8, 4, 4, 4, 4, 2, 1, 1, lib.NCHW, lib.MATH_FP32)          # empty batch: validation only
    assert L.nb200_conv2d_input_gradient(ctypes.byref(ok), None, None, None, None, 0, None) == 0
    ragged = lib.ConvDesc(2, 8, 11, 11, 8, 4, 4, 4, 4, 2, 0, 0, lib.NCHW, lib.MATH_FP32)      # (11-4)//2+1 = 4: last row/col of dx untouched
    too_big = lib.ConvDesc(2, 8, 10, 10, 8, 3, 3, 11, 10, 1, 1, 1, lib.NCHW, lib.MATH_FP32)   # Ho = 11 > 10
    assert L.nb200_conv2d_input_gradient(ctypes.byref(too_big), p, p, p, None, 0, None) == -1
    assert b"exceeds the output" in L.nb200_last_error()
    assert L.nb200_conv2d_kernels_gradient(ctypes.byref(too_big), p, p, p, None, None, 0, None) == -1
    import torch
    if not torch.cuda.is_available():
        assert L.nb200_conv2d_input_gradient(ctypes.byref(ragged), p, p, p, None, 0, None) == -2   # valid: reaches the device check
    # 46341^2 * 2 overflows int32 products and 65536^2 * 65536 overflows int64 if multiplied blindly
    huge = lib.ConvDesc(65536, 65536, 65536, 65536, 1, 1, 1, 65536, 65536, 1, 0, 0, lib.NCHW, lib.MATH_FP32)
    assert L.nb200_conv2d_forward(ctypes.byref(huge), p, p, None, 0, 0.0, p, None, 0, None) == -1
    assert b"2^32-1" in L.nb200_last_error()
    hb = lib.ConvDesc(2 ** 31 - 1, 0, 0, 0, 2 ** 31 - 1, 1, 1, 2 ** 31 - 1, 2 ** 31 - 1, 1, 0, 0, lib.NCHW, lib.MATH_FP32)
    assert L.nb200_conv2d_bias_gradient(ctypes.byref(hb), p, p, None) == -1


def test_batch_norm_entry_points_are_host_validated(L):
    """EBatchNormMode layouts (G statistics), workspace, argument checks, and no CPU fallback."""
    g = lambda N, C, H, W, mode: L.nb200_batch_norm_groups(ctypes.byref(lib.BnDesc(N, C, H, W, mode)))
    assert g(6, 5, 4, 3, lib.BN_SPATIAL) == 5 and g(6, 5, 4, 3, lib.BN_PER_ACTIVATION) == 60 and g(6, 5, 4, 3, lib.BN_INSTANCE) == 30
    d = lib.BnDesc(8, 64, 128, 128, lib.BN_SPATIAL)
    # 131072 elements per channel = 32 blocks of 4096; 3 floats per block + one row of per-channel sums
    assert L.nb200_batch_norm_workspace_bytes(ctypes.byref(d)) == (64 * 32 * 3 + 64 * 3) * 4
    bad = lib.BnDesc(8, 64, 128, 128, 5)
    p = ctypes.c_void_p(16)
    assert L.nb200_batch_norm_train(ctypes.byref(bad), p, p, p, 0.9, 1e-3, None, None, p, p, p, p, 1 << 20, None) == -1
    assert L.nb200_batch_norm_train(ctypes.byref(d), None, p, p, 0.9, 1e-3, None, None, p, p, p, p, 1 << 20, None) == -1
    assert L.nb200_batch_norm_train(ctypes.byref(d), p, p, p, 0.9, 1e-3, None, None, p, p, p, p, 16, None) == -4     # workspace too small
    assert L.nb200_batch_norm_train_from_moments(ctypes.byref(d), p, 0, p, p, p, 0.9, 1e-3, None, None, p, p, p, None) == -1
    empty = lib.BnDesc(0, 64, 8, 8, lib.BN_SPATIAL)
    assert L.nb200_batch_norm_train(ctypes.byref(empty), None, None, None, 0.9, 1e-3, None, None, None, None, None, None, 0, None) == 0
    import torch
    if not torch.cuda.is_available():
        assert L.nb200_batch_norm_train(ctypes.byref(d), p, p, p, 0.9, 1e-3, None, None, p, p, p, p, 1 << 20, None) == -2
        assert L.nb200_batch_norm_gradient(ctypes.byref(d), p, p, p, p, p, p, p, p, p, 1 << 20, None) == -2
        assert L.nb200_batch_norm(ctypes.byref(d), p, p, p, 1e-3, p, p, p, None) == -2


def test_plan_entry_points_are_host_validated(L):
    """nb200_conv2d_plan_create validates like the plain calls and needs a device (no CPU fallback); NULL plans are harmless."""
    d = lib.ConvDesc(1, 2, 6, 6, 1, 3, 3, 4, 4, 1, 0, 0, lib.NCHW, lib.MATH_FP32)
    bad = lib.ConvDesc(1, 2, 6, 6, 1, 3, 3, 5, 4, 1, 0, 0, lib.NCHW, lib.MATH_FP32)
    p = ctypes.c_void_p(16)
    h = ctypes.c_void_p()
    assert L.nb200_conv2d_plan_create(lib.OP_FORWARD, ctypes.byref(bad), p, p, p, None, 0, 0.0, 0, None, 0, ctypes.byref(h)) == -1 and not h.value
    assert L.nb200_conv2d_plan_create(7, ctypes.byref(d), p, p, p, None, 0, 0.0, 0, None, 0, ctypes.byref(h)) == -1
    assert L.nb200_conv2d_plan_create(lib.OP_FORWARD, ctypes.byref(d), p, p, p, None, 0, 0.0, 0, None, 0, None) == -1
    assert L.nb200_conv2d_plan_run(None, None) == -1 and L.nb200_conv2d_plan_kernels(None) == 0
    L.nb200_conv2d_plan_destroy(None)
    import torch
    if not torch.cuda.is_available():
        assert L.nb200_conv2d_plan_create(lib.OP_FORWARD, ctypes.byref(d), p, p, p, None, 0, 0.0, 0, None, 0, ctypes.byref(h)) == -2
        assert L.nb200_conv2d_plan_create(lib.OP_KERNELS_GRADIENT, ctypes.byref(d), p, p, p, None, 0, 0.0, 1, None, 0, ctypes.byref(h)) in (-1, -2)


def test_bench_flop_tables_are_consistent():
    """bench.py counts 2*N*K*Ho*Wo*C*R*S per conv op; its layer table, the trainer's VGG16 stack and the per-image total
    (160.4 GFLOP per op at 512x512) must agree, and the gradient buckets must tile the flat buffer from the end."""
    import torch
    import bench
    from neuro__b200.fit import ConvLayerSpec, ConvStackTrainer
    from tests.oracle_op import OracleOp
    assert abs(bench.SAMPLE_FLOPS_PER_OP / 1e9 - 160.4) < 0.1
    assert abs(bench.vgg_flops_per_image("vgg16", 512) - bench.SAMPLE_FLOPS_PER_OP) < 1.0
    assert abs(bench.stack_flops((3, 512, 512), bench.vgg_layers("vgg16")) - bench.SAMPLE_FLOPS_PER_OP) < 1.0
    assert bench.vgg_flops_per_image("vgg19", 512) > bench.SAMPLE_FLOPS_PER_OP
    layers = bench.vgg_layers("vgg16")
    assert sum(isinstance(l, ConvLayerSpec) for l in layers) == 13 and len(layers) == 18
    tr = ConvStackTrainer(OracleOp(), (3, 32, 32), layers, torch.device("cpu"), bucket_bytes=24 << 20)
    assert tr.out_shape == (512, 1, 1) and tr.params.numel() == 14710464 + 4224       # VGG16 conv kernels + biases
    assert [b[1] - b[0] for b in tr.buckets] == [7079424, 6489856, 885248, 260160]    # 28.3 + 26.0 + 3.5 MB filled from the end + a 1.0 MB tail (blocks 1-2)
    assert [b[2] for b in tr.buckets] == [14, 8, 6, 0]                                  # the layer whose kernel gradient completes each bucket
    assert tr.buckets[0][1] == tr.grads.numel() and tr.buckets[-1][0] == 0
    for name, (in_shape, stack, batch) in bench.model_stacks().items():
        t = ConvStackTrainer(OracleOp(), in_shape, stack, torch.device("cpu"))
        assert all(v > 0 for v in t.out_shape) and bench.stack_flops(in_shape, stack) > 0, name
