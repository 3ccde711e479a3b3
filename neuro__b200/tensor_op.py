"""TensorOpB200 -- Python mirror of the convolution slice of Neuro::TensorOpCpu.

Method names, argument order and meaning follow the reference virtuals
(Neuro/include/Tensors/TensorOpCpu.h:46-50) and the Tensor-level wrappers that call them
(Neuro/src/Tensors/Tensor.cpp:1757-1830), so parity tests read like the reference's own
(Neuro.Tests/src/TensorOpGpuTests.cpp:1196-1322). Tensors are torch CUDA float32 tensors used purely as
device memory: NCHW tensors have shape (N,C,H,W), NHWC (N,H,W,C), kernels (K,C,R,S). Outputs are pre-sized by
the caller and overwritten, as in the reference. All compute goes through the C ABI; there is no
torch/CPU fallback.
"""
import ctypes

import torch

from . import lib
from .lib import ConvDesc, NCHW, NHWC, check


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "device-resident contiguous fp32 expected"
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _act_extent(fmt, t):
    if fmt == NCHW:
        n, c, h, w = t.shape
    else:
        n, h, w, c = t.shape
    return n, c, h, w


def get_padding(mode, kernel_size):
    """Tensor::GetPadding (Tensor.cpp:1966-1985). mode: 'valid' | 'same' | 'full'."""
    return lib.load().nb200_padding({"valid": 0, "same": 1, "full": 2}[mode], kernel_size)


def get_conv_output_shape(in_shape, kernels_num, kernel_w, kernel_h, stride, padding_x, padding_y, fmt=NCHW):
    """Tensor::GetConvOutputShape (Tensor.cpp:2010-2029); shapes are torch-order tuples."""
    L = lib.load()
    n, c, h, w = in_shape if fmt == NCHW else (in_shape[0], in_shape[3], in_shape[1], in_shape[2])
    ho = L.nb200_conv_out_size(h, kernel_h, stride, padding_y)
    wo = L.nb200_conv_out_size(w, kernel_w, stride, padding_x)
    return (n, kernels_num, ho, wo) if fmt == NCHW else (n, ho, wo, kernels_num)


def get_conv_transpose_output_shape(in_shape, output_depth, kernel_w, kernel_h, stride, padding_x, padding_y, fmt=NCHW):
    """Tensor::GetConvTransposeOutputShape (Tensor.cpp:2032-2051)."""
    L = lib.load()
    n, c, h, w = in_shape if fmt == NCHW else (in_shape[0], in_shape[3], in_shape[1], in_shape[2])
    ho = L.nb200_conv_transpose_out_size(h, kernel_h, stride, padding_y)
    wo = L.nb200_conv_transpose_out_size(w, kernel_w, stride, padding_x)
    return (n, output_depth, ho, wo) if fmt == NCHW else (n, ho, wo, output_depth)


class PreparedKernels:
    """Filters of one layer repacked once for one op (nb200_conv2d_prepare_filters): the handle owns the workspace the
    *_prepared calls read them from. Valid while the kernels tensor is unchanged (inference, style transfer)."""

    def __init__(self, op, desc, kernels, ws, nbytes):
        self.op, self.desc, self.kernels, self.ws, self.nbytes = op, desc, kernels, ws, nbytes

    def matches(self, op, d, kernels):
        return (self.op == op and kernels.data_ptr() == self.kernels.data_ptr()
                and all(getattr(d, f) == getattr(self.desc, f) for f, _ in ConvDesc._fields_))

    def ptr(self):
        return (ctypes.c_void_p(self.ws.data_ptr()) if self.ws is not None else None), self.nbytes


class ConvPlan:
    """nb200_conv2d_plan_*: one (op, descriptor) bound to fixed device tensors, its launches recorded once; run() is one
    cudaGraphLaunch with no host-side encode / repack. Owns its workspace; keeps the tensors alive."""

    def __init__(self, L, op, d, a, b, out, bias=None, act=lib.ACT_IDENTITY, alpha=0.0, filters_constant=False):
        self._L, self.tensors = L, (a, b, out, bias)
        need = L.nb200_conv2d_workspace_bytes(op, ctypes.byref(d))
        self.ws = torch.empty(max(need, 16), dtype=torch.uint8, device=a.device)
        self.handle = ctypes.c_void_p()
        check(L.nb200_conv2d_plan_create(op, ctypes.byref(d), _ptr(a), _ptr(b), _ptr(out), _ptr(bias), act, alpha, 1 if filters_constant else 0,
                                         ctypes.c_void_p(self.ws.data_ptr()), need, ctypes.byref(self.handle)))

    def run(self):
        check(self._L.nb200_conv2d_plan_run(self.handle, _stream()))

    @property
    def kernels(self):
        return self._L.nb200_conv2d_plan_kernels(self.handle)

    def __del__(self):
        if getattr(self, "handle", None) and self.handle.value:
            self._L.nb200_conv2d_plan_destroy(self.handle)
            self.handle = ctypes.c_void_p()


class TensorOpB200:
    """Stateless apart from a grow-only device workspace per instance (reference: pooled workspace,
    TensorOpGpu.cpp:648)."""

    def __init__(self, math=lib.MATH_TF32):
        self.math = math
        self._ws = None
        self._bnws = None
        self._L = lib.load()

    # -- helpers
    def _desc(self, fmt, x_like, k_like, y_like, stride, padding_x, padding_y):
        N, C, H, W = _act_extent(fmt, x_like)
        K, C2, R, S = k_like.shape
        N2, K2, Ho, Wo = _act_extent(fmt, y_like)
        assert C2 == C and K2 == K and N2 == N, "tensor shapes do not describe one convolution"
        return ConvDesc(N, C, H, W, K, R, S, Ho, Wo, stride, padding_x, padding_y, fmt, self.math)

    def _workspace(self, op, d):
        need = self._L.nb200_conv2d_workspace_bytes(op, ctypes.byref(d))
        if need == 0:
            return None, 0
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device="cuda")
        return ctypes.c_void_p(self._ws.data_ptr()), need

    def kernel_name(self, op, d):
        return self._L.nb200_conv2d_kernel_name(op, ctypes.byref(d)).decode()

    # -- constant filters: repack once, reuse (no reference counterpart; the reference re-derives cuDNN descriptors and
    #    algorithms on every call, TensorOpGpu.cpp:629-668)
    def PrepareKernels(self, op, inputLike, kernels, outputLike, stride, paddingX, paddingY, dataFormat=NCHW):
        """op: lib.OP_FORWARD (inputLike = input, outputLike = output) or lib.OP_INPUT_GRADIENT (inputLike = the input
        gradient, outputLike = the incoming gradient). Returns a PreparedKernels to pass as `prepared=`."""
        d = self._desc(dataFormat, inputLike, kernels, outputLike, stride, paddingX, paddingY)
        need = self._L.nb200_conv2d_workspace_bytes(op, ctypes.byref(d))
        ws = torch.empty(need, dtype=torch.uint8, device="cuda") if need else None
        h = PreparedKernels(op, d, kernels, ws, need)
        p, n = h.ptr()
        check(self._L.nb200_conv2d_prepare_filters(op, ctypes.byref(d), _ptr(kernels), p, n, _stream()))
        return h

    # -- plans (include/neuro_b200.h: nb200_conv2d_plan_create)
    def PlanConv2DBiasActivation(self, input, kernels, stride, paddingX, paddingY, bias, activation, activationAlpha, output, dataFormat=NCHW,
                                 filtersConstant=False):
        d = self._desc(dataFormat, input, kernels, output, stride, paddingX, paddingY)
        return ConvPlan(self._L, lib.OP_FORWARD, d, input, kernels, output, bias, activation, activationAlpha, filtersConstant)

    def PlanConv2DInputGradient(self, gradient, kernels, stride, paddingX, paddingY, dataFormat, inputGradient, filtersConstant=False):
        d = self._desc(dataFormat, inputGradient, kernels, gradient, stride, paddingX, paddingY)
        return ConvPlan(self._L, lib.OP_INPUT_GRADIENT, d, gradient, kernels, inputGradient, filters_constant=filtersConstant)

    def PlanConv2DKernelsGradient(self, input, gradient, stride, paddingX, paddingY, dataFormat, kernelsGradient, biasGradient=None):
        d = self._desc(dataFormat, input, kernelsGradient, gradient, stride, paddingX, paddingY)
        return ConvPlan(self._L, lib.OP_KERNELS_GRADIENT, d, input, gradient, kernelsGradient, biasGradient)

    # -- the op interface (TensorOpCpu.h:46-50)
    def Conv2D(self, input, kernels, stride, paddingX, paddingY, dataFormat, output, prepared=None):
        self._forward(input, kernels, stride, paddingX, paddingY, dataFormat, None, lib.ACT_IDENTITY, 0.0, output, prepared)

    def Conv2DBiasActivation(self, input, kernels, stride, paddingX, paddingY, bias, activation, activationAlpha, output,
                             dataFormat=NCHW, prepared=None):
        assert paddingX == paddingY  # TensorOpCpu.cpp:1057
        self._forward(input, kernels, stride, paddingX, paddingY, dataFormat, bias, activation, activationAlpha, output, prepared)

    def _forward(self, input, kernels, stride, paddingX, paddingY, dataFormat, bias, activation, alpha, output, prepared):
        d = self._desc(dataFormat, input, kernels, output, stride, paddingX, paddingY)
        if prepared is not None:
            assert prepared.matches(lib.OP_FORWARD, d, kernels), "PreparedKernels belong to another problem"
            ws, n = prepared.ptr()
            fn = self._L.nb200_conv2d_forward_prepared
        else:
            ws, n = self._workspace(lib.OP_FORWARD, d)
            fn = self._L.nb200_conv2d_forward
        check(fn(ctypes.byref(d), _ptr(input), _ptr(kernels), _ptr(bias), activation, alpha, _ptr(output), ws, n, _stream()))

    def ActivationGradient(self, activation, activationAlpha, output, outputGradient, inputGradient, dataFormat=NCHW):
        """Tensor::ActivationGradient -> TensorOpCpu::{Sigmoid,Tanh,ReLU,Elu,LeakyReLU}Gradient (TensorOpCpu.cpp:813-864)."""
        self.Conv2DBiasActivationGradient(output, outputGradient, activation, activationAlpha, inputGradient, None, dataFormat)

    def Conv2DBiasActivationGradient(self, output, outputGradient, activation, activationAlpha, activationInputGradient,
                                     biasGradient=None, dataFormat=NCHW):
        """Backward prologue of Conv2dBiasActivationOp (Conv2dBiasActivationOp.cpp:47-60) in one pass:
        activationInputGradient = act'(output) * outputGradient and biasGradient = its sum over N,H,W."""
        N, K, Ho, Wo = _act_extent(dataFormat, outputGradient)
        assert output.shape == outputGradient.shape == activationInputGradient.shape
        d = ConvDesc(N, 0, 0, 0, K, 1, 1, Ho, Wo, 1, 0, 0, dataFormat, self.math)
        need = self._L.nb200_conv2d_bias_activation_gradient_workspace_bytes(ctypes.byref(d)) if biasGradient is not None else 0
        ws = None
        if need:
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device="cuda")
            ws = ctypes.c_void_p(self._ws.data_ptr())
        check(self._L.nb200_conv2d_bias_activation_gradient(ctypes.byref(d), activation, activationAlpha, _ptr(output),
                                                            _ptr(outputGradient), _ptr(activationInputGradient),
                                                            _ptr(biasGradient), ws, need, _stream()))

    def BiasActivation(self, input, bias, activation, activationAlpha, output, dataFormat=NCHW):
        """output = act(input + bias): AddOp (AddOp.cpp:38-50) + Tensor::Activation (TensorOpCpu.cpp:807-864) in one pass;
        bias may be None, output may be input."""
        N, K, Ho, Wo = _act_extent(dataFormat, input)
        d = ConvDesc(N, 0, 0, 0, K, 1, 1, Ho, Wo, 1, 0, 0, dataFormat, self.math)
        check(self._L.nb200_bias_activation(ctypes.byref(d), _ptr(input), _ptr(bias), activation, activationAlpha, _ptr(output), _stream()))

    def Conv2DBiasGradient(self, gradient, biasGradient, dataFormat=NCHW):
        N, K, Ho, Wo = _act_extent(dataFormat, gradient)
        d = ConvDesc(N, 0, 0, 0, K, 1, 1, Ho, Wo, 1, 0, 0, dataFormat, self.math)
        check(self._L.nb200_conv2d_bias_gradient(ctypes.byref(d), _ptr(gradient), _ptr(biasGradient), _stream()))

    def Conv2DInputGradient(self, gradient, kernels, stride, paddingX, paddingY, dataFormat, inputGradient, prepared=None):
        d = self._desc(dataFormat, inputGradient, kernels, gradient, stride, paddingX, paddingY)
        if prepared is not None:
            assert prepared.matches(lib.OP_INPUT_GRADIENT, d, kernels), "PreparedKernels belong to another problem"
            ws, n = prepared.ptr()
            fn = self._L.nb200_conv2d_input_gradient_prepared
        else:
            ws, n = self._workspace(lib.OP_INPUT_GRADIENT, d)
            fn = self._L.nb200_conv2d_input_gradient
        check(fn(ctypes.byref(d), _ptr(gradient), _ptr(kernels), _ptr(inputGradient), ws, n, _stream()))

    def Conv2DKernelsGradient(self, input, gradient, stride, paddingX, paddingY, dataFormat, kernelsGradient,
                              biasGradient=None):
        d = self._desc(dataFormat, input, kernelsGradient, gradient, stride, paddingX, paddingY)
        ws, n = self._workspace(lib.OP_KERNELS_GRADIENT, d)
        check(self._L.nb200_conv2d_kernels_gradient(ctypes.byref(d), _ptr(input), _ptr(gradient), _ptr(kernelsGradient),
                                                    _ptr(biasGradient), ws, n, _stream()))

    # -- Tensor-level identities for transposed convolution (Tensor.cpp:1806-1830)
    def Conv2DTransposed(self, input, kernels, stride, padding, dataFormat, result):
        """forward of Conv2DTranspose = input gradient of the matching conv; kernels are (inDepth, outDepth, F, F)."""
        self.Conv2DInputGradient(input, kernels, stride, padding, padding, dataFormat, result)

    def Conv2DTransposedInputsGradient(self, gradient, kernels, stride, padding, dataFormat, inputsGradient):
        self.Conv2D(gradient, kernels, stride, padding, padding, dataFormat, inputsGradient)

    def Conv2DTransposedKernelsGradient(self, input, gradient, stride, padding, dataFormat, kernelsGradient):
        """note the swapped (gradient, input) order, as in Tensor.cpp:1827-1830"""
        self.Conv2DKernelsGradient(gradient, input, stride, padding, padding, dataFormat, kernelsGradient)

    # -- spatial resamplers around the convolutions (TensorOpCpu.h:51-54; ConstantPad2D TensorOpCpu.cpp:528)
    def _pool_desc(self, input, filterSize, stride, type, paddingX, paddingY, dataFormat, output):
        N, C, H, W = _act_extent(dataFormat, input)
        N2, C2, Ho, Wo = _act_extent(dataFormat, output)
        assert N2 == N and C2 == C
        return lib.PoolDesc(N, C, H, W, Ho, Wo, filterSize, stride, paddingX, paddingY, type, dataFormat)

    def Pool2D(self, input, filterSize, stride, type, paddingX, paddingY, dataFormat, output):
        d = self._pool_desc(input, filterSize, stride, type, paddingX, paddingY, dataFormat, output)
        check(self._L.nb200_pool2d(ctypes.byref(d), _ptr(input), _ptr(output), _stream()))

    def Pool2DGradient(self, output, input, outputGradient, filterSize, stride, type, paddingX, paddingY, dataFormat, inputGradient):
        d = self._pool_desc(input, filterSize, stride, type, paddingX, paddingY, dataFormat, output)
        check(self._L.nb200_pool2d_gradient(ctypes.byref(d), _ptr(output), _ptr(input), _ptr(outputGradient), _ptr(inputGradient), _stream()))

    def Pool2DGradientActivationSupported(self, input, filterSize, stride, type, paddingX, paddingY, dataFormat, output):
        d = self._pool_desc(input, filterSize, stride, type, paddingX, paddingY, dataFormat, output)
        return bool(self._L.nb200_pool2d_gradient_activation_supported(ctypes.byref(d))) and input.data_ptr() % 16 == 0 and output.data_ptr() % 16 == 0

    def Pool2DGradientActivation(self, output, input, outputGradient, filterSize, stride, type, paddingX, paddingY, dataFormat, activation,
                                 activationAlpha, activationInputGradient, biasGradient=None):
        """Pool2DGradient followed by the ActivationGradient (+ Conv2DBiasGradient) of the fused conv layer that produced `input`, in
        one pass: activationInputGradient = act'(input) * poolGradient, biasGradient = its sum over N,H,W."""
        d = self._pool_desc(input, filterSize, stride, type, paddingX, paddingY, dataFormat, output)
        need = self._L.nb200_pool2d_gradient_activation_workspace_bytes(ctypes.byref(d)) if biasGradient is not None else 0
        ws = None
        if need:
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device="cuda")
            ws = ctypes.c_void_p(self._ws.data_ptr())
        check(self._L.nb200_pool2d_gradient_activation(ctypes.byref(d), activation, activationAlpha, _ptr(output), _ptr(input), _ptr(outputGradient),
                                                       _ptr(activationInputGradient), _ptr(biasGradient), ws, need, _stream()))

    def UpSample2D(self, input, scaleFactor, output):
        N, C, H, W = input.shape
        assert tuple(output.shape) == (N, C, H * scaleFactor, W * scaleFactor)   # Tensor.cpp:1859
        check(self._L.nb200_upsample2d(N, C, H, W, scaleFactor, _ptr(input), _ptr(output), _stream()))

    def UpSample2DGradient(self, outputGradient, scaleFactor, inputGradient):
        N, C, H, W = inputGradient.shape
        assert tuple(outputGradient.shape) == (N, C, H * scaleFactor, W * scaleFactor)   # Tensor.cpp:1874
        check(self._L.nb200_upsample2d_gradient(N, C, H, W, scaleFactor, _ptr(outputGradient), _ptr(inputGradient), _stream()))

    def ConstantPad2D(self, input, left, right, top, bottom, value, output):
        N, C, H, W = input.shape
        assert tuple(output.shape) == (N, C, H + top + bottom, W + left + right)   # Tensor.cpp:1502
        check(self._L.nb200_constant_pad2d(N, C, H, W, left, right, top, bottom, value, _ptr(input), _ptr(output), _stream()))

    # -- batch normalisation (TensorOpCpu.h:55-57); gamma / beta / statistics are flat tensors of G values
    def _bn(self, input, mode):
        N, C, H, W = input.shape
        d = lib.BnDesc(N, C, H, W, mode)
        need = self._L.nb200_batch_norm_workspace_bytes(ctypes.byref(d))
        if self._bnws is None or self._bnws.numel() < need:
            self._bnws = torch.empty(max(need, 16), dtype=torch.uint8, device="cuda")
        return d, ctypes.c_void_p(self._bnws.data_ptr()), need

    def BatchNormalization(self, input, mode, gamma, beta, epsilon, runningMean, runningVar, output):
        d, _, _ = self._bn(input, mode)
        check(self._L.nb200_batch_norm(ctypes.byref(d), _ptr(input), _ptr(gamma), _ptr(beta), epsilon, _ptr(runningMean),
                                       _ptr(runningVar), _ptr(output), _stream()))

    def BatchNormalizationTrain(self, input, mode, gamma, beta, momentum, epsilon, runningMean, runningVar, saveMean,
                                saveInvVariance, output):
        d, ws, n = self._bn(input, mode)
        check(self._L.nb200_batch_norm_train(ctypes.byref(d), _ptr(input), _ptr(gamma), _ptr(beta), momentum, epsilon,
                                             _ptr(runningMean), _ptr(runningVar), _ptr(saveMean), _ptr(saveInvVariance),
                                             _ptr(output), ws, n, _stream()))

    def BatchNormalizationGradient(self, input, mode, gamma, epsilon, outputGradient, savedMean, savedInvVariance,
                                   gammaGradient, betaGradient, trainable, inputGradient):
        """`epsilon` and `trainable` are accepted for signature parity; the reference ignores both (TensorOpCpu.cpp:1437-1480)."""
        d, ws, n = self._bn(input, mode)
        check(self._L.nb200_batch_norm_gradient(ctypes.byref(d), _ptr(input), _ptr(gamma), _ptr(outputGradient), _ptr(savedMean),
                                                _ptr(savedInvVariance), _ptr(gammaGradient), _ptr(betaGradient),
                                                _ptr(inputGradient), ws, n, _stream()))

    # the same ops split around the replicas' exchange (include/neuro_b200.h: nb200_batch_norm_moments ...)
    def BatchNormalizationMoments(self, input, mode, moments):
        d, ws, n = self._bn(input, mode)
        check(self._L.nb200_batch_norm_moments(ctypes.byref(d), _ptr(input), _ptr(moments), ws, n, _stream()))

    def BatchNormalizationTrainFromMoments(self, allMoments, replicas, input, mode, gamma, beta, momentum, epsilon, runningMean,
                                           runningVar, saveMean, saveInvVariance, output):
        d, _, _ = self._bn(input, mode)
        check(self._L.nb200_batch_norm_train_from_moments(ctypes.byref(d), _ptr(allMoments), replicas, _ptr(input), _ptr(gamma),
                                                          _ptr(beta), momentum, epsilon, _ptr(runningMean), _ptr(runningVar),
                                                          _ptr(saveMean), _ptr(saveInvVariance), _ptr(output), _stream()))

    def BatchNormalizationGradientSums(self, input, mode, outputGradient, savedMean, sums):
        d, ws, n = self._bn(input, mode)
        check(self._L.nb200_batch_norm_gradient_sums(ctypes.byref(d), _ptr(input), _ptr(outputGradient), _ptr(savedMean), _ptr(sums),
                                                     ws, n, _stream()))

    def BatchNormalizationGradientFromSums(self, replicas, globalSums, localSums, input, mode, gamma, outputGradient, savedMean,
                                           savedInvVariance, gammaGradient, betaGradient, inputGradient):
        d, _, _ = self._bn(input, mode)
        check(self._L.nb200_batch_norm_gradient_from_sums(ctypes.byref(d), replicas, _ptr(globalSums), _ptr(localSums), _ptr(input),
                                                          _ptr(gamma), _ptr(outputGradient), _ptr(savedMean), _ptr(savedInvVariance),
                                                          _ptr(gammaGradient), _ptr(betaGradient), _ptr(inputGradient), _stream()))

    # -- optimiser tail (TensorOpCpu.h:75-76)
    def AdamStep(self, parameter, gradient, mGrad, vGrad, lr, beta1, beta2, epsilon, gradScale=1.0):
        check(self._L.nb200_adam_step(_ptr(parameter), _ptr(gradient), _ptr(mGrad), _ptr(vGrad), parameter.numel(),
                                      gradScale, lr, beta1, beta2, epsilon, _stream()))

    def SgdStep(self, parameter, gradient, lr, gradScale=1.0):
        check(self._L.nb200_sgd_step(_ptr(parameter), _ptr(gradient), parameter.numel(), gradScale, lr, _stream()))
