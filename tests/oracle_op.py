"""Test double: the TensorOpB200 method set backed by the CPU oracle on torch CPU tensors.

Lets the host-side logic that sits above the C ABI (neuro__b200/fit.py: sharding, bucket layout, all-reduce,
optimiser sequencing) run on a GPU-less box with world_size > 1 (gloo). Test infrastructure only."""
import numpy as np
import torch

from oracle import oracle as O


def _np(t):
    return np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32)


class OracleOp:
    def Conv2DBiasActivation(self, input, kernels, stride, paddingX, paddingY, bias, activation, activationAlpha, output):
        output.copy_(torch.from_numpy(O.conv2d_bias_activation(_np(input), _np(kernels), _np(bias), stride, paddingX, activation,
                                                               activationAlpha)))

    def Conv2D(self, input, kernels, stride, paddingX, paddingY, dataFormat, output):
        output.copy_(torch.from_numpy(O.conv2d(_np(input), _np(kernels), stride, paddingX, paddingY, dataFormat)))

    def Conv2DInputGradient(self, gradient, kernels, stride, paddingX, paddingY, dataFormat, inputGradient):
        hw = tuple(inputGradient.shape[2:])
        inputGradient.copy_(torch.from_numpy(O.conv2d_input_gradient(_np(gradient), _np(kernels), stride, paddingX, paddingY, hw, dataFormat)))

    def Conv2DKernelsGradient(self, input, gradient, stride, paddingX, paddingY, dataFormat, kernelsGradient, biasGradient=None):
        rs = tuple(kernelsGradient.shape[2:])
        kernelsGradient.copy_(torch.from_numpy(O.conv2d_kernels_gradient(_np(input), _np(gradient), stride, paddingX, paddingY, rs, dataFormat)))
        if biasGradient is not None:
            biasGradient.copy_(torch.from_numpy(O.conv2d_bias_gradient(_np(gradient), dataFormat)))

    def Conv2DBiasActivationGradient(self, output, outputGradient, activation, activationAlpha, activationInputGradient,
                                     biasGradient=None, dataFormat=0):
        # the reference's two passes (Conv2dBiasActivationOp.cpp:47-60): ActivationGradient, then Conv2DBiasGradient of its result
        dz = O.activation_gradient(activation, activationAlpha, _np(output), _np(outputGradient))
        activationInputGradient.copy_(torch.from_numpy(dz))
        if biasGradient is not None:
            biasGradient.copy_(torch.from_numpy(O.conv2d_bias_gradient(dz, dataFormat)))

    def AdamStep(self, parameter, gradient, mGrad, vGrad, lr, beta1, beta2, epsilon, gradScale=1.0):
        p, g, m, v = _np(parameter), _np(gradient) * np.float32(gradScale), _np(mGrad), _np(vGrad)
        O.adam_step(p, g, m, v, lr, beta1, beta2, epsilon)
        parameter.copy_(torch.from_numpy(p)); mGrad.copy_(torch.from_numpy(m)); vGrad.copy_(torch.from_numpy(v))

    def SgdStep(self, parameter, gradient, lr, gradScale=1.0):
        p, g = _np(parameter), _np(gradient) * np.float32(gradScale)
        O.sgd_step(p, g, lr)
        parameter.copy_(torch.from_numpy(p))
