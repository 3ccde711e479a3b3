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

    # ---- layers around the convolutions (oracle restatements of TensorOpCpu.cpp:528-546, 807-864, 1187-1369)
    def BiasActivation(self, input, bias, activation, activationAlpha, output, dataFormat=0):
        x = _np(input)
        if bias is not None:
            x = x + _np(bias)[None, :, None, None]
        K = x.shape[1]
        # identity 1x1 convolution through the oracle's fused epilogue = its activation code path
        eye = np.zeros((K, K, 1, 1), np.float32); eye[np.arange(K), np.arange(K), 0, 0] = 1
        output.copy_(torch.from_numpy(O.conv2d_bias_activation(np.ascontiguousarray(x, np.float32), eye, np.zeros(K, np.float32), 1, 0,
                                                               activation, activationAlpha)))

    def Pool2D(self, input, filterSize, stride, type, paddingX, paddingY, dataFormat, output):
        output.copy_(torch.from_numpy(O.pool2d(_np(input), filterSize, stride, type, paddingX, paddingY, dataFormat)))

    def Pool2DGradient(self, output, input, outputGradient, filterSize, stride, type, paddingX, paddingY, dataFormat, inputGradient):
        inputGradient.copy_(torch.from_numpy(O.pool2d_gradient(_np(output), _np(input), _np(outputGradient), filterSize, stride, type,
                                                               paddingX, paddingY, dataFormat)))

    def UpSample2D(self, input, scaleFactor, output):
        output.copy_(torch.from_numpy(O.upsample2d(_np(input), scaleFactor)))

    def UpSample2DGradient(self, outputGradient, scaleFactor, inputGradient):
        inputGradient.copy_(torch.from_numpy(O.upsample2d_gradient(_np(outputGradient), scaleFactor)))

    def Conv2DBiasGradient(self, gradient, biasGradient, dataFormat=0):
        biasGradient.copy_(torch.from_numpy(O.conv2d_bias_gradient(_np(gradient), dataFormat)))

    def Conv2DTransposed(self, input, kernels, stride, padding, dataFormat, result):
        self.Conv2DInputGradient(input, kernels, stride, padding, padding, dataFormat, result)

    def Conv2DTransposedInputsGradient(self, gradient, kernels, stride, padding, dataFormat, inputsGradient):
        self.Conv2D(gradient, kernels, stride, padding, padding, dataFormat, inputsGradient)

    def Conv2DTransposedKernelsGradient(self, input, gradient, stride, padding, dataFormat, kernelsGradient):
        self.Conv2DKernelsGradient(gradient, input, stride, padding, padding, dataFormat, kernelsGradient)

    # ---- batch normalisation: the single-device ops are the oracle's; the replica protocol (moments / sums around the exchange,
    #      include/neuro_b200.h) is restated in float64 numpy so that world_size > 1 host logic can be tested without a GPU
    def BatchNormalizationTrain(self, input, mode, gamma, beta, momentum, epsilon, runningMean, runningVar, saveMean, saveInvVariance, output):
        rm, rv = _np(runningMean), _np(runningVar)
        y, sm, sv = O.batch_norm_train(mode, _np(input), _np(gamma), _np(beta), momentum, epsilon, rm, rv)
        for dst, src in ((output, y), (saveMean, sm), (saveInvVariance, sv), (runningMean, rm), (runningVar, rv)):
            dst.copy_(torch.from_numpy(src))

    def BatchNormalizationGradient(self, input, mode, gamma, epsilon, outputGradient, savedMean, savedInvVariance, gammaGradient, betaGradient,
                                   trainable, inputGradient):
        dx, dg, db = O.batch_norm_gradient(mode, _np(input), _np(gamma), _np(outputGradient), _np(savedMean), _np(savedInvVariance))
        inputGradient.copy_(torch.from_numpy(dx)); gammaGradient.copy_(torch.from_numpy(dg)); betaGradient.copy_(torch.from_numpy(db))

    def BatchNormalizationMoments(self, input, mode, moments):
        x = _np(input).astype(np.float64)
        mean = x.mean(axis=(0, 2, 3)); m2 = ((x - mean[None, :, None, None]) ** 2).sum(axis=(0, 2, 3))
        moments.copy_(torch.from_numpy(np.stack([mean, m2], axis=1).astype(np.float32)))

    def BatchNormalizationTrainFromMoments(self, allMoments, replicas, input, mode, gamma, beta, momentum, epsilon, runningMean, runningVar,
                                           saveMean, saveInvVariance, output):
        x = _np(input).astype(np.float64)
        am = _np(allMoments).astype(np.float64)
        ml = x.size / x.shape[1]
        n, mean, m2 = 0.0, np.zeros(x.shape[1]), np.zeros(x.shape[1])
        for r in range(replicas):   # Chan's merge in rank order
            d = am[r, :, 0] - mean
            tot = n + ml
            mean = mean + d * (ml / tot); m2 = m2 + am[r, :, 1] + d * d * (n * ml / tot); n = tot
        var = m2 / n
        inv = 1.0 / np.sqrt(var + epsilon)
        y = (x - mean[None, :, None, None]) * inv[None, :, None, None] * _np(gamma)[None, :, None, None] + _np(beta)[None, :, None, None]
        output.copy_(torch.from_numpy(y.astype(np.float32)))
        saveMean.copy_(torch.from_numpy(mean.astype(np.float32))); saveInvVariance.copy_(torch.from_numpy(inv.astype(np.float32)))
        runningMean.copy_(torch.from_numpy(((1 - momentum) * _np(runningMean) + momentum * mean).astype(np.float32)))
        runningVar.copy_(torch.from_numpy(((1 - momentum) * _np(runningVar) + momentum * var * (n / (n - 1))).astype(np.float32)))

    def BatchNormalizationGradientSums(self, input, mode, outputGradient, savedMean, sums):
        x, dy = _np(input).astype(np.float64), _np(outputGradient).astype(np.float64)
        xmu = x - _np(savedMean).astype(np.float64)[None, :, None, None]
        s = np.stack([dy.sum(axis=(0, 2, 3)), (dy * xmu).sum(axis=(0, 2, 3)), xmu.sum(axis=(0, 2, 3))], axis=1)
        sums.copy_(torch.from_numpy(s.astype(np.float32)))

    def BatchNormalizationGradientFromSums(self, replicas, globalSums, localSums, input, mode, gamma, outputGradient, savedMean, savedInvVariance,
                                           gammaGradient, betaGradient, inputGradient):
        x, dy = _np(input).astype(np.float64), _np(outputGradient).astype(np.float64)
        g, mu, inv = (_np(t).astype(np.float64) for t in (gamma, savedMean, savedInvVariance))
        gs, ls = _np(globalSums).astype(np.float64), _np(localSums).astype(np.float64)
        m = replicas * x.size / x.shape[1]
        dvar = g * gs[:, 1] * -0.5 * inv ** 3
        dmu = -inv * g * gs[:, 0] + dvar * (-2.0 * gs[:, 2] / m)
        b = lambda a: a[None, :, None, None]
        dx = dy * b(g) * b(inv) + b(dvar) * (x - b(mu)) * 2.0 / m + b(dmu) / m
        inputGradient.copy_(torch.from_numpy(dx.astype(np.float32)))
        if gammaGradient is not None:
            gammaGradient.copy_(torch.from_numpy((inv * ls[:, 1]).astype(np.float32)))
        if betaGradient is not None:
            betaGradient.copy_(torch.from_numpy(ls[:, 0].astype(np.float32)))
