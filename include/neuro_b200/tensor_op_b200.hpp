// neuro_b200/tensor_op_b200.hpp -- TensorOpB200: the B200 backend behind the reference's conv op interface.
//
// Method for method the convolution slice of Neuro::TensorOpCpu (Neuro/include/Tensors/TensorOpCpu.h:46-50,75-76),
// following the preamble every reference GPU op uses (TensorOpGpu.cpp:629-845): inputs CopyToDevice(), outputs
// OverrideDevice(), then the library call on device pointers -- here the C ABI of include/neuro_b200.h instead of
// cuDNN. No host synchronisation after the op (the reference calls cudaStreamSynchronize(0) after each, e.g. :667);
// a later CopyToHost() synchronises when host values are actually needed.
#pragma once

#include "tensor.hpp"

namespace NeuroB200
{
    class TensorOpB200 : public TensorOp
    {
    public:
        explicit TensorOpB200(int math = NB200_MATH_TF32, cudaStream_t stream = nullptr) : m_Math(math), m_Stream(stream) {}
        ~TensorOpB200() override { if (m_Workspace) cudaFree(m_Workspace); }

        EOpMode OpMode() const override { return B200; }
        bool IsDeviceBackend() const override { return true; }
        void SetMath(int math) { m_Math = math; }

        void Conv2D(const Tensor& input, const Tensor& kernels, uint32_t stride, uint32_t paddingX, uint32_t paddingY, EDataFormat dataFormat, Tensor& output) const override
        {
            Forward(input, kernels, stride, paddingX, paddingY, dataFormat, nullptr, _Identity, 0.f, output);
        }

        void Conv2DBiasActivation(const Tensor& input, const Tensor& kernels, uint32_t stride, uint32_t paddingX, uint32_t paddingY, const Tensor& bias, EActivation activation, float activationAlpha, Tensor& output) override
        {
            if (paddingX != paddingY) throw std::runtime_error("Conv2DBiasActivation: paddingX != paddingY"); // TensorOpCpu.cpp:1057
            Forward(input, kernels, stride, paddingX, paddingY, NCHW, &bias, activation, activationAlpha, output);
        }

        void Conv2DBiasGradient(const Tensor& gradient, Tensor& biasGradient) override
        {
            gradient.CopyToDevice();
            biasGradient.OverrideDevice();
            nb200_conv_desc d{};
            d.N = gradient.Batch(); d.K = gradient.Depth(); d.Ho = gradient.Height(); d.Wo = gradient.Width(); d.R = d.S = 1; d.stride = 1;
            d.fmt = NB200_NCHW; d.math = m_Math;
            Nb200Check(nb200_conv2d_bias_gradient(&d, gradient.GetDevicePtr(), biasGradient.GetDevicePtr(), m_Stream));
        }

        void ActivationGradient(EActivation activation, float alpha, const Tensor& output, const Tensor& outputGradient, Tensor& inputGradient) const override
        {
            BiasActivationGradient(output, outputGradient, activation, alpha, inputGradient, nullptr);
        }

        // one pass over HBM instead of the reference's two (ActivationGradient, then Conv2DBiasGradient re-reading its result)
        void Conv2DBiasActivationGradient(const Tensor& output, const Tensor& outputGradient, EActivation activation, float alpha, Tensor& activationInputGradient, Tensor& biasGradient) override
        {
            BiasActivationGradient(output, outputGradient, activation, alpha, activationInputGradient, &biasGradient);
        }

        void Conv2DInputGradient(const Tensor& gradient, const Tensor& kernels, uint32_t stride, uint32_t paddingX, uint32_t paddingY, EDataFormat dataFormat, Tensor& inputGradient) const override
        {
            gradient.CopyToDevice(); kernels.CopyToDevice(); inputGradient.OverrideDevice();
            const nb200_conv_desc d = Describe(inputGradient, kernels, gradient, stride, paddingX, paddingY, dataFormat);
            size_t ws = 0; void* w = Workspace(NB200_OP_INPUT_GRADIENT, d, ws);
            Nb200Check(nb200_conv2d_input_gradient(&d, gradient.GetDevicePtr(), kernels.GetDevicePtr(), inputGradient.GetDevicePtr(), w, ws, m_Stream));
        }

        void Conv2DKernelsGradient(const Tensor& input, const Tensor& gradient, uint32_t stride, uint32_t paddingX, uint32_t paddingY, EDataFormat dataFormat, Tensor& kernelsGradient) const override
        {
            input.CopyToDevice(); gradient.CopyToDevice(); kernelsGradient.OverrideDevice();
            const nb200_conv_desc d = Describe(input, kernelsGradient, gradient, stride, paddingX, paddingY, dataFormat);
            size_t ws = 0; void* w = Workspace(NB200_OP_KERNELS_GRADIENT, d, ws);
            Nb200Check(nb200_conv2d_kernels_gradient(&d, input.GetDevicePtr(), gradient.GetDevicePtr(), kernelsGradient.GetDevicePtr(), nullptr, w, ws, m_Stream));
        }

        void Pool2D(const Tensor& input, uint32_t filterSize, uint32_t stride, EPoolingMode type, uint32_t paddingX, uint32_t paddingY, EDataFormat dataFormat, Tensor& output) const override
        {
            input.CopyToDevice(); output.OverrideDevice();
            const nb200_pool_desc d = DescribePool(input, output, filterSize, stride, type, paddingX, paddingY, dataFormat);
            Nb200Check(nb200_pool2d(&d, input.GetDevicePtr(), output.GetDevicePtr(), m_Stream));
        }

        void Pool2DGradient(const Tensor& output, const Tensor& input, const Tensor& outputGradient, uint32_t filterSize, uint32_t stride, EPoolingMode type, uint32_t paddingX, uint32_t paddingY, EDataFormat dataFormat, Tensor& inputGradient) const override
        {
            output.CopyToDevice(); input.CopyToDevice(); outputGradient.CopyToDevice(); inputGradient.OverrideDevice();
            const nb200_pool_desc d = DescribePool(input, output, filterSize, stride, type, paddingX, paddingY, dataFormat);
            Nb200Check(nb200_pool2d_gradient(&d, output.GetDevicePtr(), input.GetDevicePtr(), outputGradient.GetDevicePtr(), inputGradient.GetDevicePtr(), m_Stream));
        }

        void UpSample2D(const Tensor& input, uint32_t scaleFactor, Tensor& output) const override
        {
            input.CopyToDevice(); output.OverrideDevice();
            Nb200Check(nb200_upsample2d(input.Batch(), input.Depth(), input.Height(), input.Width(), scaleFactor, input.GetDevicePtr(), output.GetDevicePtr(), m_Stream));
        }

        void UpSample2DGradient(const Tensor& outputGradient, uint32_t scaleFactor, Tensor& inputGradient) const override
        {
            outputGradient.CopyToDevice(); inputGradient.OverrideDevice();
            Nb200Check(nb200_upsample2d_gradient(inputGradient.Batch(), inputGradient.Depth(), inputGradient.Height(), inputGradient.Width(), scaleFactor,
                                                 outputGradient.GetDevicePtr(), inputGradient.GetDevicePtr(), m_Stream));
        }

        void ConstantPad2D(const Tensor& input, uint32_t left, uint32_t right, uint32_t top, uint32_t bottom, float value, Tensor& output) const override
        {
            input.CopyToDevice(); output.OverrideDevice();
            Nb200Check(nb200_constant_pad2d(input.Batch(), input.Depth(), input.Height(), input.Width(), left, right, top, bottom, value, input.GetDevicePtr(),
                                            output.GetDevicePtr(), m_Stream));
        }

        // TensorOpCpu.h:55-57 -> nb200_batch_norm*; statistics / gamma / beta are flat G-element tensors (Spatial: G = depth)
        void BatchNormalization(const Tensor& input, EBatchNormMode mode, const Tensor& gamma, const Tensor& beta, float epsilon, const Tensor* runningMean, const Tensor* runningVar, Tensor& output) const override
        {
            if (!runningMean || !runningVar) throw std::runtime_error("BatchNormalization: running statistics required");
            input.CopyToDevice(); gamma.CopyToDevice(); beta.CopyToDevice(); runningMean->CopyToDevice(); runningVar->CopyToDevice(); output.OverrideDevice();
            const nb200_bn_desc d = DescribeBn(input, mode);
            Nb200Check(nb200_batch_norm(&d, input.GetDevicePtr(), gamma.GetDevicePtr(), beta.GetDevicePtr(), epsilon, runningMean->GetDevicePtr(), runningVar->GetDevicePtr(), output.GetDevicePtr(), m_Stream));
        }

        void BatchNormalizationTrain(const Tensor& input, EBatchNormMode mode, const Tensor& gamma, const Tensor& beta, float momentum, float epsilon, Tensor* runningMean, Tensor* runningVar, Tensor& saveMean, Tensor& saveInvVariance, Tensor& output) const override
        {
            input.CopyToDevice(); gamma.CopyToDevice(); beta.CopyToDevice(); output.OverrideDevice(); saveMean.OverrideDevice(); saveInvVariance.OverrideDevice();
            if (runningMean) runningMean->CopyToDevice();
            if (runningVar) runningVar->CopyToDevice();
            const nb200_bn_desc d = DescribeBn(input, mode);
            const size_t bytes = nb200_batch_norm_workspace_bytes(&d);
            Nb200Check(nb200_batch_norm_train(&d, input.GetDevicePtr(), gamma.GetDevicePtr(), beta.GetDevicePtr(), momentum, epsilon, runningMean ? runningMean->GetDevicePtr() : nullptr,
                                              runningVar ? runningVar->GetDevicePtr() : nullptr, saveMean.GetDevicePtr(), saveInvVariance.GetDevicePtr(), output.GetDevicePtr(),
                                              bytes ? Grow(bytes) : nullptr, bytes, m_Stream));
        }

        void BatchNormalizationGradient(const Tensor& input, EBatchNormMode mode, const Tensor& gamma, float /*epsilon*/, const Tensor& outputGradient, const Tensor& savedMean, const Tensor& savedInvVariance, Tensor& gammaGradient, Tensor& betaGradient, bool /*trainable*/, Tensor& inputGradient) const override
        {
            input.CopyToDevice(); gamma.CopyToDevice(); outputGradient.CopyToDevice(); savedMean.CopyToDevice(); savedInvVariance.CopyToDevice();
            gammaGradient.OverrideDevice(); betaGradient.OverrideDevice(); inputGradient.OverrideDevice();
            const nb200_bn_desc d = DescribeBn(input, mode);
            const size_t bytes = nb200_batch_norm_workspace_bytes(&d);
            Nb200Check(nb200_batch_norm_gradient(&d, input.GetDevicePtr(), gamma.GetDevicePtr(), outputGradient.GetDevicePtr(), savedMean.GetDevicePtr(), savedInvVariance.GetDevicePtr(),
                                                 gammaGradient.GetDevicePtr(), betaGradient.GetDevicePtr(), inputGradient.GetDevicePtr(), bytes ? Grow(bytes) : nullptr, bytes, m_Stream));
        }

        // backward of "fused conv layer -> 2x2 max pooling" in one pass (include/neuro_b200.h: nb200_pool2d_gradient_activation); false = geometry not covered, caller issues the two ops
        bool Pool2DGradientActivation(const Tensor& output, const Tensor& input, const Tensor& outputGradient, uint32_t filterSize, uint32_t stride, EPoolingMode type, uint32_t padding, EActivation activation, float alpha, Tensor& activationInputGradient, Tensor& biasGradient) const
        {
            const nb200_pool_desc d = DescribePool(input, output, filterSize, stride, type, padding, padding, NCHW);
            if (!nb200_pool2d_gradient_activation_supported(&d)) return false;
            output.CopyToDevice(); input.CopyToDevice(); outputGradient.CopyToDevice(); activationInputGradient.OverrideDevice(); biasGradient.OverrideDevice();
            const size_t bytes = nb200_pool2d_gradient_activation_workspace_bytes(&d);
            Nb200Check(nb200_pool2d_gradient_activation(&d, (int)activation, alpha, output.GetDevicePtr(), input.GetDevicePtr(), outputGradient.GetDevicePtr(), activationInputGradient.GetDevicePtr(),
                                                        biasGradient.GetDevicePtr(), bytes ? Grow(bytes) : nullptr, bytes, m_Stream));
            return true;
        }

        void AdamStep(Tensor& parameter, const Tensor& gradient, Tensor& mGrad, Tensor& vGrad, float lr, float beta1, float beta2, float epsilon) const override
        {
            parameter.CopyToDevice(); gradient.CopyToDevice(); mGrad.CopyToDevice(); vGrad.CopyToDevice();
            Nb200Check(nb200_adam_step(parameter.GetDevicePtr(), gradient.GetDevicePtr(), mGrad.GetDevicePtr(), vGrad.GetDevicePtr(), parameter.Length(), 1.f, lr, beta1, beta2, epsilon, m_Stream));
        }

        void SgdStep(Tensor& parameter, const Tensor& gradient, float lr) const override
        {
            parameter.CopyToDevice(); gradient.CopyToDevice();
            Nb200Check(nb200_sgd_step(parameter.GetDevicePtr(), gradient.GetDevicePtr(), parameter.Length(), 1.f, lr, m_Stream));
        }

    private:
        // (N,C,H,W)/(Ho,Wo)/(K,R,S) from the reference Shapes: NCHW tensors are Shape(W,H,C,N), NHWC Shape(C,W,H,N), kernels Shape(S,R,C,K)
        nb200_conv_desc Describe(const Tensor& x, const Tensor& kernels, const Tensor& y, uint32_t stride, uint32_t paddingX, uint32_t paddingY, EDataFormat fmt) const
        {
            nb200_conv_desc d{};
            if (fmt == NCHW) { d.W = x.Len(0); d.H = x.Len(1); d.C = x.Len(2); d.N = x.Len(3); d.Wo = y.Len(0); d.Ho = y.Len(1); }
            else { d.C = x.Len(0); d.W = x.Len(1); d.H = x.Len(2); d.N = x.Len(3); d.Wo = y.Len(1); d.Ho = y.Len(2); }
            d.S = kernels.Len(0); d.R = kernels.Len(1); d.K = kernels.Len(3);
            if ((uint32_t)d.C != kernels.Len(2)) throw std::runtime_error("kernel depth does not match input depth");
            d.stride = stride; d.padX = paddingX; d.padY = paddingY; d.fmt = fmt; d.math = m_Math;
            return d;
        }
        void Forward(const Tensor& input, const Tensor& kernels, uint32_t stride, uint32_t paddingX, uint32_t paddingY, EDataFormat fmt, const Tensor* bias, EActivation act, float alpha, Tensor& output) const
        {
            input.CopyToDevice(); kernels.CopyToDevice(); if (bias) bias->CopyToDevice(); output.OverrideDevice();
            const nb200_conv_desc d = Describe(input, kernels, output, stride, paddingX, paddingY, fmt);
            size_t ws = 0; void* w = Workspace(NB200_OP_FORWARD, d, ws);
            Nb200Check(nb200_conv2d_forward(&d, input.GetDevicePtr(), kernels.GetDevicePtr(), bias ? bias->GetDevicePtr() : nullptr, (int)act, alpha, output.GetDevicePtr(), w, ws, m_Stream));
        }
        static nb200_bn_desc DescribeBn(const Tensor& x, EBatchNormMode mode)
        {
            nb200_bn_desc d{};
            d.N = x.Batch(); d.C = x.Depth(); d.H = x.Height(); d.W = x.Width(); d.mode = (int)mode;
            return d;
        }
        static nb200_pool_desc DescribePool(const Tensor& x, const Tensor& y, uint32_t filterSize, uint32_t stride, EPoolingMode type, uint32_t paddingX, uint32_t paddingY, EDataFormat fmt)
        {
            nb200_pool_desc d{};
            if (fmt == NCHW) { d.W = x.Len(0); d.H = x.Len(1); d.C = x.Len(2); d.Wo = y.Len(0); d.Ho = y.Len(1); }
            else { d.C = x.Len(0); d.W = x.Len(1); d.H = x.Len(2); d.Wo = y.Len(1); d.Ho = y.Len(2); }
            d.N = x.Len(3); d.filter = filterSize; d.stride = stride; d.padX = paddingX; d.padY = paddingY; d.mode = (int)type; d.fmt = fmt;
            return d;
        }
        void BiasActivationGradient(const Tensor& output, const Tensor& outputGradient, EActivation act, float alpha, Tensor& dz, Tensor* db) const
        {
            output.CopyToDevice(); outputGradient.CopyToDevice(); dz.OverrideDevice(); if (db) db->OverrideDevice();
            nb200_conv_desc d{};
            d.N = outputGradient.Batch(); d.K = outputGradient.Depth(); d.Ho = outputGradient.Height(); d.Wo = outputGradient.Width();
            d.R = d.S = 1; d.stride = 1; d.fmt = NB200_NCHW; d.math = m_Math;
            const size_t need = db ? nb200_conv2d_bias_activation_gradient_workspace_bytes(&d) : 0;
            void* w = need ? Grow(need) : nullptr;
            Nb200Check(nb200_conv2d_bias_activation_gradient(&d, (int)act, alpha, output.GetDevicePtr(), outputGradient.GetDevicePtr(), dz.GetDevicePtr(),
                                                             db ? db->GetDevicePtr() : nullptr, w, need, m_Stream));
        }
        // grow-only scratch (reference: pooled workspace per call, TensorOpGpu.cpp:648)
        void* Workspace(int op, const nb200_conv_desc& d, size_t& bytes) const
        {
            bytes = nb200_conv2d_workspace_bytes(op, &d);
            return bytes ? Grow(bytes) : nullptr;
        }
        void* Grow(size_t bytes) const
        {
            if (bytes > m_WorkspaceBytes)
            {
                if (m_Workspace) { CudaCheck(cudaDeviceSynchronize(), "workspace sync"); cudaFree(m_Workspace); }
                CudaCheck(cudaMalloc(&m_Workspace, bytes), "workspace");
                m_WorkspaceBytes = bytes;
            }
            return m_Workspace;
        }

        int m_Math;
        cudaStream_t m_Stream;
        mutable void* m_Workspace = nullptr;
        mutable size_t m_WorkspaceBytes = 0;
    };

    inline TensorOp* Tensor::GetOpFromMode(EOpMode mode)
    {
        TensorOp*& slot = OpTable()[mode];
        if (!slot && mode == B200)
            slot = new TensorOpB200(); // lazy singleton, like Tensor.cpp:2701-2716
        if (!slot)
            throw std::runtime_error("no backend registered for this EOpMode (register one with Tensor::RegisterOp)");
        return slot;
    }
}
