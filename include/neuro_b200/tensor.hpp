// neuro_b200/tensor.hpp -- C++ host layer above the C ABI (include/neuro_b200.h).
//
// A minimal, device-resident re-creation of the three reference types the convolution path touches, with the
// reference's names and semantics so code written against Neuro_ reads the same:
//
//   Shape    Neuro/include/Tensors/Shape.h:15-104, Neuro/src/Tensors/Shape.cpp:11-33   Dimensions[4] = {W,H,D,N}, W fastest
//   Tensor   Neuro/include/Tensors/Tensor.h (conv wrappers Tensor.cpp:1757-1830, shape helpers :1966-2051,
//            residency calls :2509-2634)
//   Storage  Neuro/src/Tensors/Storage.cpp:534-654: exactly one of {host, device} is authoritative; CopyToDevice /
//            CopyToHost move it, OverrideDevice / OverrideHost claim a side without copying ("I will overwrite").
//
// Differences, on purpose: storage is device-first (a tensor touched by a B200 op stays in HBM until somebody asks for
// host values; the reference mirrors every device buffer with a host allocation, Storage.cpp:246,269); "is this a
// device backend" is the virtual predicate TensorOp::IsDeviceBackend() instead of pointer equality with g_OpGpu
// (SURVEY.md section 8b lists the 11 reference sites that hard-code that comparison).
//
// Header-only; link with libneuro_b200.so and libcudart. Errors: CUDA / nb200 failures throw std::runtime_error
// (the reference asserts in debug and ignores in release; we always check).
#pragma once

#include <cuda_runtime_api.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../neuro_b200.h"

namespace NeuroB200
{
    enum EOpMode { CPU, CPU_MKL, CPU_MT, GPU, B200 };                       // Types.h:34-40 + the new backend
    enum ELocation { None, Host, Device };                                   // Types.h:42-47
    enum EPaddingMode { Valid, Same, Full };                                 // Types.h:49-54
    enum EDataFormat { NCHW = NB200_NCHW, NHWC = NB200_NHWC };               // Types.h:94-98
    enum EPoolingMode { MaxPool = NB200_POOL_MAX, AvgPool = NB200_POOL_AVG };  // Types.h:56-60
    enum EActivation { _Identity, _Sigmoid, _ReLU, _TanH, _ELU, _LeakyReLU, _Softmax }; // Types.h:83-92
    enum EBatchNormMode { PerActivation = NB200_BN_PER_ACTIVATION, Spatial = NB200_BN_SPATIAL, Instance = NB200_BN_INSTANCE }; // Types.h:70-75

    inline void CudaCheck(cudaError_t e, const char* what)
    {
        if (e != cudaSuccess)
            throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
    }

    inline void Nb200Check(int rc)
    {
        if (rc != NB200_OK)
            throw std::runtime_error(std::string("nb200: ") + nb200_last_error());
    }

    class Shape
    {
    public:
        explicit Shape(uint32_t width = 0, uint32_t height = 1, uint32_t depth = 1, uint32_t batch = 1)
        {
            Dimensions[0] = width; Dimensions[1] = height; Dimensions[2] = depth; Dimensions[3] = batch;
            Dim0 = width; Dim0Dim1 = Dim0 * height; Dim0Dim1Dim2 = Dim0Dim1 * depth; Length = Dim0Dim1Dim2 * batch;
        }
        uint32_t Width() const { return Dimensions[0]; }
        uint32_t Height() const { return Dimensions[1]; }
        uint32_t Depth() const { return Dimensions[2]; }
        uint32_t Batch() const { return Dimensions[3]; }
        uint32_t Len(size_t dim) const { return Dimensions[dim]; }
        uint32_t GetIndex(uint32_t w, uint32_t h = 0, uint32_t d = 0, uint32_t n = 0) const { return Dim0Dim1Dim2 * n + Dim0Dim1 * d + Dim0 * h + w; }
        bool operator==(const Shape& o) const { return std::memcmp(Dimensions, o.Dimensions, sizeof(Dimensions)) == 0; }
        bool operator!=(const Shape& o) const { return !(*this == o); }

        uint32_t Dimensions[4];
        uint32_t Dim0, Dim0Dim1, Dim0Dim1Dim2, Length;
    };

    class Tensor;

    // The slice of Neuro::TensorOpCpu this backend replaces (TensorOpCpu.h:46-50, 75-76): same names, argument order and
    // meaning. A CPU implementation can be plugged in by subclassing (tests do, with the oracle), mirroring
    // Tensor::GetOpFromMode (Tensor.cpp:2701-2716).
    class TensorOp
    {
    public:
        virtual ~TensorOp() {}
        virtual EOpMode OpMode() const = 0;
        virtual bool IsDeviceBackend() const = 0;
        virtual void Conv2D(const Tensor& input, const Tensor& kernels, uint32_t stride, uint32_t paddingX, uint32_t paddingY, EDataFormat dataFormat, Tensor& output) const = 0;
        virtual void Conv2DBiasActivation(const Tensor& input, const Tensor& kernels, uint32_t stride, uint32_t paddingX, uint32_t paddingY, const Tensor& bias, EActivation activation, float activationAlpha, Tensor& output) = 0;
        virtual void Conv2DBiasGradient(const Tensor& gradient, Tensor& biasGradient) = 0;
        virtual void Conv2DInputGradient(const Tensor& gradient, const Tensor& kernels, uint32_t stride, uint32_t paddingX, uint32_t paddingY, EDataFormat dataFormat, Tensor& inputGradient) const = 0;
        virtual void Conv2DKernelsGradient(const Tensor& input, const Tensor& gradient, uint32_t stride, uint32_t paddingX, uint32_t paddingY, EDataFormat dataFormat, Tensor& kernelsGradient) const = 0;
        // Tensor::ActivationGradient's targets (TensorOpCpu.h: SigmoidGradient ... LeakyReLUGradient; TensorOpCpu.cpp:813-864) behind one entry
        virtual void ActivationGradient(EActivation, float, const Tensor&, const Tensor&, Tensor&) const { throw std::runtime_error("ActivationGradient: not implemented by this backend"); }
        // the backward prologue of Conv2dBiasActivationOp (Conv2dBiasActivationOp.cpp:47-60); backends may fuse the two passes
        virtual void Conv2DBiasActivationGradient(const Tensor& output, const Tensor& outputGradient, EActivation activation, float activationAlpha, Tensor& activationInputGradient, Tensor& biasGradient)
        {
            ActivationGradient(activation, activationAlpha, output, outputGradient, activationInputGradient);
            Conv2DBiasGradient(activationInputGradient, biasGradient);
        }
        // resamplers around the convolutions (TensorOpCpu.h:51-54; ConstantPad2D TensorOpCpu.cpp:528)
        virtual void Pool2D(const Tensor&, uint32_t, uint32_t, EPoolingMode, uint32_t, uint32_t, EDataFormat, Tensor&) const { throw std::runtime_error("Pool2D: not implemented by this backend"); }
        virtual void Pool2DGradient(const Tensor&, const Tensor&, const Tensor&, uint32_t, uint32_t, EPoolingMode, uint32_t, uint32_t, EDataFormat, Tensor&) const { throw std::runtime_error("Pool2DGradient: not implemented by this backend"); }
        virtual void UpSample2D(const Tensor&, uint32_t, Tensor&) const { throw std::runtime_error("UpSample2D: not implemented by this backend"); }
        virtual void UpSample2DGradient(const Tensor&, uint32_t, Tensor&) const { throw std::runtime_error("UpSample2DGradient: not implemented by this backend"); }
        virtual void ConstantPad2D(const Tensor&, uint32_t, uint32_t, uint32_t, uint32_t, float, Tensor&) const { throw std::runtime_error("ConstantPad2D: not implemented by this backend"); }
        // batch normalisation (TensorOpCpu.h:55-57); gamma / beta / statistics hold one value per normalisation group
        virtual void BatchNormalization(const Tensor&, EBatchNormMode, const Tensor&, const Tensor&, float, const Tensor*, const Tensor*, Tensor&) const { throw std::runtime_error("BatchNormalization: not implemented by this backend"); }
        virtual void BatchNormalizationTrain(const Tensor&, EBatchNormMode, const Tensor&, const Tensor&, float, float, Tensor*, Tensor*, Tensor&, Tensor&, Tensor&) const { throw std::runtime_error("BatchNormalizationTrain: not implemented by this backend"); }
        virtual void BatchNormalizationGradient(const Tensor&, EBatchNormMode, const Tensor&, float, const Tensor&, const Tensor&, const Tensor&, Tensor&, Tensor&, bool, Tensor&) const { throw std::runtime_error("BatchNormalizationGradient: not implemented by this backend"); }
        virtual void AdamStep(Tensor& parameter, const Tensor& gradient, Tensor& mGrad, Tensor& vGrad, float lr, float beta1, float beta2, float epsilon) const = 0;
        virtual void SgdStep(Tensor& parameter, const Tensor& gradient, float lr) const = 0;
    };

    class Tensor
    {
    public:
        explicit Tensor(const Shape& shape = Shape(0), const std::string& name = "") : m_Shape(shape), m_Name(name) {}
        Tensor(const std::vector<float>& values, const Shape& shape, const std::string& name = "") : m_Shape(shape), m_Name(name)
        {
            if (values.size() != shape.Length) throw std::runtime_error("Tensor: values do not match shape");
            m_Host = values; m_Location = Host;
        }
        Tensor(const Tensor& t) : m_Shape(t.m_Shape), m_Name(t.m_Name) { t.CopyToHost(); m_Host = t.m_Host; m_Location = Host; } // deep copy to host, Storage.cpp:51-81
        Tensor& operator=(const Tensor& t)
        {
            if (this != &t) { ReleaseDevice(); t.CopyToHost(); m_Shape = t.m_Shape; m_Host = t.m_Host; m_Location = Host; m_Name = t.m_Name; }
            return *this;
        }
        ~Tensor() { ReleaseDevice(); }

        static void SetDefaultOpMode(EOpMode mode) { DefaultOpSlot() = GetOpFromMode(mode); }   // Tensor.cpp:158-161
        static void SetForcedOpMode(EOpMode mode) { ForcedOpSlot() = GetOpFromMode(mode); }     // Tensor.cpp:164-167
        static void ClearForcedOpMode() { ForcedOpSlot() = nullptr; }
        static void RegisterOp(EOpMode mode, TensorOp* op) { OpTable()[mode] = op; }
        static TensorOp* GetOpFromMode(EOpMode mode);

        uint32_t Width() const { return m_Shape.Width(); }
        uint32_t Height() const { return m_Shape.Height(); }
        uint32_t Depth() const { return m_Shape.Depth(); }
        uint32_t Batch() const { return m_Shape.Batch(); }
        uint32_t Len(size_t d) const { return m_Shape.Len(d); }
        uint32_t Length() const { return m_Shape.Length; }
        const Shape& GetShape() const { return m_Shape; }

        // ---- residency protocol (Storage.cpp:534-654) ----
        void CopyToDevice() const
        {
            if (m_Location == Device) return;
            AllocateOnDevice();
            if (m_Location == Host)
                CudaCheck(cudaMemcpy(m_Device, m_Host.data(), (size_t)Length() * sizeof(float), cudaMemcpyHostToDevice), "CopyToDevice");
            else
                CudaCheck(cudaMemset(m_Device, 0, (size_t)Length() * sizeof(float)), "CopyToDevice(zero)");
            m_Location = Device;
        }
        void CopyToHost() const
        {
            if (m_Location == Host) return;
            m_Host.resize(Length());
            if (m_Location == Device)
            {
                CudaCheck(cudaDeviceSynchronize(), "CopyToHost(sync)");
                CudaCheck(cudaMemcpy(m_Host.data(), m_Device, (size_t)Length() * sizeof(float), cudaMemcpyDeviceToHost), "CopyToHost");
            }
            else
                std::fill(m_Host.begin(), m_Host.end(), 0.f);
            m_Location = Host;
        }
        void OverrideDevice() { AllocateOnDevice(); m_Location = Device; }
        void OverrideHost() { m_Host.resize(Length()); m_Location = Host; }
        bool IsOnHost() const { return m_Location == Host; }
        bool IsOnDevice() const { return m_Location == Device; }
        const float* GetDevicePtr() const { if (m_Location != Device) throw std::runtime_error("GetDevicePtr: not on device"); return m_Device; }
        float* GetDevicePtr() { if (m_Location != Device) throw std::runtime_error("GetDevicePtr: not on device"); return m_Device; }
        float* Values() { CopyToHost(); return m_Host.data(); }
        const float* Values() const { CopyToHost(); return m_Host.data(); }

        // ---- element access / fills (host side) ----
        float GetFlat(uint32_t i) const { CopyToHost(); return m_Host[i]; }
        float Get(uint32_t w, uint32_t h = 0, uint32_t d = 0, uint32_t n = 0) const { CopyToHost(); return m_Host[m_Shape.GetIndex(w, h, d, n)]; }
        float operator()(uint32_t w, uint32_t h = 0, uint32_t d = 0, uint32_t n = 0) const { return Get(w, h, d, n); }
        Tensor& FillWithRange(float start = 0, float increment = 1)                                  // Tensor.cpp:261-267
        {
            OverrideHost();
            for (uint32_t i = 0; i < Length(); ++i) m_Host[i] = start + i * increment;
            return *this;
        }
        // counter-based U(min,max): same stream as neuro__b200/synth.py (the reference uses std::mt19937, Tensor.cpp:239-247)
        Tensor& FillWithRand(int seed = -1, float min = -1, float max = 1)
        {
            OverrideHost();
            const uint64_t key = SplitMix64((uint64_t)(seed < 0 ? 0 : seed) * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull);
            for (uint32_t i = 0; i < Length(); ++i)
            {
                const double u = (double)(SplitMix64((uint64_t)i ^ key) >> 40) * (1.0 / (1 << 24));
                m_Host[i] = (float)(min + (max - min) * u);
            }
            return *this;
        }
        void Zero() { OverrideHost(); std::fill(m_Host.begin(), m_Host.end(), 0.f); }
        bool Equals(const Tensor& other, float epsilon = 0.00001f) const                             // Tensor.cpp:2203-2220
        {
            if (m_Shape != other.m_Shape) return false;
            CopyToHost(); other.CopyToHost();
            for (uint32_t i = 0; i < Length(); ++i)
                if (std::fabs(m_Host[i] - other.m_Host[i]) > epsilon) return false;
            return true;
        }
        float MaxNormalisedError(const Tensor& reference) const
        {
            CopyToHost(); reference.CopyToHost();
            double err = 0, den = 0;
            for (uint32_t i = 0; i < Length(); ++i)
            {
                err = std::fmax(err, std::fabs((double)m_Host[i] - reference.m_Host[i]));
                den = std::fmax(den, std::fabs((double)reference.m_Host[i]));
            }
            return (float)(err / (den > 0 ? den : 1));
        }

        // ---- shape helpers (Tensor.cpp:1966-2051) ----
        static uint32_t GetPadding(EPaddingMode mode, uint32_t kernelSize) { return (uint32_t)nb200_padding((int)mode, (int)kernelSize); }
        static Shape GetConvOutputShape(const Shape& in, uint32_t kernelsNum, uint32_t kernelWidth, uint32_t kernelHeight, uint32_t stride, uint32_t paddingX, uint32_t paddingY, EDataFormat fmt)
        {
            if (fmt == NCHW)
                return Shape(nb200_conv_out_size(in.Width(), kernelWidth, stride, paddingX), nb200_conv_out_size(in.Height(), kernelHeight, stride, paddingY), kernelsNum, in.Batch());
            return Shape(kernelsNum, nb200_conv_out_size(in.Len(1), kernelWidth, stride, paddingX), nb200_conv_out_size(in.Len(2), kernelHeight, stride, paddingY), in.Len(3));
        }
        static Shape GetConvTransposeOutputShape(const Shape& in, uint32_t outputDepth, uint32_t kernelWidth, uint32_t kernelHeight, uint32_t stride, uint32_t paddingX, uint32_t paddingY, EDataFormat fmt)
        {
            if (fmt == NCHW)
                return Shape(nb200_conv_transpose_out_size(in.Width(), kernelWidth, stride, paddingX), nb200_conv_transpose_out_size(in.Height(), kernelHeight, stride, paddingY), outputDepth, in.Batch());
            return Shape(outputDepth, nb200_conv_transpose_out_size(in.Len(1), kernelWidth, stride, paddingX), nb200_conv_transpose_out_size(in.Len(2), kernelHeight, stride, paddingY), in.Len(3));
        }

        // ---- convolution wrappers (Tensor.cpp:1757-1830): shape checks + Op()->... dispatch ----
        void Conv2D(const Tensor& kernels, uint32_t stride, uint32_t padding, EDataFormat fmt, Tensor& output) const
        {
            if (GetConvOutputShape(m_Shape, kernels.Batch(), kernels.Width(), kernels.Height(), stride, padding, padding, fmt) != output.GetShape())
                throw std::runtime_error("Output shape doesn't match input shape.");
            Op()->Conv2D(*this, kernels, stride, padding, padding, fmt, output);
        }
        Tensor Conv2D(const Tensor& kernels, uint32_t stride, uint32_t padding, EDataFormat fmt) const
        {
            Tensor output(GetConvOutputShape(m_Shape, kernels.Batch(), kernels.Width(), kernels.Height(), stride, padding, padding, fmt));
            Conv2D(kernels, stride, padding, fmt, output);
            return output;
        }
        void Conv2DBiasActivation(const Tensor& kernels, uint32_t stride, uint32_t padding, const Tensor& bias, EActivation activation, float alpha, Tensor& output) const
        {
            Op()->Conv2DBiasActivation(*this, kernels, stride, padding, padding, bias, activation, alpha, output);
        }
        Tensor Conv2DBiasActivation(const Tensor& kernels, uint32_t stride, uint32_t padding, const Tensor& bias, EActivation activation, float alpha) const
        {
            Tensor output(GetConvOutputShape(m_Shape, kernels.Batch(), kernels.Width(), kernels.Height(), stride, padding, padding, NCHW));
            Conv2DBiasActivation(kernels, stride, padding, bias, activation, alpha, output);
            return output;
        }
        void Conv2DBiasGradient(const Tensor& gradient, Tensor& biasGradient) const { Op()->Conv2DBiasGradient(gradient, biasGradient); }
        // Tensor::ActivationGradient as Conv2dBiasActivationOp calls it: grad.ActivationGradient(act, alpha, output, grad, inputGrad)
        void ActivationGradient(EActivation activation, float alpha, const Tensor& output, const Tensor& outputGradient, Tensor& inputGradient) const
        {
            if (output.GetShape() != outputGradient.GetShape() || output.GetShape() != inputGradient.GetShape()) throw std::runtime_error("ActivationGradient: shapes differ");
            Op()->ActivationGradient(activation, alpha, output, outputGradient, inputGradient);
        }
        void Conv2DBiasActivationGradient(const Tensor& output, const Tensor& outputGradient, EActivation activation, float alpha, Tensor& activationInputGradient, Tensor& biasGradient) const
        {
            Op()->Conv2DBiasActivationGradient(output, outputGradient, activation, alpha, activationInputGradient, biasGradient);
        }
        void Conv2DInputsGradient(const Tensor& gradient, const Tensor& kernels, uint32_t stride, uint32_t padding, EDataFormat fmt, Tensor& inputsGradient) const
        {
            Op()->Conv2DInputGradient(gradient, kernels, stride, padding, padding, fmt, inputsGradient);
        }
        void Conv2DKernelsGradient(const Tensor& input, const Tensor& gradient, uint32_t stride, uint32_t padding, EDataFormat fmt, Tensor& kernelsGradient) const
        {
            Op()->Conv2DKernelsGradient(input, gradient, stride, padding, padding, fmt, kernelsGradient);
        }
        // ---- batch normalisation wrappers (Tensor.cpp: BatchNormalization / BatchNormalizationTrain / BatchNormalizationGradient) ----
        void BatchNormalization(const Tensor& gamma, const Tensor& beta, float epsilon, const Tensor* runningMean, const Tensor* runningVar, Tensor& result, EBatchNormMode mode = Spatial) const
        {
            Op()->BatchNormalization(*this, mode, gamma, beta, epsilon, runningMean, runningVar, result);
        }
        void BatchNormalizationTrain(const Tensor& gamma, const Tensor& beta, float momentum, float epsilon, Tensor* runningMean, Tensor* runningVar, Tensor& saveMean, Tensor& saveInvVariance, Tensor& result, EBatchNormMode mode = Spatial) const
        {
            Op()->BatchNormalizationTrain(*this, mode, gamma, beta, momentum, epsilon, runningMean, runningVar, saveMean, saveInvVariance, result);
        }
        void BatchNormalizationGradient(const Tensor& input, const Tensor& gamma, float epsilon, const Tensor& outputGradient, const Tensor& savedMean, const Tensor& savedInvVariance, Tensor& gammaGradient, Tensor& betaGradient, bool trainable, Tensor& inputGradient, EBatchNormMode mode = Spatial) const
        {
            Op()->BatchNormalizationGradient(input, mode, gamma, epsilon, outputGradient, savedMean, savedInvVariance, gammaGradient, betaGradient, trainable, inputGradient);
        }
        // ---- resampler wrappers (Tensor.cpp:1833-1876, 1500-1512): shape checks + dispatch ----
        static Shape GetPooling2DOutputShape(const Shape& in, uint32_t kernelWidth, uint32_t kernelHeight, uint32_t stride, uint32_t paddingX, uint32_t paddingY, EDataFormat fmt)
        {
            if (fmt == NCHW)
                return Shape((in.Width() + 2 * paddingX - kernelWidth) / stride + 1, (in.Height() + 2 * paddingY - kernelHeight) / stride + 1, in.Depth(), in.Batch());
            return Shape(in.Len(0), (in.Len(1) + 2 * paddingX - kernelWidth) / stride + 1, (in.Len(2) + 2 * paddingY - kernelHeight) / stride + 1, in.Len(3));
        }
        void Pool2D(uint32_t filterSize, uint32_t stride, EPoolingMode type, uint32_t padding, EDataFormat fmt, Tensor& output) const
        {
            if (GetPooling2DOutputShape(m_Shape, filterSize, filterSize, stride, padding, padding, fmt) != output.GetShape())
                throw std::runtime_error("Output shape doesn't match input shape.");
            Op()->Pool2D(*this, filterSize, stride, type, padding, padding, fmt, output);
        }
        Tensor Pool2D(uint32_t filterSize, uint32_t stride, EPoolingMode type, uint32_t padding, EDataFormat fmt) const
        {
            Tensor result(GetPooling2DOutputShape(m_Shape, filterSize, filterSize, stride, padding, padding, fmt));
            Pool2D(filterSize, stride, type, padding, fmt, result);
            return result;
        }
        void Pool2DGradient(const Tensor& output, const Tensor& input, const Tensor& outputGradient, uint32_t filterSize, uint32_t stride, EPoolingMode type, uint32_t padding, EDataFormat fmt, Tensor& result) const
        {
            Op()->Pool2DGradient(output, input, outputGradient, filterSize, stride, type, padding, padding, fmt, result);
        }
        void UpSample2D(uint32_t scaleFactor, Tensor& output) const
        {
            if (Shape(Width() * scaleFactor, Height() * scaleFactor, Depth(), Batch()) != output.GetShape())
                throw std::runtime_error("Output shape doesn't match input shape.");
            Op()->UpSample2D(*this, scaleFactor, output);
        }
        Tensor UpSample2D(uint32_t scaleFactor) const
        {
            Tensor result(Shape(Width() * scaleFactor, Height() * scaleFactor, Depth(), Batch()));
            UpSample2D(scaleFactor, result);
            return result;
        }
        void UpSample2DGradient(const Tensor& outputGradient, uint32_t scaleFactor, Tensor& inputGradient) const
        {
            if (Shape(inputGradient.Width() * scaleFactor, inputGradient.Height() * scaleFactor, inputGradient.Depth(), inputGradient.Batch()) != outputGradient.GetShape())
                throw std::runtime_error("Input gradient shape doesn't match input shape.");
            Op()->UpSample2DGradient(outputGradient, scaleFactor, inputGradient);
        }
        Tensor ConstantPad2D(uint32_t left, uint32_t right, uint32_t top, uint32_t bottom, float value) const
        {
            Tensor output(Shape(Width() + left + right, Height() + top + bottom, Depth(), Batch()));
            Op()->ConstantPad2D(*this, left, right, top, bottom, value, output);
            return output;
        }
        // transposed convolution identities (Tensor.cpp:1806-1830)
        void Conv2DTransposed(const Tensor& kernels, uint32_t stride, uint32_t padding, EDataFormat fmt, Tensor& result) const
        {
            Conv2DInputsGradient(*this, kernels, stride, padding, fmt, result);
        }
        Tensor Conv2DTransposed(const Tensor& kernels, uint32_t outputDepth, uint32_t stride, uint32_t padding, EDataFormat fmt) const
        {
            Tensor result(GetConvTransposeOutputShape(m_Shape, outputDepth, kernels.Width(), kernels.Height(), stride, padding, padding, fmt));
            Conv2DTransposed(kernels, stride, padding, fmt, result);
            return result;
        }
        void Conv2DTransposedInputsGradient(const Tensor& gradient, const Tensor& kernels, uint32_t stride, uint32_t padding, EDataFormat fmt, Tensor& inputsGradient) const
        {
            gradient.Conv2D(kernels, stride, padding, fmt, inputsGradient);
        }
        void Conv2DTransposedKernelsGradient(const Tensor& input, const Tensor& gradient, uint32_t stride, uint32_t padding, EDataFormat fmt, Tensor& kernelsGradient) const
        {
            Op()->Conv2DKernelsGradient(gradient, input, stride, padding, padding, fmt, kernelsGradient); // swapped, as in the reference
        }

        static TensorOp* ActiveOp() { return ForcedOpSlot() ? ForcedOpSlot() : DefaultOpSlot(); }

    private:
        TensorOp* Op() const { TensorOp* op = ActiveOp(); if (!op) throw std::runtime_error("no TensorOp registered for the active mode"); return op; }
        static TensorOp*& DefaultOpSlot() { static TensorOp* op = nullptr; return op; }
        static TensorOp*& ForcedOpSlot() { static TensorOp* op = nullptr; return op; }
        static TensorOp** OpTable() { static TensorOp* table[8] = {}; return table; }
        static uint64_t SplitMix64(uint64_t z)
        {
            z += 0x9E3779B97F4A7C15ull;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            return z ^ (z >> 31);
        }
        void AllocateOnDevice() const
        {
            if (!m_Device)
                CudaCheck(cudaMalloc((void**)&m_Device, (size_t)(Length() ? Length() : 1) * sizeof(float)), "cudaMalloc");
        }
        void ReleaseDevice() { if (m_Device) { cudaFree(m_Device); m_Device = nullptr; } if (m_Location == Device) m_Location = None; }

        Shape m_Shape;
        std::string m_Name;
        mutable std::vector<float> m_Host;
        mutable float* m_Device = nullptr;
        mutable ELocation m_Location = None;
    };
}
