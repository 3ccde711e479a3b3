// C++ parity tests for the host layer (include/neuro_b200/*.hpp), written the way the reference writes its own:
//   Neuro.Tests/src/TensorTests.cpp:352-425       known-answer Conv2D vectors
//   Neuro.Tests/src/TensorOpGpuTests.cpp:1196-1322  "SetForcedOpMode(CPU) ... SetForcedOpMode(GPU) ... r.Equals(r2)"
// The CPU side is the oracle (oracle/conv_oracle.c) registered as EOpMode::CPU; the device side is TensorOpB200.
// Test infrastructure: links liboracle; built and run by tests/test_cpp_host.py (-m gpu).
#include <cstdio>
#include <cstdlib>
#include <string>

#include "neuro_b200/tensor_op_b200.hpp"

using namespace NeuroB200;

extern "C"
{
    struct conv_dims { int N, C, H, W, K, R, S, Ho, Wo, stride, padX, padY, fmt; };
    void oracle_conv2d(const conv_dims*, const float*, const float*, float*);
    void oracle_conv2d_input_gradient(const conv_dims*, const float*, const float*, float*);
    void oracle_conv2d_kernels_gradient(const conv_dims*, const float*, const float*, float*);
    void oracle_conv2d_bias_activation(const conv_dims*, const float*, const float*, const float*, int, float, float*);
    void oracle_conv2d_bias_gradient(const conv_dims*, const float*, float*);
    void oracle_activation_gradient(int, float, const float*, const float*, float*, size_t);
    struct pool_dims { int N, C, H, W, Ho, Wo, filter, stride, padX, padY, mode, fmt; };
    void oracle_pool2d(const pool_dims*, const float*, float*);
    void oracle_pool2d_gradient(const pool_dims*, const float*, const float*, const float*, float*);
    void oracle_upsample2d(int, int, int, int, const float*, float*);
    void oracle_upsample2d_gradient(int, int, int, int, const float*, float*);
    void oracle_constant_pad2d(int, int, int, int, int, int, int, float, const float*, float*);
    void oracle_batch_norm_train(int, int, int, const float*, const float*, const float*, float, float, float*, float*, float*, float*, float*);
    void oracle_batch_norm(int, int, int, const float*, const float*, const float*, float, const float*, const float*, float*);
    void oracle_batch_norm_gradient(int, int, int, const float*, const float*, const float*, const float*, const float*, float*, float*, float*);
    void oracle_adam_step(float*, const float*, float*, float*, size_t, float, float, float, float);
    void oracle_sgd_step(float*, const float*, size_t, float);
}

// EOpMode::CPU for the tests = the oracle behind the same virtual interface
class TensorOpOracle : public TensorOp
{
    static conv_dims Dims(const Tensor& x, const Tensor& k, const Tensor& y, uint32_t stride, uint32_t px, uint32_t py, EDataFormat fmt)
    {
        conv_dims d{};
        if (fmt == NCHW) { d.W = x.Len(0); d.H = x.Len(1); d.C = x.Len(2); d.N = x.Len(3); d.Wo = y.Len(0); d.Ho = y.Len(1); }
        else { d.C = x.Len(0); d.W = x.Len(1); d.H = x.Len(2); d.N = x.Len(3); d.Wo = y.Len(1); d.Ho = y.Len(2); }
        d.S = k.Len(0); d.R = k.Len(1); d.K = k.Len(3); d.stride = stride; d.padX = px; d.padY = py; d.fmt = fmt;
        return d;
    }
public:
    EOpMode OpMode() const override { return CPU; }
    bool IsDeviceBackend() const override { return false; }
    void Conv2D(const Tensor& x, const Tensor& k, uint32_t s, uint32_t px, uint32_t py, EDataFormat f, Tensor& y) const override
    { const conv_dims d = Dims(x, k, y, s, px, py, f); y.OverrideHost(); oracle_conv2d(&d, x.Values(), k.Values(), y.Values()); }
    void Conv2DBiasActivation(const Tensor& x, const Tensor& k, uint32_t s, uint32_t px, uint32_t py, const Tensor& b, EActivation a, float alpha, Tensor& y) override
    { const conv_dims d = Dims(x, k, y, s, px, py, NCHW); y.OverrideHost(); oracle_conv2d_bias_activation(&d, x.Values(), k.Values(), b.Values(), (int)a, alpha, y.Values()); }
    void Conv2DBiasGradient(const Tensor& g, Tensor& db) override
    { conv_dims d{}; d.N = g.Batch(); d.K = g.Depth(); d.Ho = g.Height(); d.Wo = g.Width(); db.OverrideHost(); oracle_conv2d_bias_gradient(&d, g.Values(), db.Values()); }
    void Conv2DInputGradient(const Tensor& g, const Tensor& k, uint32_t s, uint32_t px, uint32_t py, EDataFormat f, Tensor& dx) const override
    { const conv_dims d = Dims(dx, k, g, s, px, py, f); dx.OverrideHost(); oracle_conv2d_input_gradient(&d, g.Values(), k.Values(), dx.Values()); }
    void Conv2DKernelsGradient(const Tensor& x, const Tensor& g, uint32_t s, uint32_t px, uint32_t py, EDataFormat f, Tensor& dw) const override
    { const conv_dims d = Dims(x, dw, g, s, px, py, f); dw.OverrideHost(); oracle_conv2d_kernels_gradient(&d, x.Values(), g.Values(), dw.Values()); }
    void ActivationGradient(EActivation a, float alpha, const Tensor& y, const Tensor& g, Tensor& dz) const override
    { dz.OverrideHost(); oracle_activation_gradient((int)a, alpha, y.Values(), g.Values(), dz.Values(), g.Length()); }
    static pool_dims PoolDims(const Tensor& x, const Tensor& y, uint32_t f, uint32_t s, EPoolingMode t, uint32_t px, uint32_t py, EDataFormat fmt)
    {
        pool_dims d{};
        if (fmt == NCHW) { d.W = x.Len(0); d.H = x.Len(1); d.C = x.Len(2); d.Wo = y.Len(0); d.Ho = y.Len(1); }
        else { d.C = x.Len(0); d.W = x.Len(1); d.H = x.Len(2); d.Wo = y.Len(1); d.Ho = y.Len(2); }
        d.N = x.Len(3); d.filter = f; d.stride = s; d.padX = px; d.padY = py; d.mode = (int)t; d.fmt = fmt;
        return d;
    }
    void Pool2D(const Tensor& x, uint32_t f, uint32_t s, EPoolingMode t, uint32_t px, uint32_t py, EDataFormat fmt, Tensor& y) const override
    { const pool_dims d = PoolDims(x, y, f, s, t, px, py, fmt); y.OverrideHost(); oracle_pool2d(&d, x.Values(), y.Values()); }
    void Pool2DGradient(const Tensor& y, const Tensor& x, const Tensor& g, uint32_t f, uint32_t s, EPoolingMode t, uint32_t px, uint32_t py, EDataFormat fmt, Tensor& dx) const override
    { const pool_dims d = PoolDims(x, y, f, s, t, px, py, fmt); dx.OverrideHost(); oracle_pool2d_gradient(&d, y.Values(), x.Values(), g.Values(), dx.Values()); }
    void UpSample2D(const Tensor& x, uint32_t s, Tensor& y) const override
    { y.OverrideHost(); oracle_upsample2d(x.Batch() * x.Depth(), x.Height(), x.Width(), s, x.Values(), y.Values()); }
    void UpSample2DGradient(const Tensor& g, uint32_t s, Tensor& dx) const override
    { dx.OverrideHost(); oracle_upsample2d_gradient(dx.Batch() * dx.Depth(), dx.Height(), dx.Width(), s, g.Values(), dx.Values()); }
    void ConstantPad2D(const Tensor& x, uint32_t l, uint32_t r, uint32_t t, uint32_t b, float v, Tensor& y) const override
    { y.OverrideHost(); oracle_constant_pad2d(x.Batch() * x.Depth(), x.Height(), x.Width(), l, r, t, b, v, x.Values(), y.Values()); }
    void AdamStep(Tensor& p, const Tensor& g, Tensor& m, Tensor& v, float lr, float b1, float b2, float eps) const override
    { oracle_adam_step(p.Values(), g.Values(), m.Values(), v.Values(), p.Length(), lr, b1, b2, eps); }
    void SgdStep(Tensor& p, const Tensor& g, float lr) const override { oracle_sgd_step(p.Values(), g.Values(), p.Length(), lr); }
    // Spatial mode only (what the conv stacks use): (Nn, G, S) = (N, C, H*W)
    void BatchNormalizationTrain(const Tensor& x, EBatchNormMode, const Tensor& gamma, const Tensor& beta, float momentum, float eps, Tensor* rm, Tensor* rv, Tensor& sm, Tensor& sv, Tensor& y) const override
    {
        y.OverrideHost(); sm.OverrideHost(); sv.OverrideHost();
        oracle_batch_norm_train(x.Batch(), x.Depth(), x.Height() * x.Width(), x.Values(), gamma.Values(), beta.Values(), momentum, eps, rm ? rm->Values() : nullptr, rv ? rv->Values() : nullptr,
                                sm.Values(), sv.Values(), y.Values());
    }
    void BatchNormalization(const Tensor& x, EBatchNormMode, const Tensor& gamma, const Tensor& beta, float eps, const Tensor* rm, const Tensor* rv, Tensor& y) const override
    {
        y.OverrideHost();
        oracle_batch_norm(x.Batch(), x.Depth(), x.Height() * x.Width(), x.Values(), gamma.Values(), beta.Values(), eps, rm->Values(), rv->Values(), y.Values());
    }
    void BatchNormalizationGradient(const Tensor& x, EBatchNormMode, const Tensor& gamma, float, const Tensor& g, const Tensor& sm, const Tensor& sv, Tensor& dgamma, Tensor& dbeta, bool, Tensor& dx) const override
    {
        dgamma.OverrideHost(); dbeta.OverrideHost(); dx.OverrideHost();
        oracle_batch_norm_gradient(x.Batch(), x.Depth(), x.Height() * x.Width(), x.Values(), gamma.Values(), g.Values(), sm.Values(), sv.Values(), dgamma.Values(), dbeta.Values(), dx.Values());
    }
};

static int g_Failed = 0, g_Run = 0;
#define TEST_METHOD(name) static void name(); static struct name##_reg { name##_reg() { Registry().push_back({#name, name}); } } name##_inst; static void name()
struct TestEntry { const char* name; void (*fn)(); };
static std::vector<TestEntry>& Registry() { static std::vector<TestEntry> r; return r; }
#define IsTrue(cond) do { if (!(cond)) { printf("    FAILED: %s (%s:%d)\n", #cond, __FILE__, __LINE__); ++g_Failed; } } while (0)

static float g_Tol = 1e-5f; // reference default Equals epsilon; TF32 mode uses the max-normalised 2e-3 bound instead
static bool g_Tf32 = false;
static bool Close(const Tensor& got, const Tensor& ref, float absEps = 1e-5f)
{
    return g_Tf32 ? got.MaxNormalisedError(ref) <= 2e-3f : got.Equals(ref, absEps);
}

// ---- TensorTests.cpp:352-425: literal vectors, run on the device backend ----
TEST_METHOD(Conv2D_Valid_1Kernel_1Batch)
{
    Tensor::SetDefaultOpMode(B200);
    Tensor t1(Shape(6, 6, 2)); t1.FillWithRange(0);
    Tensor t2(Shape(3, 3, 2)); t2.FillWithRange(0);
    Tensor r = t1.Conv2D(t2, 1, 0, NCHW);
    Tensor correct({ 5511, 5664, 5817, 5970, 6429, 6582, 6735, 6888, 7347, 7500, 7653, 7806, 8265, 8418, 8571, 8724 }, Shape(4, 4, 1));
    IsTrue(r.Equals(correct));
}

TEST_METHOD(Conv2D_Same_1Kernel_1Batch)
{
    Tensor::SetDefaultOpMode(B200);
    Tensor t1(Shape(6, 6, 2)); t1.FillWithRange(0);
    Tensor t2(Shape(3, 3, 2)); t2.FillWithRange(0);
    Tensor r = t1.Conv2D(t2, 1, Tensor::GetPadding(Same, 3), NCHW);
    Tensor correct({ 2492, 3674, 3794, 3914, 4034, 2624, 3765, 5511, 5664, 5817, 5970, 3855, 4413, 6429, 6582, 6735, 6888, 4431, 5061, 7347, 7500, 7653, 7806, 5007, 5709, 8265, 8418, 8571, 8724, 5583, 3416, 4898, 4982, 5066, 5150, 3260 }, Shape(6, 6, 1));
    IsTrue(r.Equals(correct));
}

// ---- TensorOpGpuTests.cpp:1196-1322, same shapes, CPU vs device ----
TEST_METHOD(Conv2D_Valid_CompareWithCpuResult)
{
    Tensor t(Shape(26, 26, 3, 3)); t.FillWithRand(11);
    Tensor kernals(Shape(3, 3, 3, 2)); kernals.FillWithRand(12);
    Tensor::SetForcedOpMode(CPU);
    Tensor r = t.Conv2D(kernals, 1, 0, NCHW);
    Tensor::SetForcedOpMode(B200);
    Tensor r2 = t.Conv2D(kernals, 1, 0, NCHW);
    IsTrue(r2.IsOnDevice());
    IsTrue(Close(r2, r));
}

TEST_METHOD(Conv2D_Same_CompareWithCpuResult)
{
    Tensor t(Shape(26, 26, 3, 3)); t.FillWithRand(11);
    Tensor kernals(Shape(3, 3, 3, 2)); kernals.FillWithRand(12);
    Tensor::SetForcedOpMode(CPU);
    Tensor r = t.Conv2D(kernals, 1, 1, NCHW);
    Tensor::SetForcedOpMode(B200);
    Tensor r2 = t.Conv2D(kernals, 1, 1, NCHW);
    IsTrue(Close(r2, r));
}

TEST_METHOD(Conv2D_NHWC_CompareWithCpuResult) // commented out in the reference's GPU tests; supported here
{
    Tensor t(Shape(3, 26, 26, 3)); t.FillWithRand(11);
    Tensor kernals(Shape(3, 3, 3, 2)); kernals.FillWithRand(12);
    Tensor::SetForcedOpMode(CPU);
    Tensor r = t.Conv2D(kernals, 1, 0, NHWC);
    Tensor::SetForcedOpMode(B200);
    Tensor r2 = t.Conv2D(kernals, 1, 0, NHWC);
    IsTrue(Close(r2, r));
}

TEST_METHOD(Conv2DBiasActivation_Valid_CompareWithCpuResult)
{
    Tensor t(Shape(26, 26, 3, 3)); t.FillWithRand(11);
    Tensor kernals(Shape(3, 3, 3, 2)); kernals.FillWithRand(12);
    Tensor bias(Shape(1, 1, 2, 1)); bias.FillWithRand(14);
    Tensor::SetForcedOpMode(CPU);
    Tensor r = t.Conv2DBiasActivation(kernals, 1, 0, bias, _ReLU, 1);
    Tensor::SetForcedOpMode(B200);
    Tensor r2 = t.Conv2DBiasActivation(kernals, 1, 0, bias, _ReLU, 1);
    IsTrue(Close(r2, r));
}

TEST_METHOD(Conv2DBiasGradient_CompareWithCpuResult)
{
    uint32_t features = 5;
    Tensor gradient(Shape(24, 24, features, 3)); gradient.FillWithRand(13);
    Tensor::SetForcedOpMode(CPU);
    Tensor biasGradient(Shape(1, 1, features, 1));
    gradient.Conv2DBiasGradient(gradient, biasGradient);
    Tensor::SetForcedOpMode(B200);
    Tensor biasGradient2(Shape(1, 1, features, 1));
    gradient.Conv2DBiasGradient(gradient, biasGradient2);
    IsTrue(biasGradient.Equals(biasGradient2, 0.0001f));
}

// Conv2dBiasActivationOp::ComputeGradientInternal (Conv2dBiasActivationOp.cpp:47-60): ActivationGradient, then Conv2DBiasGradient of
// its result -- the CPU side runs them as the reference does (two passes), TensorOpB200 in one fused pass.
TEST_METHOD(Conv2DBiasActivationGradient_CompareWithCpuResult)
{
    const EActivation acts[] = { _Sigmoid, _ReLU, _TanH, _ELU, _LeakyReLU };
    for (EActivation act : acts)
    {
        Tensor output(Shape(24, 24, 5, 3)); output.FillWithRand(15, act == _Sigmoid ? 0.f : -1.f, 1.f);
        Tensor gradient(Shape(24, 24, 5, 3)); gradient.FillWithRand(13);
        Tensor::SetForcedOpMode(CPU);
        Tensor dz(output.GetShape()), db(Shape(1, 1, 5, 1));
        gradient.Conv2DBiasActivationGradient(output, gradient, act, 0.2f, dz, db);
        Tensor::SetForcedOpMode(B200);
        Tensor dz2(output.GetShape()), db2(Shape(1, 1, 5, 1)), dz3(output.GetShape());
        gradient.Conv2DBiasActivationGradient(output, gradient, act, 0.2f, dz2, db2);
        gradient.ActivationGradient(act, 0.2f, output, gradient, dz3);
        IsTrue(dz.Equals(dz2, 0.f));    // element-wise fp32 in the reference's operation order: bit-exact
        IsTrue(dz.Equals(dz3, 0.f));
        IsTrue(db.Equals(db2, 0.0001f));
    }
}

// ---- TensorTests.cpp:427-504 literal vectors on the device, and TensorOpGpuTests-style CPU-vs-device comparisons (exact) ----
TEST_METHOD(Pool_Max_Valid_2Batches_Stride2)
{
    Tensor::SetForcedOpMode(B200);
    Tensor t1(Shape(6, 6, 1, 2)); t1.FillWithRange(0);
    Tensor r = t1.Pool2D(2, 2, MaxPool, 0, NCHW);
    Tensor correct({ 7, 9, 11, 19, 21, 23, 31, 33, 35, 43, 45, 47, 55, 57, 59, 67, 69, 71 }, Shape(3, 3, 1, 2));
    IsTrue(r.Equals(correct, 0.f));
    Tensor a = t1.Pool2D(2, 2, AvgPool, 0, NCHW);
    Tensor correctAvg({ 3.5f, 5.5f, 7.5f, 15.5f, 17.5f, 19.5f, 27.5f, 29.5f, 31.5f, 39.5f, 41.5f, 43.5f, 51.5f, 53.5f, 55.5f, 63.5f, 65.5f, 67.5f }, Shape(3, 3, 1, 2));
    IsTrue(a.Equals(correctAvg, 0.f));
}

TEST_METHOD(UpSample2D_2)
{
    Tensor::SetForcedOpMode(B200);
    Tensor t1(Shape(2, 2, 1, 2)); t1.FillWithRange(0);
    Tensor r = t1.UpSample2D(2);
    Tensor correct({ 0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3, 4, 4, 5, 5, 4, 4, 5, 5, 6, 6, 7, 7, 6, 6, 7, 7 }, Shape(4, 4, 1, 2));
    IsTrue(r.Equals(correct, 0.f));
}

TEST_METHOD(BatchNormalization_Spatial_CompareWithCpuResult) // TensorOpGpuTests.cpp:1767-1873: Shape(3,4,5,6), momentum 0.9, epsilon 0.001
{
    for (const Shape& shape : { Shape(3, 4, 5, 6), Shape(16, 16, 24, 8) })
    {
        Tensor x(shape); x.FillWithRand(31); Tensor g(shape); g.FillWithRand(32);
        Tensor gamma(Shape(1, 1, shape.Depth())); gamma.FillWithRand(33); Tensor beta(Shape(1, 1, shape.Depth())); beta.FillWithRand(34);
        Tensor out[2], sm[2], sv[2], rm[2], rv[2], dx[2], dg[2], db[2], inf[2];
        int i = 0;
        for (EOpMode mode : { CPU, B200 })
        {
            Tensor::SetForcedOpMode(mode);
            out[i] = Tensor(shape); sm[i] = Tensor(gamma.GetShape()); sv[i] = Tensor(gamma.GetShape()); dx[i] = Tensor(shape); dg[i] = Tensor(gamma.GetShape()); db[i] = Tensor(gamma.GetShape());
            rm[i] = Tensor(gamma.GetShape()); rm[i].FillWithRand(35); rv[i] = Tensor(gamma.GetShape()); rv[i].FillWithRand(36, 0.1f, 1.f); inf[i] = Tensor(shape);
            x.BatchNormalizationTrain(gamma, beta, 0.9f, 0.001f, &rm[i], &rv[i], sm[i], sv[i], out[i]);
            g.BatchNormalizationGradient(x, gamma, 0.001f, g, sm[0], sv[0], dg[i], db[i], true, dx[i]);     // both from the CPU statistics: only this op differs
            x.BatchNormalization(gamma, beta, 0.001f, &rm[i], &rv[i], inf[i]);
            ++i;
        }
        // reordered fp32 sums: tolerance, not equality (tests/test_batchnorm_gpu.py states the bound: 2e-5 max-normalised)
        IsTrue(out[1].MaxNormalisedError(out[0]) <= 2e-5f); IsTrue(sm[1].MaxNormalisedError(sm[0]) <= 2e-5f); IsTrue(sv[1].MaxNormalisedError(sv[0]) <= 2e-5f);
        IsTrue(rm[1].MaxNormalisedError(rm[0]) <= 2e-5f); IsTrue(rv[1].MaxNormalisedError(rv[0]) <= 2e-5f);
        IsTrue(dx[1].MaxNormalisedError(dx[0]) <= 2e-5f); IsTrue(dg[1].MaxNormalisedError(dg[0]) <= 2e-5f); IsTrue(db[1].MaxNormalisedError(db[0]) <= 2e-5f);
        IsTrue(inf[1].MaxNormalisedError(inf[0]) <= 2e-5f);
    }
}

TEST_METHOD(Resamplers_CompareWithCpuResult)
{
    Tensor x(Shape(16, 12, 5, 3)); x.FillWithRand(21);
    for (EPoolingMode mode : { MaxPool, AvgPool })
        for (uint32_t cfg = 0; cfg < 2; ++cfg)
        {
            const uint32_t f = cfg ? 3 : 2, st = 2, pad = cfg ? 1 : 0;   // cfg 0 takes the float4 2x2 fast path
            Tensor::SetForcedOpMode(CPU);
            Tensor y = x.Pool2D(f, st, mode, pad, NCHW);
            Tensor g(y.GetShape()); g.FillWithRand(22);
            Tensor dx(x.GetShape()); x.Pool2DGradient(y, x, g, f, st, mode, pad, NCHW, dx);
            Tensor::SetForcedOpMode(B200);
            Tensor y2 = x.Pool2D(f, st, mode, pad, NCHW);
            Tensor dx2(x.GetShape()); x.Pool2DGradient(y2, x, g, f, st, mode, pad, NCHW, dx2);
            IsTrue(y.Equals(y2, 0.f)); IsTrue(dx.Equals(dx2, 0.f));
        }
    Tensor::SetForcedOpMode(CPU);
    Tensor u = x.UpSample2D(2); Tensor du(x.GetShape()); x.UpSample2DGradient(u, 2, du);
    Tensor p = x.ConstantPad2D(1, 2, 0, 3, -1.5f);
    Tensor::SetForcedOpMode(B200);
    Tensor u2 = x.UpSample2D(2); Tensor du2(x.GetShape()); x.UpSample2DGradient(u2, 2, du2);
    Tensor p2 = x.ConstantPad2D(1, 2, 0, 3, -1.5f);
    IsTrue(u.Equals(u2, 0.f)); IsTrue(du.Equals(du2, 0.f)); IsTrue(p.Equals(p2, 0.f));
}

TEST_METHOD(Conv2DInputGradient_CompareWithCpuResult)
{
    Tensor input(Shape(26, 26, 3, 3));
    Tensor kernels(Shape(3, 3, 3, 2)); kernels.FillWithRand(12);
    Tensor gradient(Shape(24, 24, 2, 3)); gradient.FillWithRand(13);
    Tensor::SetForcedOpMode(CPU);
    Tensor inputGradient(input.GetShape());
    gradient.Conv2DInputsGradient(gradient, kernels, 1, 0, NCHW, inputGradient);
    Tensor::SetForcedOpMode(B200);
    Tensor inputGradient2(input.GetShape());
    gradient.Conv2DInputsGradient(gradient, kernels, 1, 0, NCHW, inputGradient2);
    IsTrue(Close(inputGradient2, inputGradient));
}

TEST_METHOD(Conv2DKernelsGradient_CompareWithCpuResult)
{
    Tensor input(Shape(26, 26, 3, 3)); input.FillWithRand(11);
    Tensor kernels(Shape(3, 3, 3, 2));
    Tensor gradient(Shape(24, 24, 2, 3)); gradient.FillWithRand(13);
    Tensor::SetForcedOpMode(CPU);
    Tensor kernelsGradient(kernels.GetShape());
    input.Conv2DKernelsGradient(input, gradient, 1, 0, NCHW, kernelsGradient);
    Tensor::SetForcedOpMode(B200);
    Tensor kernelsGradient2(kernels.GetShape());
    input.Conv2DKernelsGradient(input, gradient, 1, 0, NCHW, kernelsGradient2);
    IsTrue(Close(kernelsGradient2, kernelsGradient, 0.0001f)); // the reference allows 1e-4 here (TensorOpGpuTests.cpp:1320)
}

// ---- tensor-core sized layer + residency: intermediates stay on the device between ops ----
TEST_METHOD(ConvChain_StaysOnDevice_VggLikeBlock)
{
    Tensor x(Shape(64, 64, 64, 2)); x.FillWithRand(11);
    Tensor k1(Shape(3, 3, 64, 64)); k1.FillWithRand(12, -0.05f, 0.05f);
    Tensor k2(Shape(3, 3, 64, 128)); k2.FillWithRand(15, -0.05f, 0.05f);
    Tensor b1(Shape(1, 1, 64)); b1.FillWithRand(14, -0.1f, 0.1f);
    Tensor b2(Shape(1, 1, 128)); b2.FillWithRand(16, -0.1f, 0.1f);
    Tensor::SetForcedOpMode(CPU);
    Tensor y1 = x.Conv2DBiasActivation(k1, 1, 1, b1, _ReLU, 0);
    Tensor y2 = y1.Conv2DBiasActivation(k2, 1, 1, b2, _ReLU, 0);
    Tensor::SetForcedOpMode(B200);
    Tensor z1 = x.Conv2DBiasActivation(k1, 1, 1, b1, _ReLU, 0);
    IsTrue(z1.IsOnDevice());
    Tensor z2(y2.GetShape());
    z1.Conv2DBiasActivation(k2, 1, 1, b2, _ReLU, 0, z2);
    IsTrue(z1.IsOnDevice() && z2.IsOnDevice());      // no host round trip between the two layers
    IsTrue(z2.MaxNormalisedError(y2) <= (g_Tf32 ? 2e-3f : 1e-5f));
    // transposed-convolution identity: forward of Conv2DTranspose = input gradient (Tensor.cpp:1806-1810)
    Tensor kt(Shape(4, 4, 32, 128)); kt.FillWithRand(17, -0.05f, 0.05f); // (F,F,outDepth,inDepth)
    Tensor::SetForcedOpMode(CPU);
    Tensor up = y2.Conv2DTransposed(kt, 32, 2, 1, NCHW);
    Tensor::SetForcedOpMode(B200);
    Tensor up2 = z2.Conv2DTransposed(kt, 32, 2, 1, NCHW);
    IsTrue(up2.GetShape() == Shape(128, 128, 32, 2));
    IsTrue(up2.MaxNormalisedError(up) <= (g_Tf32 ? 2e-3f : 1e-5f));
}

TEST_METHOD(AdamAndSgdStep_CompareWithCpuResult)
{
    Tensor p(Shape(1000)), g(Shape(1000)), m(Shape(1000)), v(Shape(1000));
    p.FillWithRand(1); g.FillWithRand(2); m.FillWithRand(3, 0, 0.1f); v.FillWithRand(4, 0, 0.1f);
    Tensor p2(p), m2(m), v2(v);
    Tensor::GetOpFromMode(CPU)->AdamStep(p, g, m, v, 0.01f, 0.9f, 0.999f, 1e-8f);
    Tensor::GetOpFromMode(B200)->AdamStep(p2, g, m2, v2, 0.01f, 0.9f, 0.999f, 1e-8f);
    IsTrue(p.Equals(p2, 1e-6f) && m.Equals(m2, 1e-6f) && v.Equals(v2, 1e-6f));
    Tensor::GetOpFromMode(CPU)->SgdStep(p, g, 0.05f);
    Tensor::GetOpFromMode(B200)->SgdStep(p2, g, 0.05f);
    IsTrue(p.Equals(p2, 1e-6f));
}

int main(int argc, char** argv)
{
    const std::string mode = argc > 1 ? argv[1] : "fp32";
    g_Tf32 = mode == "tf32";
    Tensor::RegisterOp(CPU, new TensorOpOracle());
    Tensor::RegisterOp(B200, new TensorOpB200(g_Tf32 ? NB200_MATH_TF32 : NB200_MATH_FP32));
    for (const TestEntry& t : Registry())
    {
        const int before = g_Failed;
        ++g_Run;
        try { t.fn(); }
        catch (const std::exception& e) { printf("    EXCEPTION: %s\n", e.what()); ++g_Failed; }
        printf("[%s] %s (%s)\n", g_Failed == before ? " OK " : "FAIL", t.name, mode.c_str());
        Tensor::ClearForcedOpMode();
    }
    printf("%d tests, %d failed\n", g_Run, g_Failed);
    return g_Failed ? 1 : 0;
}
