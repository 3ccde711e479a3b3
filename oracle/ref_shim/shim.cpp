// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Host-only glue that lets the reference's own TensorOpCpu.cpp / TensorOpCpuMt.cpp / Shape.cpp
// (compiled UNMODIFIED from /root/reference by oracle/Makefile) run their convolution loops
// without the rest of Neuro_ (Tensor.cpp needs FreeImage/HDF5/cuDNN, Storage.cpp needs
// windows.h). It supplies, against the reference headers, only the
// handful of Tensor/Storage members those loops call, and a C entry point per op.
//
// Reference call sites this glue serves:
//   TensorOpCpu::Conv2D                 Neuro/src/Tensors/TensorOpCpu.cpp:1012
//   TensorOpCpu::Conv2DInputGradient    Neuro/src/Tensors/TensorOpCpu.cpp:1071
//   TensorOpCpu::Conv2DKernelsGradient  Neuro/src/Tensors/TensorOpCpu.cpp:1129
//   TensorOpCpuMt::{same three}         Neuro/src/Tensors/TensorOpCpuMt.cpp:173,220,278
//   TensorOpCpu::{Sigmoid,Tanh,ReLU,Elu,LeakyReLU}Gradient   Neuro/src/Tensors/TensorOpCpu.cpp:813-864
//   TensorOpCpu::AdamStep / SgdStep     Neuro/src/Tensors/TensorOpCpu.cpp:987-1009
//   TensorOpCpu::BatchNormalization{,Train,Gradient}          Neuro/src/Tensors/TensorOpCpu.cpp:1371-1480
// The last two groups are written in terms of Tensor arithmetic (Add/Sub/Mul/Div/Map/Sum/Mean/Pow/Sqrt, operators);
// in the reference those are one-line forwards `Op()->X(...)` (Tensor.cpp:489-975, 2720-2808) whose loops -- including
// all broadcasting rules -- live in TensorOpCpu.cpp. The forwards below do the same, so every number is still produced
// by the reference's own compiled loops.
#include <cstdlib>
#include <cstring>
#include <omp.h>

#include "Tensors/Tensor.h"
#include "Tensors/TensorOpCpu.h"
#include "Tensors/TensorOpCpuMt.h"

namespace Neuro
{
    // ---- Storage: plain host heap, single location ----
    Storage::Storage(int type, size_t size, const string& name)
        : m_Type(type), m_AllocSize(size), m_Size(size), m_Name(name)
    {
    }

    Storage::~Storage()
    {
        free(m_DataPtr);
        m_DataPtr = nullptr;
    }

    void Storage::AllocateOnHost() const
    {
        // read-only fast path: the conv loops call this once per element access from every thread
        if (m_DataPtr && m_DataLocation == Host)
            return;
        if (!m_DataPtr)
        {
            Storage* self = const_cast<Storage*>(this);
            self->m_DataPtr = (float*)calloc(m_AllocSize ? m_AllocSize : 1, sizeof(float));
        }
        m_DataLocation = Host;
    }

    void Storage::OverrideHost() { AllocateOnHost(); }
    void Storage::CopyToHost(bool) const { AllocateOnHost(); }
    const float* Storage::Data() const { AllocateOnHost(); return m_DataPtr; }
    float* Storage::Data() { AllocateOnHost(); return m_DataPtr; }

    // ---- Tensor: just enough for the conv loops ----
    TensorOpCpu* Tensor::g_DefaultOp = nullptr;
    TensorOpCpu* Tensor::g_ForcedOp = nullptr;
    TensorOpCpu* Tensor::g_OpCpu = nullptr;
    TensorOpCpu* Tensor::g_OpCpuMt = nullptr;
    TensorOpCpu* Tensor::g_OpCpuMkl = nullptr;
    TensorOpCpu* Tensor::g_OpGpu = nullptr;

    Tensor::Tensor(const Shape& shape, const string& name, EStorageType storageType)
        : m_Op(nullptr), m_Shape(shape), m_Storage(storageType, shape.Length, name), m_Name(name)
    {
    }

    void Tensor::CopyToHost(bool allowAlloc) const { m_Storage.CopyToHost(allowAlloc); }
    void Tensor::OverrideHost() { m_Storage.OverrideHost(); }
    float* Tensor::Values() { return m_Storage.Data(); }
    const float* Tensor::Values() const { return m_Storage.Data(); }

    void Tensor::Zero()
    {
        memset(m_Storage.Data(), 0, sizeof(float) * m_Shape.Length);
    }

    // Same bounds semantics as the reference (Tensor.cpp:2087-2093); used by Pool2DGradient's scatter.
    void Tensor::TrySet(float value, int w, int h, int d, int n)
    {
        if (h < 0 || h >= (int)Height() || w < 0 || w >= (int)Width() || d < 0 || d >= (int)Depth() || n < 0 || n > (int)Batch())
            return;
        Set(value, w, h, d, n);
    }

    // Tensor::Map (Tensor.cpp:814-818) forwards to the op; the op's own loop (TensorOpCpu.cpp:773-800) is the reference's.
    void Tensor::Map(const function<float(float, float)>& func, const Tensor& other, Tensor& result) const
    {
        alignas(16) static char dummy[64];
        reinterpret_cast<const TensorOpCpu*>(dummy)->TensorOpCpu::Map(func, *this, other, result);
    }

    // Same bounds semantics as the reference accessor (Tensor.cpp:2078-2084), including its
    // one-past-the-end tolerance on the batch index, which in-range callers never hit.
    float Tensor::TryGet(float def, int w, int h, int d, int n) const
    {
        const bool outside = w < 0 || w >= (int)Width() || h < 0 || h >= (int)Height() ||
                             d < 0 || d >= (int)Depth() || n < 0 || n > (int)Batch();
        return outside ? def : Get(w, h, d, n);
    }

    // ---- value semantics (deep host copy, as Storage.cpp:51-81 does for host-resident tensors) ----
    Storage::Storage(const Storage& other)
        : m_Type(other.m_Type), m_AllocSize(other.m_AllocSize), m_Size(other.m_Size), m_Name(other.m_Name)
    {
        if (other.m_DataPtr)
        {
            AllocateOnHost();
            memcpy(m_DataPtr, other.m_DataPtr, sizeof(float) * m_Size);
        }
    }

    Storage::Storage(Storage&& other) { *this = std::move(other); }

    Storage& Storage::operator=(const Storage& other)
    {
        if (this == &other)
            return *this;
        free(m_DataPtr);
        m_DataPtr = nullptr;
        m_DataLocation = None;
        m_Type = other.m_Type; m_AllocSize = other.m_AllocSize; m_Size = other.m_Size; m_Name = other.m_Name;
        if (other.m_DataPtr)
        {
            AllocateOnHost();
            memcpy(m_DataPtr, other.m_DataPtr, sizeof(float) * m_Size);
        }
        return *this;
    }

    Storage& Storage::operator=(Storage&& other)
    {
        if (this == &other)
            return *this;
        free(m_DataPtr);
        m_DataPtr = other.m_DataPtr; other.m_DataPtr = nullptr;
        m_Type = other.m_Type; m_AllocSize = other.m_AllocSize; m_Size = other.m_Size; m_Name = other.m_Name;
        m_DataLocation = other.m_DataLocation; other.m_DataLocation = None;
        return *this;
    }

    Tensor::Tensor(const Tensor& t) : m_Op(nullptr), m_Shape(t.m_Shape), m_Storage(t.m_Storage), m_Name(t.m_Name) {}
    Tensor::Tensor(Tensor&& t) : m_Op(nullptr), m_Shape(t.m_Shape), m_Storage(std::move(t.m_Storage)), m_Name(t.m_Name) {}
    Tensor& Tensor::operator=(const Tensor& t)
    {
        if (this != &t) { m_Shape = t.m_Shape; m_Storage = t.m_Storage; m_Name = t.m_Name; }
        return *this;
    }
    Tensor& Tensor::operator=(Tensor&& t)
    {
        if (this != &t) { m_Shape = t.m_Shape; m_Storage = std::move(t.m_Storage); m_Name = t.m_Name; }
        return *this;
    }

    // full-tensor copy (tau = 0 branch of Tensor.cpp:2096-2116)
    void Tensor::CopyTo(Tensor& target, float) const
    {
        memcpy(target.Values(), Values(), sizeof(float) * m_Shape.Length);
    }

    // ---- arithmetic forwards: Tensor::X -> the single-thread op's own loop ----
    namespace
    {
        alignas(16) char g_OpMem[64];
        const TensorOpCpu* RefOp() { return reinterpret_cast<const TensorOpCpu*>(g_OpMem); }
        Shape Broadcast(const Tensor& a, const Tensor& b)
        {
            return Shape(max(a.Width(), b.Width()), max(a.Height(), b.Height()), max(a.Depth(), b.Depth()), max(a.Batch(), b.Batch()));
        }
        Shape Reduced(const Tensor& t, EAxis axis)
        {
            const bool w = axis == GlobalAxis || axis == WidthAxis || axis == _01Axes || axis == _012Axes || axis == _013Axes;
            const bool h = axis == GlobalAxis || axis == HeightAxis || axis == _01Axes || axis == _012Axes || axis == _013Axes || axis == _123Axes;
            const bool d = axis == GlobalAxis || axis == DepthAxis || axis == _012Axes || axis == _123Axes;
            const bool n = axis == GlobalAxis || axis == BatchAxis || axis == _013Axes || axis == _123Axes;
            return Shape(w ? 1 : t.Width(), h ? 1 : t.Height(), d ? 1 : t.Depth(), n ? 1 : t.Batch());
        }
    }

    void Tensor::MulElem(const Tensor& t, Tensor& result) const { RefOp()->TensorOpCpu::Mul(1.f, *this, 1.f, t, result); }
    Tensor Tensor::MulElem(const Tensor& t) const { Tensor r(Broadcast(*this, t)); MulElem(t, r); return r; }
    void Tensor::Mul(float v, Tensor& result) const { RefOp()->TensorOpCpu::Mul(*this, v, result); }
    Tensor Tensor::Mul(float v) const { Tensor r(m_Shape); Mul(v, r); return r; }
    void Tensor::Div(const Tensor& t, Tensor& result) const { RefOp()->TensorOpCpu::Div(1.f, *this, 1.f, t, result); }
    Tensor Tensor::Div(const Tensor& t) const { Tensor r(m_Shape); Div(t, r); return r; }
    void Tensor::Div(float v, Tensor& result) const { Mul(1 / v, result); }
    Tensor Tensor::Div(float v) const { Tensor r(m_Shape); Div(v, r); return r; }
    void Tensor::Add(float alpha, float beta, const Tensor& t, Tensor& result) const { RefOp()->TensorOpCpu::Add(alpha, *this, beta, t, result); }
    void Tensor::Add(const Tensor& t, Tensor& result) const { Add(1, 1, t, result); }
    Tensor Tensor::Add(const Tensor& t) const { Tensor r(Broadcast(*this, t)); Add(t, r); return r; }
    void Tensor::Add(float v, Tensor& result) const { RefOp()->TensorOpCpu::Add(*this, v, result); }
    Tensor Tensor::Add(float v) const { Tensor r(m_Shape); Add(v, r); return r; }
    // TensorOpCpu::Sub is `Add(1, t1, -1, t2, output)` through the vtable (TensorOpCpu.cpp:78-81); the glue object has none, so that one line is applied here
    void Tensor::Sub(const Tensor& t, Tensor& result) const { RefOp()->TensorOpCpu::Add(1, *this, -1, t, result); }
    Tensor Tensor::Sub(const Tensor& t) const { Tensor r(Broadcast(*this, t)); Sub(t, r); return r; }
    void Tensor::Negated(Tensor& result) const { RefOp()->TensorOpCpu::Negate(*this, result); }
    Tensor Tensor::Negated() const { Tensor r(m_Shape); Negated(r); return r; }
    void Tensor::Inversed(float alpha, Tensor& result) const { RefOp()->TensorOpCpu::Inverse(alpha, *this, result); }
    Tensor Tensor::Inversed(float alpha) const { Tensor r(m_Shape); Inversed(alpha, r); return r; }
    void Tensor::Pow(float power, Tensor& result) const { RefOp()->TensorOpCpu::Pow(*this, power, result); }
    Tensor Tensor::Pow(float power) const { Tensor r(m_Shape); Pow(power, r); return r; }
    void Tensor::Sqrt(Tensor& output) const { RefOp()->TensorOpCpu::Sqrt(*this, output); }
    Tensor Tensor::Sqrt() const { Tensor r(m_Shape); Sqrt(r); return r; }
    void Tensor::Map(const function<float(float)>& func, Tensor& result) const { RefOp()->TensorOpCpu::Map(func, *this, result); }
    Tensor Tensor::Map(const function<float(float)>& func) const { Tensor r(m_Shape); Map(func, r); return r; }
    void Tensor::Sum(EAxis axis, Tensor& output) const { RefOp()->TensorOpCpu::Sum(*this, axis, output); }
    Tensor Tensor::Sum(EAxis axis) const { Tensor r(Reduced(*this, axis)); Sum(axis, r); return r; }
    void Tensor::Mean(EAxis axis, Tensor& output) const { RefOp()->TensorOpCpu::Mean(*this, axis, output); }
    Tensor Tensor::Mean(EAxis axis) const { Tensor r(Reduced(*this, axis)); Mean(axis, r); return r; }

    // free operators, Tensor.cpp:2720-2808
    Tensor operator*(const Tensor& t1, const Tensor& t2) { return t1.MulElem(t2); }
    Tensor operator*(const Tensor& t, float v) { return t.Mul(v); }
    Tensor operator/(const Tensor& t1, const Tensor& t2) { return t1.Div(t2); }
    Tensor operator/(const Tensor& t, float v) { return t.Div(v); }
    Tensor operator/(float v, const Tensor& t) { return t.Inversed(v); }
    Tensor operator+(const Tensor& t1, const Tensor& t2) { return t1.Add(t2); }
    Tensor operator+(const Tensor& t, float v) { return t.Add(v); }
    Tensor operator-(const Tensor& t1, const Tensor& t2) { return t1.Sub(t2); }
    Tensor operator-(const Tensor& t) { return t.Negated(); }
    Tensor pow(const Tensor& t, float p) { return t.Pow(p); }
    Tensor sqr(const Tensor& t) { return t.Pow(2); }
    Tensor sqrt(const Tensor& t) { return t.Sqrt(); }
    Tensor sum(const Tensor& t, EAxis axis) { return t.Sum(axis); }
    Tensor mean(const Tensor& t, EAxis axis) { return t.Mean(axis); }
}

using namespace Neuro;

namespace
{
    struct Dims { uint32_t d[4]; };

    Shape MakeShape(const uint32_t* d) { return Shape(d[0], d[1], d[2], d[3]); }

    void Load(Tensor& t, const float* src) { memcpy(t.Values(), src, sizeof(float) * t.Length()); }
    void Store(const Tensor& t, float* dst) { memcpy(dst, t.Values(), sizeof(float) * t.Length()); }

    // The op classes are stateless; their conv methods never touch `this`. Calling them
    // non-virtually on a dummy object avoids dragging in the vtable (and through it every other op).
    alignas(16) char g_Dummy[64];
    const TensorOpCpu* OpSt() { return reinterpret_cast<const TensorOpCpu*>(g_Dummy); }
    const TensorOpCpuMt* OpMt() { return reinterpret_cast<const TensorOpCpuMt*>(g_Dummy); }
}

#define REF_API extern "C" __attribute__((visibility("default")))

// All shapes are the reference's own Shape(d0,d1,d2,d3) = (fastest ... slowest) dimensions:
// NCHW tensors (W,H,C,N); NHWC tensors (C,W,H,N); kernels (S,R,C,K) in both formats.
// mt: 0 = TensorOpCpu (single thread), 1 = TensorOpCpuMt (PPL->OpenMP). fmt: 0 NCHW, 1 NHWC.

REF_API int neuro_ref_threads()
{
    return omp_get_max_threads();
}

REF_API void neuro_ref_set_threads(int n)
{
    omp_set_max_active_levels(2); // PPL load-balances nested parallel_for; OpenMP needs this enabled
    if (n > 0)
        omp_set_num_threads(n);
}

REF_API void neuro_ref_conv2d(int mt, int fmt, const float* x, const uint32_t* xDims, const float* w, const uint32_t* wDims,
                              uint32_t stride, uint32_t padX, uint32_t padY, float* y, const uint32_t* yDims)
{
    Tensor tx(MakeShape(xDims)), tw(MakeShape(wDims)), ty(MakeShape(yDims));
    Load(tx, x); Load(tw, w);
    if (mt)
        OpMt()->TensorOpCpuMt::Conv2D(tx, tw, stride, padX, padY, (EDataFormat)fmt, ty);
    else
        OpSt()->TensorOpCpu::Conv2D(tx, tw, stride, padX, padY, (EDataFormat)fmt, ty);
    Store(ty, y);
}

REF_API void neuro_ref_conv2d_input_gradient(int mt, int fmt, const float* dy, const uint32_t* dyDims, const float* w, const uint32_t* wDims,
                                             uint32_t stride, uint32_t padX, uint32_t padY, float* dx, const uint32_t* dxDims)
{
    Tensor tdy(MakeShape(dyDims)), tw(MakeShape(wDims)), tdx(MakeShape(dxDims));
    Load(tdy, dy); Load(tw, w);
    if (mt)
        OpMt()->TensorOpCpuMt::Conv2DInputGradient(tdy, tw, stride, padX, padY, (EDataFormat)fmt, tdx);
    else
        OpSt()->TensorOpCpu::Conv2DInputGradient(tdy, tw, stride, padX, padY, (EDataFormat)fmt, tdx);
    Store(tdx, dx);
}

REF_API void neuro_ref_conv2d_kernels_gradient(int mt, int fmt, const float* x, const uint32_t* xDims, const float* dy, const uint32_t* dyDims,
                                               uint32_t stride, uint32_t padX, uint32_t padY, float* dw, const uint32_t* dwDims)
{
    Tensor tx(MakeShape(xDims)), tdy(MakeShape(dyDims)), tdw(MakeShape(dwDims));
    Load(tx, x); Load(tdy, dy);
    if (mt)
        OpMt()->TensorOpCpuMt::Conv2DKernelsGradient(tx, tdy, stride, padX, padY, (EDataFormat)fmt, tdw);
    else
        OpSt()->TensorOpCpu::Conv2DKernelsGradient(tx, tdy, stride, padX, padY, (EDataFormat)fmt, tdw);
    Store(tdw, dw);
}

// dz = act'(y) * dy through the reference's own activation-gradient ops (what Tensor::ActivationGradient dispatches to,
// Conv2dBiasActivationOp.cpp:52). act follows EActivation; dims = Shape of all three tensors.
REF_API void neuro_ref_activation_gradient(int act, float alpha, const float* y, const float* dy, float* dz, const uint32_t* dims)
{
    Tensor ty(MakeShape(dims)), tg(MakeShape(dims)), tz(MakeShape(dims));
    Load(ty, y); Load(tg, dy);
    switch (act)
    {
    case 1: OpSt()->TensorOpCpu::SigmoidGradient(ty, tg, tz); break;
    case 2: OpSt()->TensorOpCpu::ReLUGradient(ty, tg, tz); break;
    case 3: OpSt()->TensorOpCpu::TanhGradient(ty, tg, tz); break;
    case 4: OpSt()->TensorOpCpu::EluGradient(ty, tg, alpha, tz); break;
    case 5: OpSt()->TensorOpCpu::LeakyReLUGradient(ty, tg, alpha, tz); break;
    default: memcpy(tz.Values(), tg.Values(), sizeof(float) * tg.Length()); break;
    }
    Store(tz, dz);
}

// ---- spatial resamplers (TensorOpCpu.cpp:1187-1369, 528-546), single-thread class; shapes as above ----
REF_API void neuro_ref_pool2d(int fmt, const float* x, const uint32_t* xDims, uint32_t filter, uint32_t stride, int mode, uint32_t padX, uint32_t padY,
                              float* y, const uint32_t* yDims)
{
    Tensor tx(MakeShape(xDims)), ty(MakeShape(yDims));
    Load(tx, x);
    OpSt()->TensorOpCpu::Pool2D(tx, filter, stride, (EPoolingMode)mode, padX, padY, (EDataFormat)fmt, ty);
    Store(ty, y);
}

REF_API void neuro_ref_pool2d_gradient(int fmt, const float* y, const uint32_t* yDims, const float* x, const uint32_t* xDims, const float* dy,
                                       uint32_t filter, uint32_t stride, int mode, uint32_t padX, uint32_t padY, float* dx)
{
    Tensor tx(MakeShape(xDims)), ty(MakeShape(yDims)), tdy(MakeShape(yDims)), tdx(MakeShape(xDims));
    Load(tx, x); Load(ty, y); Load(tdy, dy);
    OpSt()->TensorOpCpu::Pool2DGradient(ty, tx, tdy, filter, stride, (EPoolingMode)mode, padX, padY, (EDataFormat)fmt, tdx);
    Store(tdx, dx);
}

REF_API void neuro_ref_upsample2d(const float* x, const uint32_t* xDims, uint32_t scale, float* y, const uint32_t* yDims)
{
    Tensor tx(MakeShape(xDims)), ty(MakeShape(yDims));
    Load(tx, x);
    OpSt()->TensorOpCpu::UpSample2D(tx, scale, ty);
    Store(ty, y);
}

REF_API void neuro_ref_upsample2d_gradient(const float* dy, const uint32_t* yDims, uint32_t scale, float* dx, const uint32_t* xDims)
{
    Tensor tdy(MakeShape(yDims)), tdx(MakeShape(xDims));
    Load(tdy, dy);
    OpSt()->TensorOpCpu::UpSample2DGradient(tdy, scale, tdx);
    Store(tdx, dx);
}

REF_API void neuro_ref_constant_pad2d(const float* x, const uint32_t* xDims, uint32_t left, uint32_t right, uint32_t top, uint32_t bottom, float value,
                                      float* y, const uint32_t* yDims)
{
    Tensor tx(MakeShape(xDims)), ty(MakeShape(yDims));
    Load(tx, x);
    OpSt()->TensorOpCpu::ConstantPad2D(tx, left, right, top, bottom, value, ty);
    Store(ty, y);
}

// ---- optimiser updates (TensorOpCpu.cpp:987-1009); all four tensors are flat Shape(count) ----
REF_API void neuro_ref_adam_step(float* param, const float* grad, float* m, float* v, uint32_t count, float lr, float beta1, float beta2, float epsilon)
{
    const Shape flat(count);
    Tensor tp(flat), tg(flat), tm(flat), tv(flat);
    Load(tp, param); Load(tg, grad); Load(tm, m); Load(tv, v);
    OpSt()->TensorOpCpu::AdamStep(tp, tg, tm, tv, lr, beta1, beta2, epsilon);
    Store(tp, param); Store(tm, m); Store(tv, v);
}

REF_API void neuro_ref_sgd_step(float* param, const float* grad, uint32_t count, float lr)
{
    const Shape flat(count);
    Tensor tp(flat), tg(flat);
    Load(tp, param); Load(tg, grad);
    OpSt()->TensorOpCpu::SgdStep(tp, tg, lr);
    Store(tp, param);
}

// ---- batch normalisation (TensorOpCpu.cpp:1371-1480). mode follows EBatchNormMode (Types.h): 0 PerActivation, 1 Spatial, 2 Instance.
// xDims = Shape of input/output/gradients, pDims = Shape of gamma/beta/mean/variance tensors. running* may be NULL. ----
REF_API void neuro_ref_batch_norm_train(int mode, const float* x, const uint32_t* xDims, const float* gamma, const float* beta, const uint32_t* pDims,
                                        float momentum, float epsilon, float* runningMean, float* runningVar, float* saveMean, float* saveInvVar, float* y)
{
    Tensor tx(MakeShape(xDims)), tg(MakeShape(pDims)), tb(MakeShape(pDims)), trm(MakeShape(pDims)), trv(MakeShape(pDims)), tsm(MakeShape(pDims)),
        tsv(MakeShape(pDims)), ty(MakeShape(xDims));
    Load(tx, x); Load(tg, gamma); Load(tb, beta);
    if (runningMean) Load(trm, runningMean);
    if (runningVar) Load(trv, runningVar);
    tsm.Zero(); tsv.Zero();
    OpSt()->TensorOpCpu::BatchNormalizationTrain(tx, (EBatchNormMode)mode, tg, tb, momentum, epsilon, runningMean ? &trm : nullptr,
                                                 runningVar ? &trv : nullptr, tsm, tsv, ty);
    if (runningMean) Store(trm, runningMean);
    if (runningVar) Store(trv, runningVar);
    Store(tsm, saveMean); Store(tsv, saveInvVar); Store(ty, y);
}

REF_API void neuro_ref_batch_norm(int mode, const float* x, const uint32_t* xDims, const float* gamma, const float* beta, const uint32_t* pDims,
                                  float epsilon, const float* runningMean, const float* runningVar, float* y)
{
    Tensor tx(MakeShape(xDims)), tg(MakeShape(pDims)), tb(MakeShape(pDims)), trm(MakeShape(pDims)), trv(MakeShape(pDims)), ty(MakeShape(xDims));
    Load(tx, x); Load(tg, gamma); Load(tb, beta);
    if (runningMean) Load(trm, runningMean);
    if (runningVar) Load(trv, runningVar);
    OpSt()->TensorOpCpu::BatchNormalization(tx, (EBatchNormMode)mode, tg, tb, epsilon, runningMean ? &trm : nullptr, runningVar ? &trv : nullptr, ty);
    Store(ty, y);
}

REF_API void neuro_ref_batch_norm_gradient(int mode, const float* x, const uint32_t* xDims, const float* gamma, const uint32_t* pDims, float epsilon,
                                           const float* dy, const float* savedMean, const float* savedInvVar, float* dgamma, float* dbeta, int trainable,
                                           float* dx)
{
    Tensor tx(MakeShape(xDims)), tg(MakeShape(pDims)), tdy(MakeShape(xDims)), tsm(MakeShape(pDims)), tsv(MakeShape(pDims)), tdg(MakeShape(pDims)),
        tdb(MakeShape(pDims)), tdx(MakeShape(xDims));
    Load(tx, x); Load(tg, gamma); Load(tdy, dy); Load(tsm, savedMean); Load(tsv, savedInvVar);
    tdg.Zero(); tdb.Zero(); tdx.Zero();
    OpSt()->TensorOpCpu::BatchNormalizationGradient(tx, (EBatchNormMode)mode, tg, epsilon, tdy, tsm, tsv, tdg, tdb, trainable != 0, tdx);
    Store(tdg, dgamma); Store(tdb, dbeta); Store(tdx, dx);
}
