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
