/*
 * TEST INFRASTRUCTURE ONLY -- the CPU oracle for the Conv2D hot path.
 *
 * A plain-C restatement of Neuro_'s reference convolution ops. Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the product path
 * (neuro__b200/, include/) never links, imports or calls it.
 *
 * Parity status: PINNED. tests/test_oracle.py checks every function here against
 *   (1) the five literal golden vectors of Neuro.Tests/src/TensorTests.cpp:352-425, and
 *   (2) the reference's own TensorOpCpu / TensorOpCpuMt sources compiled unmodified
 *       (oracle/_ref/libneuro_ref.so, built by oracle/Makefile `ref`), bit for bit, on seeded
 *       inputs covering NCHW/NHWC, stride 1-3, pad 0-3, padX != padY, ragged (non-dividing) strides
 *       and transposed-convolution output sizes; fixtures produced from that library are
 *       committed under tests/golden/ so the pin also holds where /root/reference is absent.
 *
 * Every fp32 function accumulates in exactly the order the reference does for each output
 * element, so it is bit-identical to the reference (compile with -ffp-contract=off). The loops
 * are re-nested so that independent output elements can be computed by different threads:
 * the reference's own multi-threaded variant relies on the same property
 * (Neuro/src/Tensors/TensorOpCpuMt.cpp:173-333).
 *
 * Layouts (Neuro/src/Tensors/Shape.cpp:11-22, Neuro/include/Tensors/Shape.h:93-103):
 *   fmt 0 = NCHW: x[((n*C+c)*H+h)*W+w]      fmt 1 = NHWC: x[((n*H+h)*W+w)*C+c]
 *   kernels are KCRS in both formats: w[((k*C+c)*R+r)*S+s]
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define API __attribute__((visibility("default")))

typedef struct
{
    int N, C, H, W;   /* input (or input-gradient) extent */
    int K, R, S;      /* filters */
    int Ho, Wo;       /* output (or output-gradient) extent, supplied by the caller */
    int stride, padX, padY;
    int fmt;          /* 0 NCHW, 1 NHWC */
} conv_dims;

static inline size_t xi(const conv_dims* d, int n, int c, int h, int w)
{
    return d->fmt == 0 ? (((size_t)n * d->C + c) * d->H + h) * d->W + w
                       : (((size_t)n * d->H + h) * d->W + w) * d->C + c;
}

static inline size_t yi(const conv_dims* d, int n, int k, int h, int w)
{
    return d->fmt == 0 ? (((size_t)n * d->K + k) * d->Ho + h) * d->Wo + w
                       : (((size_t)n * d->Ho + h) * d->Wo + w) * d->K + k;
}

static inline size_t wi(const conv_dims* d, int k, int c, int r, int s)
{
    return (((size_t)k * d->C + c) * d->R + r) * d->S + s;
}

/* Forward. Follows TensorOpCpu::Conv2D, Neuro/src/Tensors/TensorOpCpu.cpp:1012-1052:
 * cross-correlation, zero padding (out-of-range taps contribute 0*w, which is added like any
 * other term), local fp32 accumulator filled in (c, r, s) order, output overwritten. */
API void oracle_conv2d(const conv_dims* d, const float* x, const float* w, float* y)
{
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int n = 0; n < d->N; ++n)
    for (int k = 0; k < d->K; ++k)
    for (int oh = 0; oh < d->Ho; ++oh)
    for (int ow = 0; ow < d->Wo; ++ow)
    {
        const int h0 = oh * d->stride - d->padY, w0 = ow * d->stride - d->padX;
        float val = 0;
        for (int c = 0; c < d->C; ++c)
        for (int r = 0; r < d->R; ++r)
        for (int s = 0; s < d->S; ++s)
        {
            const int ih = h0 + r, iw = w0 + s;
            const float xv = (ih < 0 || ih >= d->H || iw < 0 || iw >= d->W) ? 0.f : x[xi(d, n, c, ih, iw)];
            val += xv * w[wi(d, k, c, r, s)];
        }
        y[yi(d, n, k, oh, ow)] = val;
    }
}

/* Input gradient (also the forward of Conv2DTranspose, Neuro/src/Tensors/Tensor.cpp:1806-1810).
 * Follows TensorOpCpu::Conv2DInputGradient, TensorOpCpu.cpp:1071-1126: dx zeroed, then every
 * (n,k,oh,ow) scatters w*dy into the in-range taps. For one dx element the reference's additions
 * arrive ordered by (k, oh, ow, r, s); the loops below keep that order per (n,c). The dx extent
 * (H,W) is the caller's, so rows/cols the forward never read stay zero. */
API void oracle_conv2d_input_gradient(const conv_dims* d, const float* dy, const float* w, float* dx)
{
    memset(dx, 0, sizeof(float) * (size_t)d->N * d->C * d->H * d->W);
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int n = 0; n < d->N; ++n)
    for (int c = 0; c < d->C; ++c)
    for (int k = 0; k < d->K; ++k)
    for (int oh = 0; oh < d->Ho; ++oh)
    for (int ow = 0; ow < d->Wo; ++ow)
    {
        const float g = dy[yi(d, n, k, oh, ow)];
        const int h0 = oh * d->stride - d->padY, w0 = ow * d->stride - d->padX;
        for (int r = 0; r < d->R; ++r)
        {
            const int ih = h0 + r;
            if (ih < 0 || ih >= d->H) continue;
            for (int s = 0; s < d->S; ++s)
            {
                const int iw = w0 + s;
                if (iw < 0 || iw >= d->W) continue;
                dx[xi(d, n, c, ih, iw)] += w[wi(d, k, c, r, s)] * g;
            }
        }
    }
}

/* Kernel gradient. Follows TensorOpCpu::Conv2DKernelsGradient, TensorOpCpu.cpp:1129-1184: dw
 * zeroed, then dw[k,c,r,s] += x * dy over all (n,oh,ow) with the tap in range; for one dw element
 * the additions arrive ordered by (n, oh, ow). The filter extent (R,S) is the caller's. */
API void oracle_conv2d_kernels_gradient(const conv_dims* d, const float* x, const float* dy, float* dw)
{
    memset(dw, 0, sizeof(float) * (size_t)d->K * d->C * d->R * d->S);
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int k = 0; k < d->K; ++k)
    for (int c = 0; c < d->C; ++c)
    for (int n = 0; n < d->N; ++n)
    for (int oh = 0; oh < d->Ho; ++oh)
    for (int ow = 0; ow < d->Wo; ++ow)
    {
        const float g = dy[yi(d, n, k, oh, ow)];
        const int h0 = oh * d->stride - d->padY, w0 = ow * d->stride - d->padX;
        for (int r = 0; r < d->R; ++r)
        {
            const int ih = h0 + r;
            if (ih < 0 || ih >= d->H) continue;
            for (int s = 0; s < d->S; ++s)
            {
                const int iw = w0 + s;
                if (iw < 0 || iw >= d->W) continue;
                dw[wi(d, k, c, r, s)] += x[xi(d, n, c, ih, iw)] * g;
            }
        }
    }
}

/* Same three contractions accumulated in fp64 and rounded once: the tie-breaker when the
 * reference's own fp32 running sums (O(len * eps) rounding on long reductions) are the larger
 * error source. Not reference behaviour; used only to bound errors. */
API void oracle_conv2d_f64(const conv_dims* d, const float* x, const float* w, float* y)
{
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int n = 0; n < d->N; ++n)
    for (int k = 0; k < d->K; ++k)
    for (int oh = 0; oh < d->Ho; ++oh)
    for (int ow = 0; ow < d->Wo; ++ow)
    {
        const int h0 = oh * d->stride - d->padY, w0 = ow * d->stride - d->padX;
        double val = 0;
        for (int c = 0; c < d->C; ++c)
        for (int r = 0; r < d->R; ++r)
        {
            const int ih = h0 + r;
            if (ih < 0 || ih >= d->H) continue;
            for (int s = 0; s < d->S; ++s)
            {
                const int iw = w0 + s;
                if (iw < 0 || iw >= d->W) continue;
                val += (double)x[xi(d, n, c, ih, iw)] * (double)w[wi(d, k, c, r, s)];
            }
        }
        y[yi(d, n, k, oh, ow)] = (float)val;
    }
}

API void oracle_conv2d_input_gradient_f64(const conv_dims* d, const float* dy, const float* w, float* dx)
{
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int n = 0; n < d->N; ++n)
    for (int c = 0; c < d->C; ++c)
    for (int ih = 0; ih < d->H; ++ih)
    for (int iw = 0; iw < d->W; ++iw)
    {
        double val = 0;
        for (int k = 0; k < d->K; ++k)
        for (int r = 0; r < d->R; ++r)
        {
            const int th = ih + d->padY - r;
            if (th < 0 || th % d->stride) continue;
            const int oh = th / d->stride;
            if (oh >= d->Ho) continue;
            for (int s = 0; s < d->S; ++s)
            {
                const int tw = iw + d->padX - s;
                if (tw < 0 || tw % d->stride) continue;
                const int ow = tw / d->stride;
                if (ow >= d->Wo) continue;
                val += (double)w[wi(d, k, c, r, s)] * (double)dy[yi(d, n, k, oh, ow)];
            }
        }
        dx[xi(d, n, c, ih, iw)] = (float)val;
    }
}

API void oracle_conv2d_kernels_gradient_f64(const conv_dims* d, const float* x, const float* dy, float* dw)
{
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int k = 0; k < d->K; ++k)
    for (int c = 0; c < d->C; ++c)
    for (int r = 0; r < d->R; ++r)
    for (int s = 0; s < d->S; ++s)
    {
        double val = 0;
        for (int n = 0; n < d->N; ++n)
        for (int oh = 0; oh < d->Ho; ++oh)
        {
            const int ih = oh * d->stride - d->padY + r;
            if (ih < 0 || ih >= d->H) continue;
            for (int ow = 0; ow < d->Wo; ++ow)
            {
                const int iw = ow * d->stride - d->padX + s;
                if (iw < 0 || iw >= d->W) continue;
                val += (double)x[xi(d, n, c, ih, iw)] * (double)dy[yi(d, n, k, oh, ow)];
            }
        }
        dw[wi(d, k, c, r, s)] = (float)val;
    }
}

/* Activations as the reference computes them (TensorOpCpu.cpp:807-864): exp evaluated in double
 * (the C++ overload picked for `(float)exp(-x)` with a float argument is exp(float) under
 * <cmath>; both g++ and MSVC resolve `exp(float)` to the float overload, so expf is used). */
static float activate(int act, float alpha, float v)
{
    switch (act)
    {
    case 1: return 1 / (1 + expf(-v));              /* _Sigmoid   :808 */
    case 2: return v > 0.f ? v : 0.f;               /* _ReLU      :832 */
    case 3: return 2 / (1 + expf(-2 * v)) - 1;      /* _TanH      :820 */
    case 4: return v >= 0 ? v : alpha * (expf(v) - 1); /* _ELU    :844 */
    case 5: return v >= 0 ? v : alpha * v;          /* _LeakyReLU :856 */
    default: return v;                              /* _Identity */
    }
}

/* Fused op. Follows TensorOpCpu::Conv2DBiasActivation, TensorOpCpu.cpp:1055-1062:
 * conv (NCHW) -> broadcast add of bias[k] -> activation. act uses EActivation numbering
 * (Neuro/include/Types.h:83-92). The reference only defines it for NCHW; fmt is honoured here
 * so the NHWC fused path of the product can be checked too. */
API void oracle_conv2d_bias_activation(const conv_dims* d, const float* x, const float* w, const float* bias,
                                       int act, float alpha, float* y)
{
    oracle_conv2d(d, x, w, y);
#pragma omp parallel for collapse(2)
    for (int n = 0; n < d->N; ++n)
    for (int k = 0; k < d->K; ++k)
    for (int oh = 0; oh < d->Ho; ++oh)
    for (int ow = 0; ow < d->Wo; ++ow)
    {
        const size_t i = yi(d, n, k, oh, ow);
        y[i] = activate(act, alpha, y[i] + bias[k]);
    }
}

/* Bias gradient. Follows TensorOpCpu::Conv2DBiasGradient, TensorOpCpu.cpp:1065-1068 =
 * Sum over the W,H,N axes (SumTemplate<1,1,0,1>, :284-296): one fp32 running sum per channel,
 * additions ordered by (n, h, w). */
API void oracle_conv2d_bias_gradient(const conv_dims* d, const float* dy, float* db)
{
#pragma omp parallel for
    for (int k = 0; k < d->K; ++k)
    {
        float acc = 0;
        for (int n = 0; n < d->N; ++n)
        for (int oh = 0; oh < d->Ho; ++oh)
        for (int ow = 0; ow < d->Wo; ++ow)
            acc += dy[yi(d, n, k, oh, ow)];
        db[k] = acc;
    }
}

/* Activation gradient through the activation's OUTPUT y, times the incoming gradient. Follows
 * TensorOpCpu::SigmoidGradient :813-816, TanhGradient :825-828, ReLUGradient :837-840, EluGradient :849-852,
 * LeakyReLUGradient :861-864 (dispatched by Tensor::ActivationGradient from
 * Conv2dBiasActivationOp::ComputeGradientInternal, Conv2dBiasActivationOp.cpp:47-53). act = EActivation. */
API void oracle_activation_gradient(int act, float alpha, const float* y, const float* dy, float* dz, size_t n)
{
#pragma omp parallel for
    for (long long i = 0; i < (long long)n; ++i)
    {
        const float x = y[i], x2 = dy[i];
        float r;
        switch (act)
        {
        case 1: r = x * (1 - x) * x2; break;
        case 2: r = x > 0 ? x2 : 0; break;
        case 3: r = (1 - x * x) * x2; break;
        case 4: r = (x > 0 ? 1 : (x + alpha)) * x2; break;
        case 5: r = (x > 0 ? 1 : alpha) * x2; break;
        default: r = x2; break;
        }
        dz[i] = r;
    }
}

/* Optimiser updates that follow the gradient exchange. Follow TensorOpCpu::AdamStep /
 * SgdStep, TensorOpCpu.cpp:987-1009:  m = b1*m + (1-b1)*g;  v = v*b2 + (1-b2)*g*g;
 * p = p - m / (sqrt(v) + eps) * lr   (sqrt evaluated in double, as `(float)::sqrt(x)`). */
API void oracle_adam_step(float* p, const float* g, float* m, float* v, size_t n,
                          float lr, float beta1, float beta2, float eps)
{
#pragma omp parallel for
    for (long long i = 0; i < (long long)n; ++i)
    {
        m[i] = beta1 * m[i] + (1 - beta1) * 1.f * g[i];
        v[i] = v[i] * beta2 + (1 - beta2) * 1.f * g[i] * g[i];
        p[i] = p[i] - m[i] / ((float)sqrt((double)v[i]) + eps) * lr;
    }
}

API void oracle_sgd_step(float* p, const float* g, size_t n, float lr)
{
#pragma omp parallel for
    for (long long i = 0; i < (long long)n; ++i)
        p[i] = 1 * p[i] + -lr * g[i];
}


/* ---- spatial resamplers ---- */
#include <float.h>

typedef struct pool_dims { int N, C, H, W, Ho, Wo, filter, stride, padX, padY, mode, fmt; } pool_dims;

static inline size_t ai(int fmt, int C, int H, int W, int n, int c, int h, int w)
{
    return fmt == 0 ? (((size_t)n * C + c) * H + h) * W + w : (((size_t)n * H + h) * W + w) * C + c;
}

/* Pooling. Follows TensorOpCpu::Pool2D, TensorOpCpu.cpp:1187-1246: window scanned (poolY, poolX); TryGet outside the
 * tensor returns -FLT_MAX (max) or 0 (avg); the average divides by filter*filter. mode 0 = MaxPool, 1 = AvgPool. */
API void oracle_pool2d(const pool_dims* d, const float* x, float* y)
{
#pragma omp parallel for collapse(2)
    for (int n = 0; n < d->N; ++n)
    for (int c = 0; c < d->C; ++c)
    for (int oh = 0; oh < d->Ho; ++oh)
    for (int ow = 0; ow < d->Wo; ++ow)
    {
        const int h = oh * d->stride - d->padY, w = ow * d->stride - d->padX;
        float acc = d->mode == 0 ? -FLT_MAX : 0.f;
        for (int py = 0; py < d->filter; ++py)
        for (int px = 0; px < d->filter; ++px)
        {
            const int in = h + py >= 0 && h + py < d->H && w + px >= 0 && w + px < d->W;
            if (d->mode == 0)
            {
                const float v = in ? x[ai(d->fmt, d->C, d->H, d->W, n, c, h + py, w + px)] : -FLT_MAX;
                acc = acc > v ? acc : v;
            }
            else
                acc += in ? x[ai(d->fmt, d->C, d->H, d->W, n, c, h + py, w + px)] : 0.f;
        }
        y[ai(d->fmt, d->C, d->Ho, d->Wo, n, c, oh, ow)] = d->mode == 0 ? acc : acc / (float)(d->filter * d->filter);
    }
}

/* Pooling gradient. Follows TensorOpCpu::Pool2DGradient, TensorOpCpu.cpp:1249-1338: dx zeroed, then windows visited in
 * (outH, outW) order; max: the first element (poolH, poolW order) whose value equals the pooled output gets += dy (a
 * match on an out-of-range tap is dropped by TrySet but still ends the search); avg: every in-range element gets
 * += dy / filter^2. */
API void oracle_pool2d_gradient(const pool_dims* d, const float* y, const float* x, const float* dy, float* dx)
{
    memset(dx, 0, sizeof(float) * (size_t)d->N * d->C * d->H * d->W);
#pragma omp parallel for collapse(2)
    for (int n = 0; n < d->N; ++n)
    for (int c = 0; c < d->C; ++c)
    for (int oh = 0; oh < d->Ho; ++oh)
    for (int ow = 0; ow < d->Wo; ++ow)
    {
        const int h = oh * d->stride - d->padY, w = ow * d->stride - d->padX;
        const size_t yo = ai(d->fmt, d->C, d->Ho, d->Wo, n, c, oh, ow);
        if (d->mode == 0)
        {
            int found = 0;
            for (int py = 0; py < d->filter && !found; ++py)
            for (int px = 0; px < d->filter; ++px)
            {
                const int in = h + py >= 0 && h + py < d->H && w + px >= 0 && w + px < d->W;
                const float v = in ? x[ai(d->fmt, d->C, d->H, d->W, n, c, h + py, w + px)] : -FLT_MAX;
                if (v == y[yo])
                {
                    if (in)
                        dx[ai(d->fmt, d->C, d->H, d->W, n, c, h + py, w + px)] += dy[yo];
                    found = 1;
                    break;
                }
            }
        }
        else
        {
            const float f2 = (float)(d->filter * d->filter);
            for (int py = 0; py < d->filter; ++py)
            for (int px = 0; px < d->filter; ++px)
                if (h + py >= 0 && h + py < d->H && w + px >= 0 && w + px < d->W)
                    dx[ai(d->fmt, d->C, d->H, d->W, n, c, h + py, w + px)] += dy[yo] / f2;
        }
    }
}

/* Nearest-neighbour up-sampling. Follows TensorOpCpu::UpSample2D, TensorOpCpu.cpp:1340-1354 (NCHW planes). */
API void oracle_upsample2d(int planes, int H, int W, int s, const float* x, float* y)
{
#pragma omp parallel for
    for (int p = 0; p < planes; ++p)
    for (int h = 0; h < H; ++h)
    for (int w = 0; w < W; ++w)
        for (int oh = h * s; oh < (h + 1) * s; ++oh)
        for (int ow = w * s; ow < (w + 1) * s; ++ow)
            y[((size_t)p * H * s + oh) * W * s + ow] = x[((size_t)p * H + h) * W + w];
}

/* Follows TensorOpCpu::UpSample2DGradient, TensorOpCpu.cpp:1357-1369: dx zeroed, dx(w/s, h/s) += dy(w, h) walking h then w. */
API void oracle_upsample2d_gradient(int planes, int H, int W, int s, const float* dy, float* dx)
{
    memset(dx, 0, sizeof(float) * (size_t)planes * H * W);
#pragma omp parallel for
    for (int p = 0; p < planes; ++p)
    for (int oh = 0; oh < H * s; ++oh)
    for (int ow = 0; ow < W * s; ++ow)
        dx[((size_t)p * H + oh / s) * W + ow / s] += dy[((size_t)p * H * s + oh) * W * s + ow];
}

/* Follows TensorOpCpu::ConstantPad2D, TensorOpCpu.cpp:528-546 (NCHW planes). */
API void oracle_constant_pad2d(int planes, int H, int W, int left, int right, int top, int bottom, float value, const float* x, float* y)
{
    const int Ho = H + top + bottom, Wo = W + left + right;
#pragma omp parallel for
    for (int p = 0; p < planes; ++p)
    for (int h = 0; h < Ho; ++h)
    for (int w = 0; w < Wo; ++w)
    {
        float v = value;
        if (w >= left && h >= top && w < W + left && h < H + top)
            v = x[((size_t)p * H + (h - top)) * W + (w - left)];
        y[((size_t)p * Ho + h) * Wo + w] = v;
    }
}


/* ---- batch normalisation ----
 * Follows TensorOpCpu::BatchNormalizationTrain / BatchNormalizationGradient / BatchNormalization,
 * Neuro/src/Tensors/TensorOpCpu.cpp:1392-1480, 1371-1389. The reference writes them with Tensor operators; every operator is a
 * loop of TensorOpCpu.cpp:28-75 (Add: alpha*a + beta*b), :136-183 (Mul: alpha*a*beta*b), :186-196 (Mul by scalar),
 * :285-299 + :327-351 (Sum: zero, then += walking the tensor n, d, h, w), :354-376 (Mean: Sum, then times 1/count -- Tensor::Div(float)
 * is Mul(1/v), Tensor.cpp:554-557), :379-389 (Pow through double ::pow), :433-442 (Sqrt), :472-481 (Inverse). The statements below
 * apply those loops in the same order with the same intermediate roundings, so results are bit-identical to the compiled
 * reference (tests/test_oracle.py).
 *
 * One layout serves the three EBatchNormMode values. A tensor is viewed as x[n][g][s], n < Nn, g < G, s < S; statistics are taken
 * per group g over (n, s); gamma/beta/mean/variance hold G values:
 *   Spatial        (axis _013Axes, m = W*H*N):  Nn = N, G = C,     S = H*W
 *   PerActivation  (axis BatchAxis, m = N):     Nn = N, G = W*H*C, S = 1
 *   Instance       (axis _01Axes,  m = W*H):    Nn = 1, G = C*N,   S = H*W   (gamma is Shape(1,1,C,N) there) */
static inline size_t bi(int G, int S, int n, int g, int s) { return ((size_t)n * G + g) * S + s; }

API void oracle_batch_norm_train(int Nn, int G, int S, const float* x, const float* gamma, const float* beta, float momentum, float epsilon,
                                 float* runningMean, float* runningVar, float* saveMean, float* saveInvVar, float* y)
{
    const float m = (float)((unsigned)S * (unsigned)Nn);
    if (m == 1)
    {
        memcpy(y, x, sizeof(float) * (size_t)Nn * G * S); /* "cannot normalize single values so just copy input to output" */
        return;
    }
    const float invm = 1 / m;
    for (int g = 0; g < G; ++g)
    {
        float sum = 0.f;
        for (int n = 0; n < Nn; ++n)
            for (int s = 0; s < S; ++s)
                sum += x[bi(G, S, n, g, s)];
        const float mean = sum * invm;
        saveMean[g] = mean;
        float sq = 0.f;
        for (int n = 0; n < Nn; ++n)
            for (int s = 0; s < S; ++s)
            {
                const float xmu = 1.f * x[bi(G, S, n, g, s)] + -1.f * mean;
                sq += (float)pow((double)xmu, (double)2.f);
            }
        const float var = sq * invm;
        const float inv = 1.f / sqrtf(var + epsilon);
        saveInvVar[g] = inv;
        for (int n = 0; n < Nn; ++n)
            for (int s = 0; s < S; ++s)
            {
                const float xmu = 1.f * x[bi(G, S, n, g, s)] + -1.f * mean;
                const float xnorm = 1.f * xmu * 1.f * inv;
                y[bi(G, S, n, g, s)] = 1.f * (1.f * xnorm * 1.f * gamma[g]) + 1.f * beta[g];
            }
        if (runningMean)
            runningMean[g] = (1 - momentum) * runningMean[g] + momentum * mean;
        if (runningVar)
        {
            const float temp = var * (m / (m - 1)); /* "according to the original BN paper" */
            runningVar[g] = (1 - momentum) * runningVar[g] + momentum * temp;
        }
    }
}

/* Inference form, TensorOpCpu.cpp:1371-1389 with running statistics. */
API void oracle_batch_norm(int Nn, int G, int S, const float* x, const float* gamma, const float* beta, float epsilon,
                           const float* runningMean, const float* runningVar, float* y)
{
    for (int g = 0; g < G; ++g)
    {
        const float f = 1.f / sqrtf(runningVar[g] + epsilon);
        for (int n = 0; n < Nn; ++n)
            for (int s = 0; s < S; ++s)
            {
                const float xmu = 1.f * x[bi(G, S, n, g, s)] + -1.f * runningMean[g];
                const float xnorm = 1.f * xmu * 1.f * f;
                y[bi(G, S, n, g, s)] = 1.f * (1.f * xnorm * 1.f * gamma[g]) + 1.f * beta[g];
            }
    }
}

API void oracle_batch_norm_gradient(int Nn, int G, int S, const float* x, const float* gamma, const float* dy, const float* savedMean,
                                    const float* savedInvVar, float* dgamma, float* dbeta, float* dx)
{
    const float m = (float)((unsigned)S * (unsigned)Nn);
    if (m == 1)
    {
        memcpy(dx, dy, sizeof(float) * (size_t)Nn * G * S);
        memset(dgamma, 0, sizeof(float) * G);
        memset(dbeta, 0, sizeof(float) * G);
        return;
    }
    const float invm = 1 / m;
    for (int g = 0; g < G; ++g)
    {
        const float mean = savedMean[g], inv = savedInvVar[g], ninv = -inv;
        float sumA = 0.f, sumB = 0.f, sumC = 0.f, sumG = 0.f, sumD = 0.f;
        for (int n = 0; n < Nn; ++n)
            for (int s = 0; s < S; ++s)
            {
                const size_t i = bi(G, S, n, g, s);
                const float xmu = 1.f * x[i] + -1.f * mean;
                const float xnorm = 1.f * xmu * 1.f * inv;
                const float dxn = 1.f * dy[i] * 1.f * gamma[g];
                sumA += 1.f * dxn * 1.f * xmu;
                sumB += 1.f * dxn * 1.f * ninv;
                sumC += xmu * -2.f;
                sumG += 1.f * dy[i] * 1.f * xnorm;
                sumD += dy[i];
            }
        const float dVar = 1.f * (sumA * -.5f) * 1.f * (float)pow((double)inv, (double)3.f);
        const float dMu = 1.f * sumB + 1.f * (1.f * dVar * 1.f * (sumC * invm));
        const float dMuM = dMu * invm;
        for (int n = 0; n < Nn; ++n)
            for (int s = 0; s < S; ++s)
            {
                const size_t i = bi(G, S, n, g, s);
                const float xmu = 1.f * x[i] + -1.f * mean;
                const float dxn = 1.f * dy[i] * 1.f * gamma[g];
                const float a = 1.f * dxn * 1.f * inv;
                const float b = ((1.f * dVar * 1.f * xmu) * 2.f) * invm;
                dx[i] = 1.f * (1.f * a + 1.f * b) + 1.f * dMuM;
            }
        dgamma[g] = sumG;
        dbeta[g] = sumD;
    }
}
