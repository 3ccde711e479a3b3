// Spatial resamplers on either side of the convolutions: Pool2D / Pool2DGradient,
// UpSample2D / UpSample2DGradient, ConstantPad2D. All HBM-bound (a handful of loads per stored element), so the design
// rules are the memory ones: one thread per OUTPUT element (gather form, every output written exactly once, no atomics,
// no zero-fill pass), consecutive threads on consecutive addresses of the tensor being written, grid-stride loops over a
// grid of a few waves of 148 SMs. Every kernel adds in exactly the order the reference's loops add, so results are
// bit-identical to TensorOpCpu (checked against the compiled reference in tests/).
#include <float.h>

#include "common.cuh"

namespace nb200
{
    namespace
    {
        constexpr int kThreads = 256;

        inline unsigned grid_for(long long n)
        {
            const long long blocks = (n + kThreads - 1) / kThreads;
            const long long cap = 148ll * 16;
            return (unsigned)(blocks < 1 ? 1 : blocks > cap ? cap : blocks);
        }

        struct Geo
        {
            int N, C, H, W, Ho, Wo, F, stride, padX, padY;
            ActStrides xs, ys; // element strides of the input-side / output-side tensors in the caller's data format
            int nhwc;
        };

        // decompose a flat index of a (N, C, h, w) tensor stored in the caller's format into its coordinates
        __device__ __forceinline__ void coords(long long i, int nhwc, int C, int Hh, int Ww, int& n, int& c, int& h, int& w)
        {
            if (nhwc)
            {
                c = (int)(i % C); i /= C;
                w = (int)(i % Ww); i /= Ww;
                h = (int)(i % Hh); n = (int)(i / Hh);
            }
            else
            {
                w = (int)(i % Ww); i /= Ww;
                h = (int)(i % Hh); i /= Hh;
                c = (int)(i % C); n = (int)(i / C);
            }
        }

        // TensorOpCpu::Pool2D (TensorOpCpu.cpp:1187-1246): window scanned (poolY, poolX); out-of-range taps read -FLT_MAX (max) or 0
        // (avg); the average always divides by F*F.
        template <bool MAX>
        __global__ void __launch_bounds__(kThreads) pool2d_kernel(const float* __restrict__ x, float* __restrict__ y, Geo g, long long total)
        {
            for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads)
            {
                int n, c, oh, ow;
                coords(i, g.nhwc, g.C, g.Ho, g.Wo, n, c, oh, ow);
                const float* xp = x + n * g.xs.n + c * g.xs.c;
                const int h0 = oh * g.stride - g.padY, w0 = ow * g.stride - g.padX;
                float acc = MAX ? -FLT_MAX : 0.f;
                for (int py = 0; py < g.F; ++py)
                    for (int px = 0; px < g.F; ++px)
                    {
                        const int h = h0 + py, w = w0 + px;
                        const bool in = h >= 0 && h < g.H && w >= 0 && w < g.W;
                        const float v = in ? __ldg(xp + h * g.xs.h + w * g.xs.w) : (MAX ? -FLT_MAX : 0.f);
                        acc = MAX ? fmaxf(acc, v) : __fadd_rn(acc, v);
                    }
                y[i] = MAX ? acc : __fdiv_rn(acc, (float)(g.F * g.F));
            }
        }

        // TensorOpCpu::Pool2DGradient (TensorOpCpu.cpp:1249-1338) in gather form. The reference walks the windows in (outH, outW)
        // order and scatters: max -> the FIRST window element (poolH, poolW order) equal to the pooled value receives the
        // gradient; avg -> every in-range element receives gradient / (F*F). Here each input element visits the windows
        // that cover it in the same (outH, outW) order and adds their contributions in that order.
        template <bool MAX>
        __global__ void __launch_bounds__(kThreads)
        pool2d_gradient_kernel(const float* __restrict__ y, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, Geo g,
                               long long total)
        {
            for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads)
            {
                int n, c, h, w;
                coords(i, g.nhwc, g.C, g.H, g.W, n, c, h, w);
                const float* xp = x + n * g.xs.n + c * g.xs.c;
                const long long yb = n * g.ys.n + c * g.ys.c;
                // windows with oh*stride - padY <= h < oh*stride - padY + F
                const int hh = h + g.padY, ww = w + g.padX;
                int ohLo = hh - g.F + 1; ohLo = ohLo <= 0 ? 0 : (ohLo + g.stride - 1) / g.stride;
                int owLo = ww - g.F + 1; owLo = owLo <= 0 ? 0 : (owLo + g.stride - 1) / g.stride;
                const int ohHi = min(g.Ho - 1, hh / g.stride), owHi = min(g.Wo - 1, ww / g.stride);
                float acc = 0.f;
                const float mine = MAX ? __ldg(xp + h * g.xs.h + w * g.xs.w) : 0.f;
                for (int oh = ohLo; oh <= ohHi; ++oh)
                    for (int ow = owLo; ow <= owHi; ++ow)
                    {
                        const long long yo = yb + oh * g.ys.h + ow * g.ys.w;
                        const float go = __ldg(dy + yo);
                        if (!MAX)
                        {
                            acc = __fadd_rn(acc, __fdiv_rn(go, (float)(g.F * g.F)));
                            continue;
                        }
                        const float m = __ldg(y + yo);
                        if (mine != m)
                            continue;
                        // am I the first element of this window equal to its maximum? (out-of-range taps hold -FLT_MAX)
                        const int h0 = oh * g.stride - g.padY, w0 = ow * g.stride - g.padX;
                        bool first = true;
                        for (int py = 0; py < g.F && first; ++py)
                            for (int px = 0; px < g.F; ++px)
                            {
                                const int h2 = h0 + py, w2 = w0 + px;
                                if (h2 == h && w2 == w)
                                {
                                    py = g.F; // reached myself: nobody earlier matched
                                    break;
                                }
                                const bool in = h2 >= 0 && h2 < g.H && w2 >= 0 && w2 < g.W;
                                const float v = in ? __ldg(xp + h2 * g.xs.h + w2 * g.xs.w) : -FLT_MAX;
                                if (v == m)
                                {
                                    first = false;
                                    break;
                                }
                            }
                        if (first)
                            acc = __fadd_rn(acc, go);
                    }
                dx[i] = acc;
            }
        }

        // TensorOpCpu::UpSample2D (TensorOpCpu.cpp:1340-1354): nearest neighbour, NCHW planes. One thread per output element.
        __global__ void __launch_bounds__(kThreads) upsample2d_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int s, long long total)
        {
            const int Wo = W * s, Ho = H * s;
            for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads)
            {
                const int ow = (int)(i % Wo);
                const long long r = i / Wo;
                const int oh = (int)(r % Ho);
                const long long plane = r / Ho;
                y[i] = __ldg(x + (plane * H + oh / s) * W + ow / s);
            }
        }

        // TensorOpCpu::UpSample2DGradient (TensorOpCpu.cpp:1357-1369): dx(w/s, h/s) += dy(w, h) walking h then w, i.e. each input
        // element adds its s x s block row by row, left to right, starting from 0.
        __global__ void __launch_bounds__(kThreads)
        upsample2d_gradient_kernel(const float* __restrict__ dy, float* __restrict__ dx, int H, int W, int s, long long total)
        {
            const int Wo = W * s;
            for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads)
            {
                const int w = (int)(i % W);
                const long long r = i / W;
                const int h = (int)(r % H);
                const long long plane = r / H;
                const float* p = dy + ((plane * H + h) * s) * (long long)Wo + (long long)w * s;
                float acc = 0.f;
                for (int a = 0; a < s; ++a)
                    for (int b = 0; b < s; ++b)
                        acc = __fadd_rn(acc, __ldg(p + (long long)a * Wo + b));
                dx[i] = acc;
            }
        }

        // TensorOpCpu::ConstantPad2D (TensorOpCpu.cpp:528-546), NCHW planes.
        __global__ void __launch_bounds__(kThreads)
        constant_pad2d_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int left, int top, int Ho, int Wo, float value, long long total)
        {
            for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads)
            {
                const int ow = (int)(i % Wo);
                const long long r = i / Wo;
                const int oh = (int)(r % Ho);
                const long long plane = r / Ho;
                const int h = oh - top, w = ow - left;
                y[i] = (h >= 0 && h < H && w >= 0 && w < W) ? __ldg(x + (plane * H + h) * W + w) : value;
            }
        }

        int check_pool(const nb200_pool_desc* d)
        {
            if (!d)
                return fail(NB200_E_INVALID, "null descriptor");
            if (d->N < 0 || d->C < 0 || d->H < 0 || d->W < 0 || d->Ho < 0 || d->Wo < 0 || d->filter < 1 || d->stride < 1 || d->padX < 0 || d->padY < 0)
                return fail(NB200_E_INVALID, "pooling descriptor out of range");
            if (d->fmt != NB200_NCHW && d->fmt != NB200_NHWC)
                return fail(NB200_E_INVALID, "unknown data format %d", d->fmt);
            if (d->mode != NB200_POOL_MAX && d->mode != NB200_POOL_AVG)
                return fail(NB200_E_INVALID, "unknown pooling mode %d", d->mode);
            if ((long long)d->N * d->C * d->H * d->W > 0)
            {
                // Tensor::GetPooling2DOutputShape (Tensor.cpp:1988-2007)
                if (d->H + 2 * d->padY < d->filter || d->W + 2 * d->padX < d->filter)
                    return fail(NB200_E_INVALID, "pooling window larger than padded input");
                const int ho = (d->H + 2 * d->padY - d->filter) / d->stride + 1, wo = (d->W + 2 * d->padX - d->filter) / d->stride + 1;
                if (ho != d->Ho || wo != d->Wo)
                    return fail(NB200_E_INVALID, "output extent %dx%d does not match GetPooling2DOutputShape %dx%d", d->Ho, d->Wo, ho, wo);
            }
            return NB200_OK;
        }

        Geo make_geo(const nb200_pool_desc& d)
        {
            Geo g;
            g.N = d.N; g.C = d.C; g.H = d.H; g.W = d.W; g.Ho = d.Ho; g.Wo = d.Wo; g.F = d.filter; g.stride = d.stride; g.padX = d.padX; g.padY = d.padY;
            g.xs = act_strides(d.fmt, d.C, d.H, d.W); g.ys = act_strides(d.fmt, d.C, d.Ho, d.Wo); g.nhwc = d.fmt == NB200_NHWC;
            return g;
        }

        int device_ok()
        {
            int dev = -1, major = 0;
            if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
            {
                cudaGetLastError();
                return fail(NB200_E_NO_DEVICE, "no usable CUDA device");
            }
            if (major != 10)
                return fail(NB200_E_NO_DEVICE, "device %d has compute capability %d.x; this library is built for sm_100a only", dev, major);
            return NB200_OK;
        }
    }
}

using namespace nb200;

extern "C"
{
    int nb200_pool2d(const nb200_pool_desc* d, const float* x, float* y, void* stream)
    {
        int rc = check_pool(d);
        if (rc) return rc;
        const long long total = (long long)d->N * d->C * d->Ho * d->Wo;
        if (total == 0) return NB200_OK;
        if (!x || !y) return fail(NB200_E_INVALID, "null tensor pointer");
        if ((rc = device_ok())) return rc;
        const Geo g = make_geo(*d);
        if (d->mode == NB200_POOL_MAX)
            pool2d_kernel<true><<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>(x, y, g, total);
        else
            pool2d_kernel<false><<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>(x, y, g, total);
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int nb200_pool2d_gradient(const nb200_pool_desc* d, const float* y, const float* x, const float* dy, float* dx, void* stream)
    {
        int rc = check_pool(d);
        if (rc) return rc;
        const long long total = (long long)d->N * d->C * d->H * d->W;
        if (total == 0) return NB200_OK;
        if (!dy || !dx || (d->mode == NB200_POOL_MAX && (!x || !y))) return fail(NB200_E_INVALID, "null tensor pointer");
        if ((rc = device_ok())) return rc;
        const Geo g = make_geo(*d);
        if (d->mode == NB200_POOL_MAX)
            pool2d_gradient_kernel<true><<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>(y, x, dy, dx, g, total);
        else
            pool2d_gradient_kernel<false><<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>(y, x, dy, dx, g, total);
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int nb200_upsample2d(int32_t N, int32_t C, int32_t H, int32_t W, int32_t scale, const float* x, float* y, void* stream)
    {
        if (N < 0 || C < 0 || H < 0 || W < 0 || scale < 1) return fail(NB200_E_INVALID, "up-sampling extents out of range");
        const long long total = (long long)N * C * H * W * scale * scale;
        if (total == 0) return NB200_OK;
        if (total > 0xFFFFFFFFll) return fail(NB200_E_INVALID, "tensor exceeds 2^32-1 elements");
        if (!x || !y) return fail(NB200_E_INVALID, "null tensor pointer");
        int rc = device_ok();
        if (rc) return rc;
        upsample2d_kernel<<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>(x, y, H, W, scale, total);
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int nb200_upsample2d_gradient(int32_t N, int32_t C, int32_t H, int32_t W, int32_t scale, const float* dy, float* dx, void* stream)
    {
        if (N < 0 || C < 0 || H < 0 || W < 0 || scale < 1) return fail(NB200_E_INVALID, "up-sampling extents out of range");
        const long long total = (long long)N * C * H * W;
        if (total == 0) return NB200_OK;
        if (total * scale * scale > 0xFFFFFFFFll) return fail(NB200_E_INVALID, "tensor exceeds 2^32-1 elements");
        if (!dy || !dx) return fail(NB200_E_INVALID, "null tensor pointer");
        int rc = device_ok();
        if (rc) return rc;
        upsample2d_gradient_kernel<<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>(dy, dx, H, W, scale, total);
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int nb200_constant_pad2d(int32_t N, int32_t C, int32_t H, int32_t W, int32_t left, int32_t right, int32_t top, int32_t bottom, float value,
                             const float* x, float* y, void* stream)
    {
        if (N < 0 || C < 0 || H < 0 || W < 0 || left < 0 || right < 0 || top < 0 || bottom < 0) return fail(NB200_E_INVALID, "padding extents out of range");
        const int Ho = H + top + bottom, Wo = W + left + right;
        const long long total = (long long)N * C * Ho * Wo;
        if (total == 0) return NB200_OK;
        if (total > 0xFFFFFFFFll) return fail(NB200_E_INVALID, "tensor exceeds 2^32-1 elements");
        if (!y || (!x && (long long)H * W > 0)) return fail(NB200_E_INVALID, "null tensor pointer");
        int rc = device_ok();
        if (rc) return rc;
        constant_pad2d_kernel<<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>(x, y, H, W, left, top, Ho, Wo, value, total);
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }
}
