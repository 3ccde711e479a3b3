// CUDA-core fp32 convolution kernels: the general path.
//
// They serve (a) NB200_MATH_FP32, (b) every problem the tcgen05 implicit-GEMM kernels do not take
// (NHWC, W not a multiple of 4, rectangular filters, padX != padY, tiny feature maps, ...), and
// (c) the HBM-bound small-channel layers (C or K in {1,3,6}), where tensor cores have nothing to add
// (4-25 FLOP/B, far below the ~127 FLOP/B ridge).
//
// Semantics follow the reference CPU ops (Neuro/src/Tensors/TensorOpCpu.cpp:1012-1184): cross-correlation,
// zero padding, outputs overwritten, caller-supplied extents on both sides.
#include "common.cuh"

namespace nb200
{
    namespace
    {
        constexpr int kFiltersPerThread = 8;   // register blocking over output channels (fwd) / input channels (dgrad)
        constexpr int kPixelsPerBlock = 128;

        struct Geo
        {
            int N, C, H, W, K, R, S, Ho, Wo, stride, padX, padY;
            ActStrides xs, ys;
        };

        Geo make_geo(const nb200_conv_desc& d)
        {
            Geo g;
            g.N = d.N; g.C = d.C; g.H = d.H; g.W = d.W; g.K = d.K; g.R = d.R; g.S = d.S; g.Ho = d.Ho; g.Wo = d.Wo;
            g.stride = d.stride; g.padX = d.padX; g.padY = d.padY;
            g.xs = act_strides(d.fmt, d.C, d.H, d.W);
            g.ys = act_strides(d.fmt, d.K, d.Ho, d.Wo);
            return g;
        }

        // y[n,k,oh,ow] = act(bias[k] + sum_{c,r,s} x[n,c,oh*st-pY+r,ow*st-pX+s] * w[k,c,r,s])
        // One thread = one output pixel x kFiltersPerThread filters; the x value is loaded once and reused across
        // the filter block, the weight loads are warp-uniform (one L1 broadcast each).
        __global__ void __launch_bounds__(kPixelsPerBlock)
        direct_fprop_kernel(Geo g, const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                            int act, float alpha, float* __restrict__ y)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const long long pixels = (long long)g.N * g.Ho * g.Wo;
            const long long p = (long long)blockIdx.x * kPixelsPerBlock + threadIdx.x;
            if (p >= pixels)
                return;
            const int ow = (int)(p % g.Wo);
            const int oh = (int)((p / g.Wo) % g.Ho);
            const int n = (int)(p / ((long long)g.Wo * g.Ho));
            const int k0 = blockIdx.y * kFiltersPerThread;
            const int h0 = oh * g.stride - g.padY, w0 = ow * g.stride - g.padX;

            float acc[kFiltersPerThread];
#pragma unroll
            for (int j = 0; j < kFiltersPerThread; ++j)
                acc[j] = 0.f;

            const long long filt = (long long)g.C * g.R * g.S;
            const float* xn = x + n * g.xs.n;
            for (int c = 0; c < g.C; ++c)
                for (int r = 0; r < g.R; ++r)
                {
                    const int ih = h0 + r;
                    if (ih < 0 || ih >= g.H)
                        continue;
                    for (int s = 0; s < g.S; ++s)
                    {
                        const int iw = w0 + s;
                        if (iw < 0 || iw >= g.W)
                            continue;
                        const float xv = __ldg(xn + c * g.xs.c + ih * g.xs.h + iw * g.xs.w);
                        const float* wp = w + (long long)k0 * filt + ((long long)c * g.R + r) * g.S + s;
#pragma unroll
                        for (int j = 0; j < kFiltersPerThread; ++j)
                            if (k0 + j < g.K)
                                acc[j] = fmaf(xv, __ldg(wp + j * filt), acc[j]);
                    }
                }

            float* yp = y + n * g.ys.n + oh * g.ys.h + ow * g.ys.w;
#pragma unroll
            for (int j = 0; j < kFiltersPerThread; ++j)
                if (k0 + j < g.K)
                {
                    float v = acc[j];
                    if (bias)
                        v += __ldg(bias + k0 + j);
                    yp[(k0 + j) * g.ys.c] = apply_activation(act, alpha, v);
                }
        }

        // dx[n,c,ih,iw] = sum_{k,r,s : (ih+pY-r) = oh*st, (iw+pX-s) = ow*st, oh<Ho, ow<Wo} w[k,c,r,s] * dy[n,k,oh,ow]
        // Gather form of the reference's scatter loops: every dx element is written exactly once (zeros where no tap
        // reaches), so no zero-fill pass and no atomics. One thread = one dx pixel x kFiltersPerThread input channels.
        __global__ void __launch_bounds__(kPixelsPerBlock)
        direct_dgrad_kernel(Geo g, const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const long long pixels = (long long)g.N * g.H * g.W;
            const long long p = (long long)blockIdx.x * kPixelsPerBlock + threadIdx.x;
            if (p >= pixels)
                return;
            const int iw = (int)(p % g.W);
            const int ih = (int)((p / g.W) % g.H);
            const int n = (int)(p / ((long long)g.W * g.H));
            const int c0 = blockIdx.y * kFiltersPerThread;

            float acc[kFiltersPerThread];
#pragma unroll
            for (int j = 0; j < kFiltersPerThread; ++j)
                acc[j] = 0.f;

            const long long filt = (long long)g.C * g.R * g.S;
            const int rs = g.R * g.S;
            const float* dyn = dy + n * g.ys.n;
            for (int r = 0; r < g.R; ++r)
            {
                const int th = ih + g.padY - r;
                if (th < 0 || th % g.stride)
                    continue;
                const int oh = th / g.stride;
                if (oh >= g.Ho)
                    continue;
                for (int s = 0; s < g.S; ++s)
                {
                    const int tw = iw + g.padX - s;
                    if (tw < 0 || tw % g.stride)
                        continue;
                    const int ow = tw / g.stride;
                    if (ow >= g.Wo)
                        continue;
                    const float* gp = dyn + oh * g.ys.h + ow * g.ys.w;
                    const float* wp = w + (long long)c0 * rs + r * g.S + s;
                    for (int k = 0; k < g.K; ++k)
                    {
                        const float gv = __ldg(gp + k * g.ys.c);
                        const float* wk = wp + k * filt;
#pragma unroll
                        for (int j = 0; j < kFiltersPerThread; ++j)
                            if (c0 + j < g.C)
                                acc[j] = fmaf(gv, __ldg(wk + j * rs), acc[j]);
                    }
                }
            }

            float* xp = dx + n * g.xs.n + ih * g.xs.h + iw * g.xs.w;
#pragma unroll
            for (int j = 0; j < kFiltersPerThread; ++j)
                if (c0 + j < g.C)
                    xp[(c0 + j) * g.xs.c] = acc[j];
        }

        // dw[k,c,r,s] = sum_{n,oh,ow} x[n,c,oh*st-pY+r,ow*st-pX+s] * dy[n,k,oh,ow]
        // Split reduction: blockIdx.z owns a contiguous slice of the (n,oh,ow) range and writes one partial per
        // (slice,k,c,r,s) into the workspace (or straight into dw when there is a single slice); a second kernel
        // adds the slices in a fixed order, so the result is deterministic (no atomics).
        // One block = one (k,c) pair; threads stride over the slice's pixels keeping R*S running sums.
        constexpr int kWgradThreads = 256;
        constexpr int kMaxTaps = 25;

        __global__ void __launch_bounds__(kWgradThreads)
        direct_wgrad_kernel(Geo g, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ out,
                            int slices)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const int k = blockIdx.x, c = blockIdx.y, slice = blockIdx.z;
            const int taps = g.R * g.S;
            const long long pixels = (long long)g.N * g.Ho * g.Wo;
            const long long per = (pixels + slices - 1) / slices;
            const long long begin = slice * per;
            const long long end = begin + per < pixels ? begin + per : pixels;

            __shared__ float red[kWgradThreads / 32];
            float* dst = out + (((long long)slice * g.K + k) * g.C + c) * taps;

            for (int t0 = 0; t0 < taps; t0 += kMaxTaps)
            {
                const int nt = taps - t0 < kMaxTaps ? taps - t0 : kMaxTaps;
                float acc[kMaxTaps];
#pragma unroll
                for (int t = 0; t < kMaxTaps; ++t)
                    acc[t] = 0.f;

                for (long long p = begin + threadIdx.x; p < end; p += kWgradThreads)
                {
                    const int ow = (int)(p % g.Wo);
                    const int oh = (int)((p / g.Wo) % g.Ho);
                    const int n = (int)(p / ((long long)g.Wo * g.Ho));
                    const float gv = __ldg(dy + n * g.ys.n + k * g.ys.c + oh * g.ys.h + ow * g.ys.w);
                    const float* xc = x + n * g.xs.n + c * g.xs.c;
                    const int h0 = oh * g.stride - g.padY, w0 = ow * g.stride - g.padX;
#pragma unroll
                    for (int t = 0; t < kMaxTaps; ++t)
                    {
                        if (t >= nt)
                            break;
                        const int r = (t0 + t) / g.S, s = (t0 + t) % g.S;
                        const int ih = h0 + r, iw = w0 + s;
                        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
                            acc[t] = fmaf(__ldg(xc + ih * g.xs.h + iw * g.xs.w), gv, acc[t]);
                    }
                }

#pragma unroll
                for (int t = 0; t < kMaxTaps; ++t)
                {
                    if (t >= nt)
                        break;
                    float v = acc[t];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
                        v += __shfl_xor_sync(0xffffffffu, v, o);
                    if ((threadIdx.x & 31) == 0)
                        red[threadIdx.x >> 5] = v;
                    __syncthreads();
                    if (threadIdx.x == 0)
                    {
                        float tot = 0.f;
#pragma unroll
                        for (int i = 0; i < kWgradThreads / 32; ++i)
                            tot += red[i];
                        dst[t0 + t] = tot;
                    }
                    __syncthreads();
                }
            }
        }

        __global__ void reduce_slices_kernel(const float* __restrict__ part, float* __restrict__ out, long long count, int slices)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= count)
                return;
            float v = 0.f;
            for (int s = 0; s < slices; ++s)
                v += part[s * count + i];
            out[i] = v;
        }

        int wgrad_slices(const nb200_conv_desc& d)
        {
            // enough blocks to fill 148 SMs a few times over, without slicing the reduction below ~2K pixels
            const long long pixels = (long long)d.N * d.Ho * d.Wo;
            const long long pairs = (long long)d.K * d.C;
            long long want = (148 * 8 + pairs - 1) / pairs;
            long long cap = pixels / 2048;
            if (cap < 1) cap = 1;
            if (want > cap) want = cap;
            if (want > 64) want = 64;
            return (int)(want < 1 ? 1 : want);
        }
    }

    int direct_forward(const nb200_conv_desc& d, const float* x, const float* w, const float* bias, int act, float alpha,
                       float* y, cudaStream_t st)
    {
        const Geo g = make_geo(d);
        const long long pixels = (long long)d.N * d.Ho * d.Wo;
        if (pixels == 0 || d.K == 0)
            return NB200_OK;
        dim3 grid(ceil_div(pixels, kPixelsPerBlock), ceil_div(d.K, kFiltersPerThread));
        NB200_CUDA_TRY(launch_kernel(direct_fprop_kernel, dim3(grid), dim3(kPixelsPerBlock), 0, st, g, x, w, bias, act, alpha, y));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int direct_input_gradient(const nb200_conv_desc& d, const float* dy, const float* w, float* dx, cudaStream_t st)
    {
        const Geo g = make_geo(d);
        const long long pixels = (long long)d.N * d.H * d.W;
        if (pixels == 0 || d.C == 0)
            return NB200_OK;
        dim3 grid(ceil_div(pixels, kPixelsPerBlock), ceil_div(d.C, kFiltersPerThread));
        NB200_CUDA_TRY(launch_kernel(direct_dgrad_kernel, dim3(grid), dim3(kPixelsPerBlock), 0, st, g, dy, w, dx));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    size_t direct_kernels_gradient_workspace(const nb200_conv_desc& d)
    {
        const int slices = wgrad_slices(d);
        return slices > 1 ? (size_t)slices * d.K * d.C * d.R * d.S * sizeof(float) : 0;
    }

    int direct_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes,
                                cudaStream_t st)
    {
        const Geo g = make_geo(d);
        const long long count = (long long)d.K * d.C * d.R * d.S;
        if (count == 0)
            return NB200_OK;
        if ((long long)d.N * d.Ho * d.Wo == 0)
        {
            NB200_CUDA_TRY(cudaMemsetAsync(dw, 0, count * sizeof(float), st));
            return NB200_OK;
        }
        if (d.K > 65535 * 32 || d.C > 65535)
            return fail(NB200_E_UNSUPPORTED, "direct wgrad: K or C too large for the launch grid");
        int slices = wgrad_slices(d);
        if (slices > 1 && wsBytes < (size_t)slices * count * sizeof(float))
            return fail(NB200_E_WORKSPACE, "direct wgrad needs %zu workspace bytes, got %zu", (size_t)slices * count * sizeof(float), wsBytes);
        float* out = slices > 1 ? (float*)ws : dw;
        dim3 grid(d.K, d.C, slices);
        NB200_CUDA_TRY(launch_kernel(direct_wgrad_kernel, dim3(grid), dim3(kWgradThreads), 0, st, g, x, dy, out, slices));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        if (slices > 1)
        {
            NB200_CUDA_TRY(launch_kernel(reduce_slices_kernel, dim3(ceil_div(count, 256)), dim3(256), 0, st, (const float*)ws, dw, count, slices));
            NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        }
        return NB200_OK;
    }
}
