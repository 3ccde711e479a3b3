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

        // decompose a flat index of a (N, C, h, w) tensor stored in the caller's format into its coordinates. Tensors hold
        // < 2^32 elements (Neuro::Shape::Length is uint32_t), so the index arithmetic is 32-bit: a 64-bit division costs
        // ~100 instructions on the CUDA cores and made the first version of these kernels issue-bound at 0.5-1.7 TB/s.
        __device__ __forceinline__ void coords(unsigned i, int nhwc, unsigned C, unsigned Hh, unsigned Ww, unsigned& n, unsigned& c, unsigned& h, unsigned& w)
        {
            if (nhwc)
            {
                c = i % C; i /= C;
                w = i % Ww; i /= Ww;
                h = i % Hh; n = i / Hh;
            }
            else
            {
                w = i % Ww; i /= Ww;
                h = i % Hh; i /= Hh;
                c = i % C; n = i / C;
            }
        }

        // TensorOpCpu::Pool2D (TensorOpCpu.cpp:1187-1246): window scanned (poolY, poolX); out-of-range taps read -FLT_MAX (max) or 0
        // (avg); the average always divides by F*F.
        template <bool MAX>
        __global__ void __launch_bounds__(kThreads) pool2d_kernel(const float* __restrict__ x, float* __restrict__ y, Geo g, long long total)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            for (long long i64 = (long long)blockIdx.x * kThreads + threadIdx.x; i64 < total; i64 += (long long)gridDim.x * kThreads)
            {
                const unsigned i = (unsigned)i64;
                unsigned n, c, oh, ow;
                coords(i, g.nhwc, g.C, g.Ho, g.Wo, n, c, oh, ow);
                const float* xp = x + n * g.xs.n + c * g.xs.c;
                const int h0 = (int)oh * g.stride - g.padY, w0 = (int)ow * g.stride - g.padX;
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

        // The pooling every model in the reference uses (VGG16/19, the conv autoencoder): 2x2 windows, stride 2, no padding,
        // NCHW, even H, W % 8 == 0. Thread = 4 consecutive outputs of one row = two aligned float4 loads from each of the two
        // input rows and one float4 store: every byte moves once, fully coalesced. Same operation order as above.
        template <bool MAX>
        __global__ void __launch_bounds__(kThreads)
        pool2x2_kernel(const float4* __restrict__ x, float4* __restrict__ y, unsigned H, unsigned W4, unsigned Ho, unsigned Wo4, long long quads)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            for (long long q64 = (long long)blockIdx.x * kThreads + threadIdx.x; q64 < quads; q64 += (long long)gridDim.x * kThreads)
            {
                const unsigned q = (unsigned)q64;
                const unsigned ow4 = q % Wo4, t = q / Wo4, oh = t % Ho, plane = t / Ho;
                const float4* r0 = x + ((size_t)plane * H + 2 * oh) * W4 + 2 * ow4;
                const float4 a0 = __ldcs(r0), a1 = __ldcs(r0 + 1), b0 = __ldcs(r0 + W4), b1 = __ldcs(r0 + W4 + 1);
                float4 o;
                if (MAX)
                {
                    o.x = fmaxf(fmaxf(fmaxf(fmaxf(-FLT_MAX, a0.x), a0.y), b0.x), b0.y);
                    o.y = fmaxf(fmaxf(fmaxf(fmaxf(-FLT_MAX, a0.z), a0.w), b0.z), b0.w);
                    o.z = fmaxf(fmaxf(fmaxf(fmaxf(-FLT_MAX, a1.x), a1.y), b1.x), b1.y);
                    o.w = fmaxf(fmaxf(fmaxf(fmaxf(-FLT_MAX, a1.z), a1.w), b1.z), b1.w);
                }
                else
                {
                    o.x = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a0.x, a0.y), b0.x), b0.y), 4.f);   // 0 + a is a, exactly
                    o.y = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a0.z, a0.w), b0.z), b0.w), 4.f);
                    o.z = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a1.x, a1.y), b1.x), b1.y), 4.f);
                    o.w = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a1.z, a1.w), b1.z), b1.w), 4.f);
                }
                y[q] = o;
            }
        }

        // first-match rule of the reference's max-pool gradient on one 2x2 window (scan order e00, e01, e10, e11)
        __device__ __forceinline__ void first_max(float e00, float e01, float e10, float e11, float m, float g, float& d00, float& d01, float& d10,
                                                  float& d11)
        {
            const bool m0 = e00 == m, m1 = !m0 && e01 == m, m2 = !m0 && !m1 && e10 == m, m3 = !m0 && !m1 && !m2 && e11 == m;
            d00 = m0 ? g : 0.f; d01 = m1 ? g : 0.f; d10 = m2 ? g : 0.f; d11 = m3 ? g : 0.f;
        }

        template <bool MAX>
        __global__ void __launch_bounds__(kThreads)
        pool2x2_gradient_kernel(const float4* __restrict__ y, const float4* __restrict__ x, const float4* __restrict__ dy, float4* __restrict__ dx,
                                unsigned H, unsigned W4, unsigned Ho, unsigned Wo4, long long quads)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            for (long long q64 = (long long)blockIdx.x * kThreads + threadIdx.x; q64 < quads; q64 += (long long)gridDim.x * kThreads)
            {
                const unsigned q = (unsigned)q64;
                const unsigned ow4 = q % Wo4, t = q / Wo4, oh = t % Ho, plane = t / Ho;
                const size_t r0 = ((size_t)plane * H + 2 * oh) * W4 + 2 * ow4;
                const float4 g = __ldcs(dy + q);
                float4 t0, t1, u0, u1; // dx rows 2*oh (t) and 2*oh+1 (u), two float4 each
                if (MAX)
                {
                    const float4 m = __ldcs(y + q);
                    const float4 a0 = __ldcs(x + r0), a1 = __ldcs(x + r0 + 1), b0 = __ldcs(x + r0 + W4), b1 = __ldcs(x + r0 + W4 + 1);
                    first_max(a0.x, a0.y, b0.x, b0.y, m.x, g.x, t0.x, t0.y, u0.x, u0.y);
                    first_max(a0.z, a0.w, b0.z, b0.w, m.y, g.y, t0.z, t0.w, u0.z, u0.w);
                    first_max(a1.x, a1.y, b1.x, b1.y, m.z, g.z, t1.x, t1.y, u1.x, u1.y);
                    first_max(a1.z, a1.w, b1.z, b1.w, m.w, g.w, t1.z, t1.w, u1.z, u1.w);
                }
                else
                {
                    const float qx = __fdiv_rn(g.x, 4.f), qy = __fdiv_rn(g.y, 4.f), qz = __fdiv_rn(g.z, 4.f), qw = __fdiv_rn(g.w, 4.f);
                    t0 = make_float4(qx, qx, qy, qy); t1 = make_float4(qz, qz, qw, qw);
                    u0 = t0; u1 = t1;
                }
                __stcs(dx + r0, t0); __stcs(dx + r0 + 1, t1); __stcs(dx + r0 + W4, u0); __stcs(dx + r0 + W4 + 1, u1);
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
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            for (long long i64 = (long long)blockIdx.x * kThreads + threadIdx.x; i64 < total; i64 += (long long)gridDim.x * kThreads)
            {
                const unsigned i = (unsigned)i64;
                unsigned n, c, hu, wu;
                coords(i, g.nhwc, g.C, g.H, g.W, n, c, hu, wu);
                const int h = (int)hu, w = (int)wu;
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

        // TensorOpCpu::UpSample2D (TensorOpCpu.cpp:1340-1354): nearest neighbour, NCHW planes. Thread = VEC consecutive outputs
        // of one row (VEC = 4: one 16-byte store; the 1-2 source values they replicate come from L1).
        template <int VEC>
        __global__ void __launch_bounds__(kThreads)
        upsample2d_kernel(const float* __restrict__ x, float* __restrict__ y, unsigned H, unsigned W, unsigned s, long long groups)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const unsigned Wo = W * s, Ho = H * s, WoG = Wo / VEC;
            for (long long i64 = (long long)blockIdx.x * kThreads + threadIdx.x; i64 < groups; i64 += (long long)gridDim.x * kThreads)
            {
                const unsigned i = (unsigned)i64;
                const unsigned ow0 = (i % WoG) * VEC, r = i / WoG, oh = r % Ho, plane = r / Ho;
                const float* src = x + ((size_t)plane * H + oh / s) * W;
                float v[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    v[j] = __ldg(src + (ow0 + j) / s);
                float* dst = y + (size_t)r * Wo + ow0;
                if (VEC == 4)
                    __stcs(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
                else
                    dst[0] = v[0];
            }
        }

        // TensorOpCpu::UpSample2DGradient (TensorOpCpu.cpp:1357-1369): dx(w/s, h/s) += dy(w, h) walking h then w, i.e. each input
        // element adds its s x s block row by row, left to right, starting from 0.
        __global__ void __launch_bounds__(kThreads)
        upsample2d_gradient_kernel(const float* __restrict__ dy, float* __restrict__ dx, unsigned H, unsigned W, unsigned s, long long total)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const unsigned Wo = W * s;
            for (long long i64 = (long long)blockIdx.x * kThreads + threadIdx.x; i64 < total; i64 += (long long)gridDim.x * kThreads)
            {
                const unsigned i = (unsigned)i64;
                const unsigned w = i % W, r = i / W, h = r % H, plane = r / H;
                const float* p = dy + ((size_t)plane * H + h) * s * Wo + (size_t)w * s;
                float acc = 0.f;
                for (unsigned a = 0; a < s; ++a)
                    for (unsigned b = 0; b < s; ++b)
                        acc = __fadd_rn(acc, __ldg(p + (size_t)a * Wo + b));
                dx[i] = acc;
            }
        }

        // scale 2 (every model in the reference): thread = 2 consecutive dx elements = one aligned float4 from each of the two
        // dy rows; same order of additions (row a = 0: left, right; row a = 1: left, right)
        __global__ void __launch_bounds__(kThreads)
        upsample2x_gradient_kernel(const float4* __restrict__ dy, float2* __restrict__ dx, unsigned H, unsigned W2, long long pairs)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            for (long long i64 = (long long)blockIdx.x * kThreads + threadIdx.x; i64 < pairs; i64 += (long long)gridDim.x * kThreads)
            {
                const unsigned i = (unsigned)i64;
                const unsigned w2 = i % W2, r = i / W2; // r = plane * H + h
                const float4* p = dy + (size_t)r * 2 * W2 + w2; // dy row 2h of this plane: W2 float4 per row
                const float4 a = __ldcs(p), b = __ldcs(p + W2);
                float2 o;
                o.x = __fadd_rn(__fadd_rn(__fadd_rn(a.x, a.y), b.x), b.y);
                o.y = __fadd_rn(__fadd_rn(__fadd_rn(a.z, a.w), b.z), b.w);
                dx[i] = o;
            }
        }

        // TensorOpCpu::ConstantPad2D (TensorOpCpu.cpp:528-546), NCHW planes. Thread = VEC consecutive outputs of one row.
        template <int VEC>
        __global__ void __launch_bounds__(kThreads)
        constant_pad2d_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int left, int top, unsigned Ho, unsigned Wo, float value,
                              long long groups)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const unsigned WoG = Wo / VEC;
            for (long long i64 = (long long)blockIdx.x * kThreads + threadIdx.x; i64 < groups; i64 += (long long)gridDim.x * kThreads)
            {
                const unsigned i = (unsigned)i64;
                const unsigned ow0 = (i % WoG) * VEC, r = i / WoG, oh = r % Ho, plane = r / Ho;
                const int h = (int)oh - top;
                const bool rowIn = h >= 0 && h < H;
                const float* src = x + ((size_t)plane * H + (rowIn ? h : 0)) * W;
                float v[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                {
                    const int w = (int)(ow0 + j) - left;
                    v[j] = (rowIn && w >= 0 && w < W) ? __ldg(src + w) : value;
                }
                float* dst = y + (size_t)r * Wo + ow0;
                if (VEC == 4)
                    __stcs(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
                else
                    dst[0] = v[0];
            }
        }

        // 2x2 stride-2 unpadded NCHW pooling on rows of whole float4 pairs; every pointer that is passed must be 16-byte aligned
        // Layout pass between the reference's two data formats: per image, dst[c][p] = src[p][c] (toNchw) or dst[p][c] = src[c][p],
        // p = h*W + w; 32 x 32 tiles through padded shared memory, both sides coalesced. Bound: HBM, 8 bytes per element.
        __global__ void __launch_bounds__(256)
        layout_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            // src is [rows][cols] per image, dst [cols][rows]
            __shared__ float tile[32][33];
            const long long img = (long long)blockIdx.z * rows * cols;
            const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
            const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
            for (int i = 0; i < 32; i += 8)
            {
                const int r = r0 + ty + i, c = c0 + tx;
                if (r < rows && c < cols)
                    tile[ty + i][tx] = __ldg(src + img + (long long)r * cols + c);
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 32; i += 8)
            {
                const int c = c0 + ty + i, r = r0 + tx;
                if (r < rows && c < cols)
                    dst[img + (long long)c * rows + r] = tile[tx][ty + i];
            }
        }

        bool pool2x2_fast(const nb200_pool_desc& d, const void* a, const void* b, const void* c, const void* e)
        {
            return d.fmt == NB200_NCHW && d.filter == 2 && d.stride == 2 && d.padX == 0 && d.padY == 0 && d.H % 2 == 0 && d.W % 8 == 0 &&
                   (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)e) & 15) == 0;
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
            if ((long long)d->N * d->C * d->H * d->W > 0xFFFFFFFFll || (long long)d->N * d->C * d->Ho * d->Wo > 0xFFFFFFFFll)
                return fail(NB200_E_INVALID, "tensor exceeds 2^32-1 elements"); // Neuro::Shape::Length is uint32_t; the kernels index with 32 bits
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

namespace nb200
{
    // NHWC (N, H*W, C) <-> NCHW (N, C, H*W): the tensor-core kernels are NCHW; NHWC problems run them between two layout passes
    int layout_transpose(const float* src, float* dst, int N, int C, int HW, bool toNchw, cudaStream_t st)
    {
        if ((long long)N * C * HW == 0)
            return NB200_OK;
        if (N > 65535 || (C + 31) / 32 > 65535 || (HW + 31) / 32 > 65535)
            return fail(NB200_E_UNSUPPORTED, "layout pass: extent too large for one grid");
        const int rows = toNchw ? HW : C, cols = toNchw ? C : HW;
        const dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32), (unsigned)N);
        NB200_CUDA_TRY(launch_kernel(layout_transpose_kernel, grid, dim3(256), 0, st, src, dst, rows, cols));
        count_launch();
        return NB200_OK;
    }
}

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
        if (pool2x2_fast(*d, x, y, nullptr, nullptr))
        {
            const long long quads = total / 4;
            if (d->mode == NB200_POOL_MAX)
                NB200_CUDA_TRY(launch_kernel(pool2x2_kernel<true>, dim3(grid_for(quads)), dim3(kThreads), 0, (cudaStream_t)stream, (const float4*)x, (float4*)y, d->H, d->W / 4, d->Ho, d->Wo / 4, quads));
            else
                NB200_CUDA_TRY(launch_kernel(pool2x2_kernel<false>, dim3(grid_for(quads)), dim3(kThreads), 0, (cudaStream_t)stream, (const float4*)x, (float4*)y, d->H, d->W / 4, d->Ho, d->Wo / 4, quads));
        }
        else if (d->mode == NB200_POOL_MAX)
            NB200_CUDA_TRY(launch_kernel(pool2d_kernel<true>, dim3(grid_for(total)), dim3(kThreads), 0, (cudaStream_t)stream, x, y, g, total));
        else
            NB200_CUDA_TRY(launch_kernel(pool2d_kernel<false>, dim3(grid_for(total)), dim3(kThreads), 0, (cudaStream_t)stream, x, y, g, total));
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
        if (pool2x2_fast(*d, d->mode == NB200_POOL_MAX ? x : dy, d->mode == NB200_POOL_MAX ? y : dy, dy, dx))
        {
            const long long quads = (long long)d->N * d->C * d->Ho * d->Wo / 4;
            if (d->mode == NB200_POOL_MAX)
                NB200_CUDA_TRY(launch_kernel(pool2x2_gradient_kernel<true>, dim3(grid_for(quads)), dim3(kThreads), 0, (cudaStream_t)stream, (const float4*)y, (const float4*)x, (const float4*)dy, (float4*)dx, d->H, d->W / 4, d->Ho, d->Wo / 4, quads));
            else
                NB200_CUDA_TRY(launch_kernel(pool2x2_gradient_kernel<false>, dim3(grid_for(quads)), dim3(kThreads), 0, (cudaStream_t)stream, nullptr, nullptr, (const float4*)dy, (float4*)dx, d->H, d->W / 4, d->Ho, d->Wo / 4, quads));
        }
        else if (d->mode == NB200_POOL_MAX)
            NB200_CUDA_TRY(launch_kernel(pool2d_gradient_kernel<true>, dim3(grid_for(total)), dim3(kThreads), 0, (cudaStream_t)stream, y, x, dy, dx, g, total));
        else
            NB200_CUDA_TRY(launch_kernel(pool2d_gradient_kernel<false>, dim3(grid_for(total)), dim3(kThreads), 0, (cudaStream_t)stream, y, x, dy, dx, g, total));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int32_t nb200_pool2d_gradient_activation_supported(const nb200_pool_desc* d)
    {
        return check_pool(d) == NB200_OK && pool_act_bias_supported(*d) ? 1 : 0;
    }

    size_t nb200_pool2d_gradient_activation_workspace_bytes(const nb200_pool_desc* d)
    {
        return nb200_pool2d_gradient_activation_supported(d) ? pool_act_bias_workspace(*d) : 0;
    }

    int nb200_pool2d_gradient_activation(const nb200_pool_desc* d, int32_t act, float alpha, const float* y, const float* x, const float* dy,
                                         float* dz, float* db, void* workspace, size_t workspace_bytes, void* stream)
    {
        int rc = check_pool(d);
        if (rc) return rc;
        if (act < NB200_ACT_IDENTITY || act > NB200_ACT_LEAKY_RELU)
            return fail(NB200_E_INVALID, "activation %d has no gradient here", act);
        if (!pool_act_bias_supported(*d) || ((((uintptr_t)y | (uintptr_t)x | (uintptr_t)dy | (uintptr_t)dz) & 15) != 0))
            return fail(NB200_E_UNSUPPORTED, "fused pool + activation gradient takes 2x2 stride-2 max pooling of aligned NCHW tensors only");
        if ((long long)d->N * d->C * d->H * d->W == 0)
        {
            if (db && d->C > 0)
            {
                if ((rc = device_ok())) return rc;
                NB200_CUDA_TRY(cudaMemsetAsync(db, 0, d->C * sizeof(float), (cudaStream_t)stream));
            }
            return NB200_OK;
        }
        if (!y || !x || !dy || !dz)
            return fail(NB200_E_INVALID, "null tensor pointer");
        if ((rc = device_ok())) return rc;
        return pool_act_bias_gradient(*d, act, alpha, y, x, dy, dz, db, workspace, workspace_bytes, (cudaStream_t)stream);
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
        if ((W * scale) % 4 == 0 && ((uintptr_t)y & 15) == 0)
            NB200_CUDA_TRY(launch_kernel(upsample2d_kernel<4>, dim3(grid_for(total / 4)), dim3(kThreads), 0, (cudaStream_t)stream, x, y, H, W, scale, total / 4));
        else
            NB200_CUDA_TRY(launch_kernel(upsample2d_kernel<1>, dim3(grid_for(total)), dim3(kThreads), 0, (cudaStream_t)stream, x, y, H, W, scale, total));
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
        if (scale == 2 && W % 2 == 0 && (((uintptr_t)dy & 15) | ((uintptr_t)dx & 7)) == 0)
            NB200_CUDA_TRY(launch_kernel(upsample2x_gradient_kernel, dim3(grid_for(total / 2)), dim3(kThreads), 0, (cudaStream_t)stream, (const float4*)dy, (float2*)dx, H, W / 2, total / 2));
        else
            NB200_CUDA_TRY(launch_kernel(upsample2d_gradient_kernel, dim3(grid_for(total)), dim3(kThreads), 0, (cudaStream_t)stream, dy, dx, H, W, scale, total));
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
        if (Wo % 4 == 0 && ((uintptr_t)y & 15) == 0)
            NB200_CUDA_TRY(launch_kernel(constant_pad2d_kernel<4>, dim3(grid_for(total / 4)), dim3(kThreads), 0, (cudaStream_t)stream, x, y, H, W, left, top, Ho, Wo, value, total / 4));
        else
            NB200_CUDA_TRY(launch_kernel(constant_pad2d_kernel<1>, dim3(grid_for(total)), dim3(kThreads), 0, (cudaStream_t)stream, x, y, H, W, left, top, Ho, Wo, value, total));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }
}
