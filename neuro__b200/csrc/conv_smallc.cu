// HBM-bound first-layer kernels: 3x3, stride 1, NCHW, C <= 4 input channels (VGG block1_conv1: 3 -> 64 at 512x512,
// 12.9 FLOP/B). Tensor cores have nothing to add at this arithmetic intensity, so these are
// CUDA-core fp32 kernels organised around the only thing that matters here: touch the K-channel tensor (y or dy)
// exactly once with full 128-byte lines, and keep the tiny x / filter tensors in registers, L1 or shared memory.
//
//   forward          thread = 4 consecutive output pixels; the 3x3xC input window (C*18 floats) lives in registers and is
//                    reused for every filter; filters come from shared memory as broadcast LDS.128; y is written with
//                    coalesced 16-byte stores (one 512-byte run per warp and filter).
//   input gradient   thread = 4 consecutive dx pixels x C channels; streams dy once (aligned float4 + two edge scalars per
//                    row), flipped filters from shared memory.
//   kernel gradient  warp = one filter k, lanes = pixels: streams dy[k] once with coalesced float4 loads, x windows come
//                    from L1 (shared by the 8 warps of the block), 9*C running sums per lane, one shuffle reduction at the
//                    end; deterministic split over output rows + fixed-order second pass.
#include <stdlib.h>

#include "common.cuh"

namespace nb200
{
    namespace
    {
        constexpr int kSmallThreads = 256;

        struct SmallGeo
        {
            int N, H, W, K, Ho, Wo, pad;
            int aligned; // every tensor base is 16-byte aligned (vector paths allowed)
        };

        // ------------------------------------------------------------ forward
        // ACT >= 0: activation known at compile time (identity / ReLU: the epilogues the reference emits for its fused layers); ACT < 0:
        // run-time switch. With the switch inside the filter loop the loop body was 454 SASS instructions around 108 FFMA.
        template <int ACT>
        __device__ __forceinline__ float activate(int act, float alpha, float v)
        {
            if (ACT == NB200_ACT_IDENTITY) return v;
            if (ACT == NB200_ACT_RELU) return v > 0.f ? v : 0.f;
            return apply_activation(act, alpha, v);
        }

        template <int C, int ACT>
        __global__ void __launch_bounds__(kSmallThreads)
        smallc_fprop_kernel(SmallGeo g, const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                            int act, float alpha, float* __restrict__ y)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            constexpr int T = C * 9;              // taps per filter
            constexpr int TP = (T + 3) & ~3;      // padded to a multiple of 4 for LDS.128
            extern __shared__ float sw[];         // [K][TP]
            for (int i = threadIdx.x; i < g.K * TP; i += kSmallThreads)
            {
                const int k = i / TP, t = i - k * TP;
                sw[i] = t < T ? w[k * T + t] : 0.f;
            }
            __syncthreads();

            const int quadsPerRow = (g.Wo + 3) >> 2;
            const long long quads = (long long)g.N * g.Ho * quadsPerRow;
            const long long q = (long long)blockIdx.x * kSmallThreads + threadIdx.x;
            if (q >= quads)
                return;
            const int qw = (int)(q % quadsPerRow);
            const int oh = (int)((q / quadsPerRow) % g.Ho);
            const int n = (int)(q / ((long long)quadsPerRow * g.Ho));
            const int ow0 = qw * 4;

            // input window: C x 3 rows x 6 columns (ow0-pad .. ow0-pad+5), zero padded
            float xin[C][3][6];
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int r = 0; r < 3; ++r)
                {
                    const int ih = oh - g.pad + r;
                    const float* row = x + (((long long)n * C + c) * g.H + ih) * g.W;
#pragma unroll
                    for (int j = 0; j < 6; ++j)
                    {
                        const int iw = ow0 - g.pad + j;
                        xin[c][r][j] = (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) ? __ldg(row + iw) : 0.f;
                    }
                }

            const long long plane = (long long)g.Ho * g.Wo;
            float* yp = y + (long long)n * g.K * plane + (long long)oh * g.Wo + ow0;
            const bool vec = g.aligned && (g.Wo & 3) == 0; // then every quad is complete and 16-byte aligned
            for (int k = 0; k < g.K; ++k)
            {
                float wk[TP];
                const float4* wp = (const float4*)(sw + k * TP);
#pragma unroll
                for (int i = 0; i < TP / 4; ++i)
                {
                    const float4 f = wp[i];
                    wk[4 * i] = f.x; wk[4 * i + 1] = f.y; wk[4 * i + 2] = f.z; wk[4 * i + 3] = f.w;
                }
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int s = 0; s < 3; ++s)
                        {
                            const float wv = wk[(c * 3 + r) * 3 + s];
                            a0 = fmaf(xin[c][r][s], wv, a0);
                            a1 = fmaf(xin[c][r][s + 1], wv, a1);
                            a2 = fmaf(xin[c][r][s + 2], wv, a2);
                            a3 = fmaf(xin[c][r][s + 3], wv, a3);
                        }
                const float b = bias ? __ldg(bias + k) : 0.f;
                a0 = activate<ACT>(act, alpha, a0 + b); a1 = activate<ACT>(act, alpha, a1 + b);
                a2 = activate<ACT>(act, alpha, a2 + b); a3 = activate<ACT>(act, alpha, a3 + b);
                float* dst = yp + k * plane;
                if (vec)
                    __stcs((float4*)dst, make_float4(a0, a1, a2, a3)); // streaming store: y is not re-read by this kernel
                else
                {
                    if (ow0 < g.Wo) dst[0] = a0;
                    if (ow0 + 1 < g.Wo) dst[1] = a1;
                    if (ow0 + 2 < g.Wo) dst[2] = a2;
                    if (ow0 + 3 < g.Wo) dst[3] = a3;
                }
            }
        }

        // ------------------------------------------------------------ input gradient
        // dx[n][c][h][w] = sum_k sum_{r,s} dy[n][k][h+pad-r][w+pad-s] * w[k][c][r][s]
        template <int C>
        __global__ void __launch_bounds__(kSmallThreads)
        smallc_dgrad_kernel(SmallGeo g, const float* __restrict__ dy, const float* __restrict__ w, const float* __restrict__ bias, int act,
                            float alpha, float* __restrict__ dx)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            constexpr int T = C * 9;
            constexpr int TP = (T + 3) & ~3;
            extern __shared__ float sw[];         // [K][TP], taps flipped: sw[k][(c*3+a)*3+b] = w[k][c][2-a][2-b]
            for (int i = threadIdx.x; i < g.K * TP; i += kSmallThreads)
            {
                const int k = i / TP, t = i - k * TP;
                float v = 0.f;
                if (t < T)
                {
                    const int c = t / 9, a = (t % 9) / 3, b = t % 3;
                    v = w[(k * C + c) * 9 + (2 - a) * 3 + (2 - b)];
                }
                sw[i] = v;
            }
            __syncthreads();

            const int quadsPerRow = (g.W + 3) >> 2;
            const long long quads = (long long)g.N * g.H * quadsPerRow;
            const long long q = (long long)blockIdx.x * kSmallThreads + threadIdx.x;
            if (q >= quads)
                return;
            const int qw = (int)(q % quadsPerRow);
            const int h = (int)((q / quadsPerRow) % g.H);
            const int n = (int)(q / ((long long)quadsPerRow * g.H));
            const int w0 = qw * 4;
            // with flipped taps this is a forward conv of dy with pad' = 2 - pad: window rows h-pad'+a, cols w0-pad'+j
            const int pp = 2 - g.pad;

            float acc[C][4];
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[c][j] = 0.f;

            const long long plane = (long long)g.Ho * g.Wo;
            const float* dyn = dy + (long long)n * g.K * plane;
            const bool fast = g.aligned && pp == 1 && (g.Wo & 3) == 0 && w0 + 4 <= g.Wo; // aligned float4 centre + 2 edge scalars
            // the multiply-accumulate of one filter's 3x6 window into the C x 4 sums
            auto fma_window = [&](const float (&win)[3][6], int k)
            {
                float wk[TP];
                const float4* wp = (const float4*)(sw + k * TP);
#pragma unroll
                for (int i = 0; i < TP / 4; ++i)
                {
                    const float4 f = wp[i];
                    wk[4 * i] = f.x; wk[4 * i + 1] = f.y; wk[4 * i + 2] = f.z; wk[4 * i + 3] = f.w;
                }
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int b = 0; b < 3; ++b)
                        {
                            const float wv = wk[(c * 3 + a) * 3 + b];
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                acc[c][j] = fmaf(win[a][b + j], wv, acc[c][j]);
                        }
            };
            if (fast)
            {
                // Everything that does not depend on the filter is computed once: three row pointers that advance by one plane per
                // filter and five predicates. (The first version recomputed rows, bounds and 64-bit addresses inside the loop: 678
                // SASS instructions around 108 FFMA, which is what made this HBM-sized kernel issue-bound at 1.3-1.9 TB/s.)
                const float* rp[3];
                bool rowOk[3];
#pragma unroll
                for (int a = 0; a < 3; ++a)
                {
                    const int oh = h - 1 + a;
                    rowOk[a] = oh >= 0 && oh < g.Ho;
                    rp[a] = dyn + (long long)(rowOk[a] ? oh : 0) * g.Wo + w0;
                }
                const bool leftOk = w0 > 0, rightOk = w0 + 4 < g.Wo;
                for (int k = 0; k < g.K; ++k)
                {
                    float win[3][6];
#pragma unroll
                    for (int a = 0; a < 3; ++a)
                    {
                        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                        float l = 0.f, rr = 0.f;
                        if (rowOk[a])
                        {
                            f = __ldcs((const float4*)rp[a]);
                            if (leftOk) l = __ldg(rp[a] - 1);
                            if (rightOk) rr = __ldg(rp[a] + 4);
                        }
                        win[a][0] = l; win[a][1] = f.x; win[a][2] = f.y; win[a][3] = f.z; win[a][4] = f.w; win[a][5] = rr;
                        rp[a] += plane;
                    }
                    fma_window(win, k);
                }
            }
            else
            {
                for (int k = 0; k < g.K; ++k)
                {
                    float win[3][6];
#pragma unroll
                    for (int a = 0; a < 3; ++a)
                    {
                        const int oh = h - pp + a;
                        const bool rowOk = oh >= 0 && oh < g.Ho;
                        const float* row = dyn + k * plane + (long long)oh * g.Wo;
#pragma unroll
                        for (int j = 0; j < 6; ++j)
                        {
                            const int ow = w0 - pp + j;
                            win[a][j] = (rowOk && ow >= 0 && ow < g.Wo) ? __ldg(row + ow) : 0.f;
                        }
                    }
                    fma_window(win, k);
                }
            }

            if (bias != nullptr || act != NB200_ACT_IDENTITY)
            {
                // only when this kernel serves as the FORWARD of a few-filter layer (roles swapped, see smallk_* below)
#pragma unroll
                for (int c = 0; c < C; ++c)
                {
                    const float b = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        acc[c][j] = apply_activation(act, alpha, acc[c][j] + b);
                }
            }
            const bool vec = g.aligned && (g.W & 3) == 0;
#pragma unroll
            for (int c = 0; c < C; ++c)
            {
                float* dst = dx + (((long long)n * C + c) * g.H + h) * g.W + w0;
                if (vec)
                    *(float4*)dst = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
                else
                {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (w0 + j < g.W) dst[j] = acc[c][j];
                }
            }
        }

        // ------------------------------------------------------------ input gradient, two dx rows per thread
        // Same arithmetic for the common geometry (pad 1, aligned, W % 4 == 0, even H): a thread owns the 4-pixel quads of TWO
        // consecutive dx rows, so the 4 dy rows it loads per filter serve 216 FMAs instead of 3 rows serving 108 -- a third less
        // L2->L1 traffic per output (every dy row used to be fetched by the blocks of three dx rows).
        template <int C>
        __global__ void __launch_bounds__(kSmallThreads)
        smallc_dgrad2_kernel(SmallGeo g, const float* __restrict__ dy, const float* __restrict__ w, const float* __restrict__ bias, int act,
                             float alpha, float* __restrict__ dx)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            constexpr int T = C * 9;
            constexpr int TP = (T + 3) & ~3;
            extern __shared__ float sw[];         // [K][TP], taps flipped: sw[k][(c*3+a)*3+b] = w[k][c][2-a][2-b]
            for (int i = threadIdx.x; i < g.K * TP; i += kSmallThreads)
            {
                const int k = i / TP, t = i - k * TP;
                float v = 0.f;
                if (t < T)
                {
                    const int c = t / 9, a = (t % 9) / 3, b = t % 3;
                    v = w[(k * C + c) * 9 + (2 - a) * 3 + (2 - b)];
                }
                sw[i] = v;
            }
            __syncthreads();

            const int quadsPerRow = g.W >> 2, H2 = g.H >> 1;
            const long long quads = (long long)g.N * H2 * quadsPerRow;
            const long long q = (long long)blockIdx.x * kSmallThreads + threadIdx.x;
            if (q >= quads)
                return;
            const int qw = (int)(q % quadsPerRow);
            const int h = 2 * (int)((q / quadsPerRow) % H2);
            const int n = (int)(q / ((long long)quadsPerRow * H2));
            const int w0 = qw * 4;

            float acc[2][C][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        acc[i][c][j] = 0.f;

            const long long plane = (long long)g.Ho * g.Wo;
            const float* rp[4];
            bool rowOk[4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
            {
                const int oh = h - 1 + r;
                rowOk[r] = oh >= 0 && oh < g.Ho;
                rp[r] = dy + (long long)n * g.K * plane + (long long)(rowOk[r] ? oh : 0) * g.Wo + w0;
            }
            const bool leftOk = w0 > 0, rightOk = w0 + 4 < g.Wo;
            for (int k = 0; k < g.K; ++k)
            {
                float win[4][6];
#pragma unroll
                for (int r = 0; r < 4; ++r)
                {
                    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                    float l = 0.f, rr = 0.f;
                    if (rowOk[r])
                    {
                        f = __ldcs((const float4*)rp[r]);
                        if (leftOk) l = __ldg(rp[r] - 1);
                        if (rightOk) rr = __ldg(rp[r] + 4);
                    }
                    win[r][0] = l; win[r][1] = f.x; win[r][2] = f.y; win[r][3] = f.z; win[r][4] = f.w; win[r][5] = rr;
                    rp[r] += plane;
                }
                float wk[TP];
                const float4* wp = (const float4*)(sw + k * TP);
#pragma unroll
                for (int i = 0; i < TP / 4; ++i)
                {
                    const float4 f = wp[i];
                    wk[4 * i] = f.x; wk[4 * i + 1] = f.y; wk[4 * i + 2] = f.z; wk[4 * i + 3] = f.w;
                }
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int b = 0; b < 3; ++b)
                        {
                            const float wv = wk[(c * 3 + a) * 3 + b];
#pragma unroll
                            for (int i = 0; i < 2; ++i)
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    acc[i][c][j] = fmaf(win[i + a][b + j], wv, acc[i][c][j]);
                        }
            }

            if (bias != nullptr || act != NB200_ACT_IDENTITY)
            {
#pragma unroll
                for (int c = 0; c < C; ++c)
                {
                    const float b = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            acc[i][c][j] = apply_activation(act, alpha, acc[i][c][j] + b);
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int c = 0; c < C; ++c)
                    *(float4*)(dx + (((long long)n * C + c) * g.H + h + i) * g.W + w0) = make_float4(acc[i][c][0], acc[i][c][1], acc[i][c][2], acc[i][c][3]);
        }

        // ------------------------------------------------------------ input gradient, filters split over the warps of a block
        // Same arithmetic as smallc_dgrad_kernel for grids that would leave most of the chip idle (few pixels, many filters:
        // the forward of a few-filter output layer, e.g. DCGAN's 128 -> 3 at 32x32): all 8 warps of a block work on the SAME 32
        // pixel quads, warp w reducing filters [w*K/8, (w+1)*K/8); the eight partial sums meet in shared memory and are added in
        // warp order (deterministic). 8x the blocks, 1/8 of the serial filter loop per thread.
        template <int C>
        __global__ void __launch_bounds__(kSmallThreads)
        smallc_dgrad_ksplit_kernel(SmallGeo g, const float* __restrict__ dy, const float* __restrict__ w, const float* __restrict__ bias, int act,
                            float alpha, float* __restrict__ dx)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            constexpr int T = C * 9;
            constexpr int TP = (T + 3) & ~3;
            extern __shared__ float sw[];         // [K][TP], taps flipped: sw[k][(c*3+a)*3+b] = w[k][c][2-a][2-b]
            for (int i = threadIdx.x; i < g.K * TP; i += kSmallThreads)
            {
                const int k = i / TP, t = i - k * TP;
                float v = 0.f;
                if (t < T)
                {
                    const int c = t / 9, a = (t % 9) / 3, b = t % 3;
                    v = w[(k * C + c) * 9 + (2 - a) * 3 + (2 - b)];
                }
                sw[i] = v;
            }
            __syncthreads();

            const int quadsPerRow = (g.W + 3) >> 2;
            const long long quads = (long long)g.N * g.H * quadsPerRow;
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            const long long qRaw = (long long)blockIdx.x * 32 + lane;
            const bool live = qRaw < quads;
            const long long q = live ? qRaw : quads - 1;   // idle lanes shadow the last quad (no early return: block-wide barriers below)
            const int qw = (int)(q % quadsPerRow);
            const int h = (int)((q / quadsPerRow) % g.H);
            const int n = (int)(q / ((long long)quadsPerRow * g.H));
            const int kPer = (g.K + 7) >> 3;
            const int kBegin = warp * kPer, kEnd = min(g.K, kBegin + kPer);
            const int w0 = qw * 4;
            // with flipped taps this is a forward conv of dy with pad' = 2 - pad: window rows h-pad'+a, cols w0-pad'+j
            const int pp = 2 - g.pad;

            float acc[C][4];
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[c][j] = 0.f;

            const long long plane = (long long)g.Ho * g.Wo;
            const float* dyn = dy + (long long)n * g.K * plane;
            const bool fast = g.aligned && pp == 1 && (g.Wo & 3) == 0 && w0 + 4 <= g.Wo; // aligned float4 centre + 2 edge scalars
            for (int k = kBegin; k < kEnd; ++k)
            {
                float win[3][6];
#pragma unroll
                for (int a = 0; a < 3; ++a)
                {
                    const int oh = h - pp + a;
                    const bool rowOk = oh >= 0 && oh < g.Ho;
                    const float* row = dyn + k * plane + (long long)oh * g.Wo;
                    if (fast)
                    {
                        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                        float l = 0.f, rr = 0.f;
                        if (rowOk)
                        {
                            f = __ldcs((const float4*)(row + w0));
                            if (w0 > 0) l = __ldg(row + w0 - 1);
                            if (w0 + 4 < g.Wo) rr = __ldg(row + w0 + 4);
                        }
                        win[a][0] = l; win[a][1] = f.x; win[a][2] = f.y; win[a][3] = f.z; win[a][4] = f.w; win[a][5] = rr;
                    }
                    else
                    {
#pragma unroll
                        for (int j = 0; j < 6; ++j)
                        {
                            const int ow = w0 - pp + j;
                            win[a][j] = (rowOk && ow >= 0 && ow < g.Wo) ? __ldg(row + ow) : 0.f;
                        }
                    }
                }
                float wk[TP];
                const float4* wp = (const float4*)(sw + k * TP);
#pragma unroll
                for (int i = 0; i < TP / 4; ++i)
                {
                    const float4 f = wp[i];
                    wk[4 * i] = f.x; wk[4 * i + 1] = f.y; wk[4 * i + 2] = f.z; wk[4 * i + 3] = f.w;
                }
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int b = 0; b < 3; ++b)
                        {
                            const float wv = wk[(c * 3 + a) * 3 + b];
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                acc[c][j] = fmaf(win[a][b + j], wv, acc[c][j]);
                        }
            }

            // partial sums of the 8 filter slices -> warp 0 adds them in slice order
            float* red = sw + g.K * TP;               // [8 warps][C * 4][32 lanes]
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    red[(warp * C * 4 + c * 4 + j) * 32 + lane] = acc[c][j];
            __syncthreads();
            if (warp != 0 || !live)
                return;
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    float v = 0.f;
#pragma unroll
                    for (int wv = 0; wv < 8; ++wv)
                        v += red[(wv * C * 4 + c * 4 + j) * 32 + lane];
                    acc[c][j] = v;
                }
            if (bias != nullptr || act != NB200_ACT_IDENTITY)
            {
                // only when this kernel serves as the FORWARD of a few-filter layer (roles swapped, see smallk_* below)
#pragma unroll
                for (int c = 0; c < C; ++c)
                {
                    const float b = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        acc[c][j] = apply_activation(act, alpha, acc[c][j] + b);
                }
            }
            const bool vec = g.aligned && (g.W & 3) == 0;
#pragma unroll
            for (int c = 0; c < C; ++c)
            {
                float* dst = dx + (((long long)n * C + c) * g.H + h) * g.W + w0;
                if (vec)
                    *(float4*)dst = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
                else
                {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (w0 + j < g.W) dst[j] = acc[c][j];
                }
            }
        }

        // ------------------------------------------------------------ kernel gradient
        // dw[k][c][r][s] = sum_{n,oh,ow} dy[n][k][oh][ow] * x[n][c][oh+r-pad][ow+s-pad]
        // grid (ceil(K/8), slices): warp wIdx of the block owns filter k = blockIdx.x*8 + wIdx and the slice's output rows.
        template <int C>
        __global__ void __launch_bounds__(kSmallThreads)
        smallc_wgrad_kernel(SmallGeo g, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part,
                            int rowsPerSlice)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            constexpr int T = C * 9;
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            const int k = blockIdx.x * 8 + warp;
            const int slice = blockIdx.y;
            const int rows = g.N * g.Ho;
            const int rowBegin = slice * rowsPerSlice;
            const int rowEnd = min(rowBegin + rowsPerSlice, rows);

            float acc[T];
#pragma unroll
            for (int t = 0; t < T; ++t)
                acc[t] = 0.f;

            if (k < g.K)
            {
                const long long plane = (long long)g.Ho * g.Wo;
                const bool vec = g.aligned && (g.Wo & 3) == 0;
                const bool xfast = g.aligned && g.pad == 1 && (g.W & 3) == 0;
                for (int row = rowBegin; row < rowEnd; ++row)
                {
                    const int n = row / g.Ho, oh = row - n * g.Ho;
                    const float* drow = dy + ((long long)n * g.K + k) * plane + (long long)oh * g.Wo;
                    for (int ow0 = lane * 4; ow0 < g.Wo; ow0 += 128)
                    {
                        float d[4];
                        if (vec)
                        {
                            const float4 f = __ldcs((const float4*)(drow + ow0));
                            d[0] = f.x; d[1] = f.y; d[2] = f.z; d[3] = f.w;
                        }
                        else
                        {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                d[j] = ow0 + j < g.Wo ? __ldg(drow + ow0 + j) : 0.f;
                        }
#pragma unroll
                        for (int c = 0; c < C; ++c)
#pragma unroll
                            for (int r = 0; r < 3; ++r)
                            {
                                const int ih = oh - g.pad + r;
                                if (ih < 0 || ih >= g.H)
                                    continue; // warp-uniform
                                const float* xr = x + (((long long)n * C + c) * g.H + ih) * g.W;
                                float xv[6];
                                if (xfast && ow0 + 4 <= g.W)
                                {
                                    const float4 f = __ldg((const float4*)(xr + ow0));
                                    xv[0] = ow0 > 0 ? __ldg(xr + ow0 - 1) : 0.f;
                                    xv[1] = f.x; xv[2] = f.y; xv[3] = f.z; xv[4] = f.w;
                                    xv[5] = ow0 + 4 < g.W ? __ldg(xr + ow0 + 4) : 0.f;
                                }
                                else
                                {
#pragma unroll
                                    for (int j = 0; j < 6; ++j)
                                    {
                                        const int iw = ow0 - g.pad + j;
                                        xv[j] = (iw >= 0 && iw < g.W) ? __ldg(xr + iw) : 0.f;
                                    }
                                }
#pragma unroll
                                for (int s = 0; s < 3; ++s)
                                {
                                    float a = acc[(c * 3 + r) * 3 + s];
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        a = fmaf(d[j], xv[s + j], a);
                                    acc[(c * 3 + r) * 3 + s] = a;
                                }
                            }
                    }
                }
            }

#pragma unroll
            for (int t = 0; t < T; ++t)
            {
                float v = acc[t];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && k < g.K)
                    part[((long long)slice * g.K + k) * T + t] = v;
            }
        }

        __global__ void smallc_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int count, int slices)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const int i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= count)
                return;
            float v = 0.f;
            for (int s = 0; s < slices; ++s)
                v += part[(long long)s * count + i];
            out[i] = v;
        }

        // ------------------------------------------------------------ strided kernel gradient, few channels
        // dw[k][c][r][s] = sum_{n,oh,ow} dy[n][k][oh][ow] * x[n][c][oh*st+r-padY][ow*st+s-padX]   (first layers of the GAN
        // discriminators / U-Net encoder: 3 or 6 channels, 3x3 or 4x4 filters, stride 2 -- HBM-bound, and 10-60x off their HBM
        // time on the gathered tensor-core kernel, whose 128-channel A tile holds 3 real channels). Warp = KPW filters, lanes =
        // output pixels of a row; the C*F*F*KPW running sums live in registers, x comes from L1 (the 8 warps of a block read the
        // same rows), dy is streamed once with coalesced loads. Deterministic split over output rows + fixed-order second pass.
        struct StridedGeo
        {
            int N, H, W, K, Ho, Wo, stride, padX, padY;
        };

        template <int C, int F, int KPW>
        __global__ void __launch_bounds__(kSmallThreads)
        strided_wgrad_kernel(StridedGeo g, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part, int pixPerSlice)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            constexpr int T = C * F * F;
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            const int kBase = (blockIdx.x * 8 + warp) * KPW;
            const int slice = blockIdx.y;
            const unsigned plane = (unsigned)(g.Ho * g.Wo);
            const unsigned total = (unsigned)g.N * plane;               // output pixels, flattened (n, oh, ow): lanes stay busy on narrow maps
            const unsigned pBegin = (unsigned)slice * (unsigned)pixPerSlice;
            const unsigned pEnd = min(pBegin + (unsigned)pixPerSlice, total);

            float acc[KPW][T];
#pragma unroll
            for (int j = 0; j < KPW; ++j)
#pragma unroll
                for (int t = 0; t < T; ++t)
                    acc[j][t] = 0.f;

            if (kBase < g.K)
            {
                for (unsigned p = pBegin + lane; p < pEnd; p += 32)
                {
                    const unsigned n = p / plane, rem = p - n * plane;
                    const int oh = (int)(rem / (unsigned)g.Wo), ow = (int)(rem - (unsigned)oh * (unsigned)g.Wo);
                    float d[KPW];
#pragma unroll
                    for (int j = 0; j < KPW; ++j)
                        d[j] = kBase + j < g.K ? __ldcs(dy + ((long long)n * g.K + kBase + j) * plane + rem) : 0.f;
                    const int ih0 = oh * g.stride - g.padY, iw0 = ow * g.stride - g.padX;
#pragma unroll
                    for (int c = 0; c < C; ++c)
#pragma unroll
                        for (int r = 0; r < F; ++r)
                        {
                            const int ih = ih0 + r;
                            const bool rowOk = ih >= 0 && ih < g.H;
                            const float* xr = x + (((long long)n * C + c) * g.H + (rowOk ? ih : 0)) * g.W;
#pragma unroll
                            for (int s2 = 0; s2 < F; ++s2)
                            {
                                const int iw = iw0 + s2;
                                const float xv = (rowOk && iw >= 0 && iw < g.W) ? __ldg(xr + iw) : 0.f;
#pragma unroll
                                for (int j = 0; j < KPW; ++j)
                                    acc[j][(c * F + r) * F + s2] = fmaf(d[j], xv, acc[j][(c * F + r) * F + s2]);
                            }
                        }
                }
            }

#pragma unroll
            for (int j = 0; j < KPW; ++j)
#pragma unroll
                for (int t = 0; t < T; ++t)
                {
                    float v = acc[j][t];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
                        v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0 && kBase + j < g.K)
                        part[((long long)slice * g.K + kBase + j) * T + t] = v;
                }
        }

        constexpr int strided_kpw(int C, int F) { return C * F * F <= 27 ? 4 : C * F * F <= 54 ? 2 : 1; }

        int strided_slices(const nb200_conv_desc& d)
        {
            const int kpw = strided_kpw(d.C, d.R);
            const int kBlocks = ceil_div(d.K, 8 * kpw);
            const long long chunks = ((long long)d.N * d.Ho * d.Wo + 31) / 32; // a slice is a whole number of 32-pixel steps
            long long want = ceil_div(148 * 4, kBlocks);
            if (want > chunks) want = chunks;
            return want < 1 ? 1 : (int)want;
        }

        // out[b][a][2-r][2-s] = in[a][b][r][s]: filters transposed and rotated by 180 degrees (3x3 taps: index t -> 8 - t)
        __global__ void swap_filters_kernel(const float* __restrict__ in, float* __restrict__ out, int A, int B)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const int i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= A * B * 9)
                return;
            const int t = i % 9, ab = i / 9, b = ab % B, a = ab / B;
            out[(b * A + a) * 9 + (8 - t)] = in[i];
        }

        SmallGeo small_geo(const nb200_conv_desc& d, const void* a, const void* b, const void* c)
        {
            const int aligned = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
            return SmallGeo{d.N, d.H, d.W, d.K, d.Ho, d.Wo, d.padX, aligned};
        }

        int wgrad_slices(const nb200_conv_desc& d)
        {
            const int rows = d.N * d.Ho;
            const int kBlocks = ceil_div(d.K, 8);
            int want = ceil_div(148 * 6, kBlocks);
            if (want > rows) want = rows;
            if (want < 1) want = 1;
            return want;
        }
    }

    bool smallc_supported(const nb200_conv_desc& d)
    {
        return d.fmt == NB200_NCHW && d.C >= 1 && d.C <= 4 && d.R == 3 && d.S == 3 && d.stride == 1 && d.padX == d.padY && d.padX <= 2 &&
               d.K >= 1 && d.K * ((d.C * 9 + 3) & ~3) * 4 <= 96 * 1024 &&
               d.Ho == d.H + 2 * d.padY - 2 && d.Wo == d.W + 2 * d.padX - 2 && d.N >= 1 && d.Ho >= 1 && d.Wo >= 1;
    }

    size_t smallc_wgrad_workspace(const nb200_conv_desc& d)
    {
        return (size_t)wgrad_slices(d) * d.K * d.C * 9 * sizeof(float);
    }

#define SMALLC_DISPATCH(CALL)          \
    switch (d.C)                       \
    {                                  \
    case 1: { CALL(1); break; }        \
    case 2: { CALL(2); break; }        \
    case 3: { CALL(3); break; }        \
    default: { CALL(4); break; }       \
    }

    int smallc_forward(const nb200_conv_desc& d, const float* x, const float* w, const float* bias, int act, float alpha, float* y,
                       cudaStream_t st)
    {
        const SmallGeo g = small_geo(d, x, y, nullptr);
        const long long quads = (long long)d.N * d.Ho * ((d.Wo + 3) / 4);
        const size_t smem = (size_t)d.K * ((d.C * 9 + 3) & ~3) * 4;
#define CALL_ACT(CC, AA)                                                                                                      \
        {                                                                                                                      \
            if (smem > 48 * 1024)                                                                                              \
                NB200_CUDA_TRY(cudaFuncSetAttribute(smallc_fprop_kernel<CC, AA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            NB200_CUDA_TRY(launch_kernel(smallc_fprop_kernel<CC, AA>, dim3(ceil_div(quads, kSmallThreads)), dim3(kSmallThreads), smem, st, g, x, w, bias, act, alpha, y)); \
        }
#define CALL(CC)                                                                                                              \
        if (act == NB200_ACT_IDENTITY) CALL_ACT(CC, NB200_ACT_IDENTITY)                                                        \
        else if (act == NB200_ACT_RELU) CALL_ACT(CC, NB200_ACT_RELU)                                                           \
        else CALL_ACT(CC, -1)
        SMALLC_DISPATCH(CALL)
#undef CALL
#undef CALL_ACT
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int smallc_input_gradient(const nb200_conv_desc& d, const float* dy, const float* w, float* dx, cudaStream_t st)
    {
        return smallc_input_gradient_epilogue(d, dy, w, nullptr, NB200_ACT_IDENTITY, 0.f, dx, st);
    }

    int smallc_input_gradient_epilogue(const nb200_conv_desc& d, const float* dy, const float* w, const float* bias, int act, float alpha,
                                       float* dx, cudaStream_t st)
    {
        const SmallGeo g = small_geo(d, dy, dx, nullptr);
        const long long quads = (long long)d.N * d.H * ((d.W + 3) / 4);
        const size_t smem = (size_t)d.K * ((d.C * 9 + 3) & ~3) * 4;
        {
            // two dx rows per thread where the geometry allows it and the halved grid still fills the chip
            static const char* env2 = getenv("NB200_SMALLC_DGRAD2"); // 0 disables (profiling)
            const long long quads2 = (long long)d.N * (d.H / 2) * (d.W / 4);
            if (g.aligned && d.padX == 1 && (d.W & 3) == 0 && d.W == d.Wo && d.H == d.Ho && (d.H & 1) == 0 && ceil_div(quads2, kSmallThreads) >= 148 &&
                !(env2 && env2[0] == '0'))
            {
#define CALL(CC)                                                                                                              \
                if (smem > 48 * 1024)                                                                                          \
                    NB200_CUDA_TRY(cudaFuncSetAttribute(smallc_dgrad2_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                NB200_CUDA_TRY(launch_kernel(smallc_dgrad2_kernel<CC>, dim3(ceil_div(quads2, kSmallThreads)), dim3(kSmallThreads), smem, st, g, dy, w, bias, act, alpha, dx));
                SMALLC_DISPATCH(CALL)
#undef CALL
                NB200_CUDA_TRY(cudaGetLastError());
                count_launch();
                return NB200_OK;
            }
        }
        // less than one block per SM and enough filters to share out: split the filters over the warps of a block instead
        // (DCGAN 128 -> 3 @32x32 batch 128, 128 blocks: 0.089 -> 0.070 ms; at 256 blocks -- VGG 64 -> 3 @512x512 batch 1 -- the
        // split is slower, 0.067 vs 0.050 ms, because every block reloads the filters)
        static const char* env = getenv("NB200_SMALLC_KSPLIT"); // 0 disables (profiling)
        if (ceil_div(quads, kSmallThreads) < 148 && d.K >= 32 && !(env && env[0] == '0'))
        {
            const size_t smem2 = smem + (size_t)8 * d.C * 4 * 32 * sizeof(float);
#define CALL(CC)                                                                                                              \
            if (smem2 > 48 * 1024)                                                                                              \
                NB200_CUDA_TRY(cudaFuncSetAttribute(smallc_dgrad_ksplit_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
            NB200_CUDA_TRY(launch_kernel(smallc_dgrad_ksplit_kernel<CC>, dim3(ceil_div(quads, 32)), dim3(kSmallThreads), smem2, st, g, dy, w, bias, act, alpha, dx));
            SMALLC_DISPATCH(CALL)
#undef CALL
            NB200_CUDA_TRY(cudaGetLastError());
            count_launch();
            return NB200_OK;
        }
#define CALL(CC)                                                                                                              \
        if (smem > 48 * 1024)                                                                                                  \
            NB200_CUDA_TRY(cudaFuncSetAttribute(smallc_dgrad_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        NB200_CUDA_TRY(launch_kernel(smallc_dgrad_kernel<CC>, dim3(ceil_div(quads, kSmallThreads)), dim3(kSmallThreads), smem, st, g, dy, w, bias, act, alpha, dx));
        SMALLC_DISPATCH(CALL)
#undef CALL
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int smallc_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes,
                                cudaStream_t st)
    {
        const SmallGeo g = small_geo(d, x, dy, nullptr);
        const int slices = wgrad_slices(d);
        const size_t need = smallc_wgrad_workspace(d);
        if (wsBytes < need || !ws)
            return fail(NB200_E_WORKSPACE, "small-channel kernel gradient needs %zu workspace bytes, got %zu", need, wsBytes);
        const int rowsPerSlice = ceil_div(d.N * d.Ho, slices);
        dim3 grid(ceil_div(d.K, 8), slices);
#define CALL(CC) NB200_CUDA_TRY(launch_kernel(smallc_wgrad_kernel<CC>, dim3(grid), dim3(kSmallThreads), 0, st, g, x, dy, (float*)ws, rowsPerSlice));
        SMALLC_DISPATCH(CALL)
#undef CALL
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        const int count = d.K * d.C * 9;
        NB200_CUDA_TRY(launch_kernel(smallc_reduce_kernel, dim3(ceil_div(count, 256)), dim3(256), 0, st, (const float*)ws, dw, count, slices));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }
    // ---- few-FILTER layers (K <= 4: the RGB / mask output convolutions of the GAN generators and the autoencoder) ----
    // A 3x3 stride-1 convolution with many channels and a few filters is the small-channel problem with the roles of its
    // two activation tensors exchanged: with w'[c][k][r][s] = w[k][c][2-r][2-s] and pad' = 2 - pad
    //   forward(x -> y)        = small-channel INPUT GRADIENT of (dy' = x) giving (dx' = y)      [+ bias / activation]
    //   input gradient(dy->dx) = small-channel FORWARD of (x' = dy) giving (y' = dx)
    //   kernel gradient        = small-channel KERNEL GRADIENT with (x' = dy, dy' = x), then dw[k][c][r][s] = dw'[c][k][2-r][2-s]
    // so these HBM-bound layers stream their big tensor once through the kernels above instead of running as a GEMM with
    // 3 of 64 accumulator columns in use.
    nb200_conv_desc smallk_swapped(const nb200_conv_desc& d)
    {
        nb200_conv_desc s = d;
        s.C = d.K; s.K = d.C;
        s.H = d.Ho; s.W = d.Wo; s.Ho = d.H; s.Wo = d.W;
        s.padX = 2 - d.padX; s.padY = 2 - d.padY;
        return s;
    }

    bool smallk_supported(const nb200_conv_desc& d)
    {
        if (!(d.K >= 1 && d.K <= 4 && d.C > 4 && d.R == 3 && d.S == 3 && d.stride == 1 && d.padX <= 2 && d.padY <= 2))
            return false;
        return smallc_supported(smallk_swapped(d));
    }

    int smallc_swap_filters(const float* in, float* out, int A, int B, cudaStream_t st)
    {
        NB200_CUDA_TRY(launch_kernel(swap_filters_kernel, dim3(ceil_div((long long)A * B * 9, 256)), dim3(256), 0, st, in, out, A, B));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }
    // ---- strided few-channel kernel gradient (strided_wgrad_kernel) ----
    bool strided_wgrad_supported(const nb200_conv_desc& d)
    {
        // measured (profiles/README.md): 3 channels x 3x3 (27 sums, 4 filters per warp) halves the gathered kernel's time; 6 channels x 4x4
        // (96 sums, 237 registers, one block per SM) is 2.4x SLOWER than it, so the family stops at 36 running sums per filter
        const bool cOk = d.C >= 1 && d.C <= 4 && d.C * d.R * d.S <= 36;
        return d.fmt == NB200_NCHW && cOk && d.R == d.S && (d.R == 3 || d.R == 4) && d.stride == 2 && d.K >= 1 && d.N >= 1 && d.Ho >= 1 &&
               d.Wo >= 1 && d.H >= 1 && d.W >= 1 && (long long)d.N * d.Ho * d.Wo <= 0x7fffffffll;
    }

    size_t strided_wgrad_workspace(const nb200_conv_desc& d)
    {
        return (size_t)strided_slices(d) * d.K * d.C * d.R * d.S * sizeof(float);
    }

    int strided_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes, cudaStream_t st)
    {
        const size_t need = strided_wgrad_workspace(d);
        if (wsBytes < need || !ws)
            return fail(NB200_E_WORKSPACE, "strided few-channel kernel gradient needs %zu workspace bytes, got %zu", need, wsBytes);
        const StridedGeo g{d.N, d.H, d.W, d.K, d.Ho, d.Wo, d.stride, d.padX, d.padY};
        const int slices = strided_slices(d);
        const long long chunks = ((long long)d.N * d.Ho * d.Wo + 31) / 32;
        const int rowsPerSlice = (int)((chunks + slices - 1) / slices) * 32; // pixels per slice
#define SW_CALL(CC, FF)                                                                                                        \
        {                                                                                                                      \
            constexpr int kpw = strided_kpw(CC, FF);                                                                           \
            dim3 grid(ceil_div(d.K, 8 * kpw), slices);                                                                         \
            NB200_CUDA_TRY(launch_kernel(strided_wgrad_kernel<CC, FF, kpw>, dim3(grid), dim3(kSmallThreads), 0, st, g, x, dy, (float*)ws, rowsPerSlice));            \
        }
        if (d.R == 3)
        {
            switch (d.C) { case 1: SW_CALL(1, 3) break; case 2: SW_CALL(2, 3) break; case 3: SW_CALL(3, 3) break; default: SW_CALL(4, 3) break; }
        }
        else
        {
            switch (d.C) { case 1: SW_CALL(1, 4) break; default: SW_CALL(2, 4) break; }
        }
#undef SW_CALL
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        const int count = d.K * d.C * d.R * d.S;
        NB200_CUDA_TRY(launch_kernel(smallc_reduce_kernel, dim3(ceil_div(count, 256)), dim3(256), 0, st, (const float*)ws, dw, count, slices));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }
}
