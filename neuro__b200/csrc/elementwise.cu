// HBM-bound neighbours of the convolution: bias gradient and the optimiser updates that follow the
// gradient exchange.
#include "common.cuh"

namespace nb200
{
    namespace
    {
        constexpr int kReduceThreads = 1024;

        // db[k] = sum_{n,oh,ow} dy[n,k,oh,ow]   (TensorOpCpu::Conv2DBiasGradient, TensorOpCpu.cpp:1065-1068)
        // One block per channel; fixed reduction tree => deterministic.
        __global__ void __launch_bounds__(kReduceThreads)
        bias_gradient_kernel(const float* __restrict__ dy, float* __restrict__ db, int N, int K, int HW, ActStrides ys, int nchw)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const int k = blockIdx.x;
            float acc = 0.f;
            if (nchw)
            {
                // (n, hw) planes are contiguous runs of HW floats
                for (int n = 0; n < N; ++n)
                {
                    const float* p = dy + n * ys.n + k * ys.c;
                    for (int i = threadIdx.x; i < HW; i += kReduceThreads)
                        acc += __ldg(p + i);
                }
            }
            else
            {
                const long long pixels = (long long)N * HW;
                for (long long i = threadIdx.x; i < pixels; i += kReduceThreads)
                    acc += __ldg(dy + i * K + k);
            }
            __shared__ float red[kReduceThreads / 32];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if ((threadIdx.x & 31) == 0)
                red[threadIdx.x >> 5] = acc;
            __syncthreads();
            if (threadIdx.x < 32)
            {
                float v = threadIdx.x < kReduceThreads / 32 ? red[threadIdx.x] : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    v += __shfl_xor_sync(0xffffffffu, v, o);
                if (threadIdx.x == 0)
                    db[k] = v;
            }
        }

        // Derivative of the activation expressed through its OUTPUT, times the incoming gradient, exactly as the reference
        // evaluates it (TensorOpCpu.cpp:813-864: SigmoidGradient, TanhGradient, ReLUGradient, EluGradient, LeakyReLUGradient).
        template <int ACT>
        __device__ __forceinline__ float activation_gradient(float y, float g, float alpha)
        {
            // explicit round-to-nearest steps: no FMA contraction, so the result is bit-identical to the reference's fp32 loop
            if (ACT == NB200_ACT_SIGMOID) return __fmul_rn(__fmul_rn(y, __fsub_rn(1.f, y)), g);
            if (ACT == NB200_ACT_RELU) return y > 0.f ? g : 0.f;
            if (ACT == NB200_ACT_TANH) return __fmul_rn(__fsub_rn(1.f, __fmul_rn(y, y)), g);
            if (ACT == NB200_ACT_ELU) return __fmul_rn(y > 0.f ? 1.f : __fadd_rn(y, alpha), g);
            if (ACT == NB200_ACT_LEAKY_RELU) return __fmul_rn(y > 0.f ? 1.f : alpha, g);
            return g;
        }

        constexpr int kAgThreads = 256;
        constexpr int kAgSeg = 8192; // elements of one (n, k) plane handled by one block: 8 float4 per thread

        __device__ __forceinline__ float block_sum_256(float acc)
        {
            __shared__ float red[kAgThreads / 32];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if ((threadIdx.x & 31) == 0)
                red[threadIdx.x >> 5] = acc;
            __syncthreads();
            float v = threadIdx.x < kAgThreads / 32 ? red[threadIdx.x] : 0.f;
            if (threadIdx.x < 32)
            {
#pragma unroll
                for (int o = 4; o > 0; o >>= 1)
                    v += __shfl_xor_sync(0xffffffffu, v, o);
            }
            return v; // valid in thread 0
        }

        // Backward prologue of Conv2dBiasActivationOp (Conv2dBiasActivationOp.cpp:47-60) in ONE pass over HBM:
        //   dz = act'(y) * dy                              (Tensor::ActivationGradient)
        //   partial[k][n * segs + seg] = sum of this block's dz   (first half of Conv2DBiasGradient, TensorOpCpu.cpp:1065-1068)
        // NCHW: block (seg, k, n) owns elements [seg*kAgSeg, ...) of plane (n, k): every partial belongs to one channel.
        template <int ACT, bool VEC>
        __global__ void __launch_bounds__(kAgThreads)
        act_bias_gradient_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dz,
                                 float* __restrict__ partial, int HW, int K, int segs, float alpha)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const int seg = blockIdx.x, k = blockIdx.y, n = blockIdx.z;
            const long long base = ((long long)n * K + k) * HW;
            const int lo = seg * kAgSeg;
            const int hi = min(HW, lo + kAgSeg);
            float acc = 0.f;
            if (VEC)
            {
                const float4* y4 = reinterpret_cast<const float4*>(y + base);
                const float4* g4 = reinterpret_cast<const float4*>(dy + base);
                float4* z4 = reinterpret_cast<float4*>(dz + base);
                for (int i = lo / 4 + threadIdx.x; i < hi / 4; i += kAgThreads)
                {
                    const float4 a = __ldcs(y4 + i), g = __ldcs(g4 + i);
                    float4 z;
                    z.x = activation_gradient<ACT>(a.x, g.x, alpha);
                    z.y = activation_gradient<ACT>(a.y, g.y, alpha);
                    z.z = activation_gradient<ACT>(a.z, g.z, alpha);
                    z.w = activation_gradient<ACT>(a.w, g.w, alpha);
                    z4[i] = z;
                    acc += (z.x + z.y) + (z.z + z.w);
                }
            }
            else
            {
                for (int i = lo + threadIdx.x; i < hi; i += kAgThreads)
                {
                    const float z = activation_gradient<ACT>(y[base + i], dy[base + i], alpha);
                    dz[base + i] = z;
                    acc += z;
                }
            }
            if (partial)
            {
                const float v = block_sum_256(acc);
                if (threadIdx.x == 0)
                    partial[(long long)k * gridDim.z * segs + (long long)n * segs + seg] = v;
            }
        }

        // db[k] = partials of channel k added in index order by one warp (fixed order => deterministic)
        __global__ void bias_partial_reduce_kernel(const float* __restrict__ partial, float* __restrict__ db, int K, int per)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const int k = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
            if (k >= K)
                return;
            const float* p = partial + (long long)k * per;
            float acc = 0.f;
            for (int i = threadIdx.x & 31; i < per; i += 32)
                acc += p[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if ((threadIdx.x & 31) == 0)
                db[k] = acc;
        }

        // first-match rule of the reference's max-pool gradient on one 2x2 window (scan order e00, e01, e10, e11;
        // TensorOpCpu.cpp:1249-1338), then the activation gradient through the window's own values (they ARE the activation's
        // output) -- every step rounded as the two separate reference passes round it
        template <int ACT>
        __device__ __forceinline__ float pool_act_window(float e00, float e01, float e10, float e11, float m, float g, float alpha, float& d00,
                                                         float& d01, float& d10, float& d11)
        {
            const bool m0 = e00 == m, m1 = !m0 && e01 == m, m2 = !m0 && !m1 && e10 == m, m3 = !m0 && !m1 && !m2 && e11 == m;
            d00 = activation_gradient<ACT>(e00, m0 ? g : 0.f, alpha); d01 = activation_gradient<ACT>(e01, m1 ? g : 0.f, alpha);
            d10 = activation_gradient<ACT>(e10, m2 ? g : 0.f, alpha); d11 = activation_gradient<ACT>(e11, m3 ? g : 0.f, alpha);
            return (d00 + d01) + (d10 + d11);
        }

        constexpr int kPgSeg = 2048; // float4 groups of the POOLED plane per block: 8 per thread, 32 K elements of dz

        // Pool2DGradient (2x2 max, stride 2) + ActivationGradient + first half of Conv2DBiasGradient in one pass: the backward of
        // "fused conv layer -> max pooling" (every VGG block boundary). Block (seg, k, n) owns float4 groups [seg*kPgSeg, ...) of
        // pooled plane (n, k); x = the pooling input = the convolution's activation output.
        template <int ACT>
        __global__ void __launch_bounds__(kAgThreads)
        pool2x2_gradient_act_bias_kernel(const float4* __restrict__ y, const float4* __restrict__ x, const float4* __restrict__ dy, float4* __restrict__ dz,
                                         float* __restrict__ partial, unsigned H, unsigned W4, unsigned Ho, unsigned Wo4, int segs, float alpha)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const int seg = blockIdx.x, k = blockIdx.y, n = blockIdx.z;
            const unsigned plane = (unsigned)n * gridDim.y + k;
            const unsigned quadsPerPlane = Ho * Wo4;
            const unsigned lo = seg * kPgSeg, hi = min(quadsPerPlane, lo + kPgSeg);
            float acc = 0.f;
            for (unsigned i = lo + threadIdx.x; i < hi; i += kAgThreads)
            {
                const unsigned ow4 = i % Wo4, oh = i / Wo4;
                const size_t q = (size_t)plane * quadsPerPlane + i;
                const size_t r0 = ((size_t)plane * H + 2 * oh) * W4 + 2 * ow4;
                const float4 g = __ldcs(dy + q), m = __ldcs(y + q);
                const float4 a0 = __ldcs(x + r0), a1 = __ldcs(x + r0 + 1), b0 = __ldcs(x + r0 + W4), b1 = __ldcs(x + r0 + W4 + 1);
                float4 t0, t1, u0, u1; // dz rows 2*oh (t) and 2*oh+1 (u), two float4 each
                acc += pool_act_window<ACT>(a0.x, a0.y, b0.x, b0.y, m.x, g.x, alpha, t0.x, t0.y, u0.x, u0.y);
                acc += pool_act_window<ACT>(a0.z, a0.w, b0.z, b0.w, m.y, g.y, alpha, t0.z, t0.w, u0.z, u0.w);
                acc += pool_act_window<ACT>(a1.x, a1.y, b1.x, b1.y, m.z, g.z, alpha, t1.x, t1.y, u1.x, u1.y);
                acc += pool_act_window<ACT>(a1.z, a1.w, b1.z, b1.w, m.w, g.w, alpha, t1.z, t1.w, u1.z, u1.w);
                dz[r0] = t0; dz[r0 + 1] = t1; dz[r0 + W4] = u0; dz[r0 + W4 + 1] = u1;
            }
            if (partial)
            {
                const float v = block_sum_256(acc);
                if (threadIdx.x == 0)
                    partial[(long long)k * gridDim.z * segs + (long long)n * segs + seg] = v;
            }
        }

        template <int ACT>
        int launch_pool_act_bias(const nb200_pool_desc& d, float alpha, const float* y, const float* x, const float* dy, float* dz, float* db,
                                 float* partial, cudaStream_t st)
        {
            const int segs = ceil_div((long long)d.Ho * d.Wo / 4, kPgSeg);
            const dim3 grid((unsigned)segs, (unsigned)d.C, (unsigned)d.N);
            NB200_CUDA_TRY(launch_kernel(pool2x2_gradient_act_bias_kernel<ACT>, dim3(grid), dim3(kAgThreads), 0, st, (const float4*)y, (const float4*)x, (const float4*)dy, (float4*)dz, db ? partial : nullptr, d.H, d.W / 4, d.Ho, d.Wo / 4, segs, alpha));
            NB200_CUDA_TRY(cudaGetLastError());
            count_launch();
            if (db)
            {
                NB200_CUDA_TRY(launch_kernel(bias_partial_reduce_kernel, dim3(ceil_div(d.C, 8)), dim3(256), 0, st, partial, db, d.C, d.N * segs));
                NB200_CUDA_TRY(cudaGetLastError());
                count_launch();
            }
            return NB200_OK;
        }

        // any layout, no reduction (NHWC callers: the channel of an element is not a block-uniform)
        template <int ACT>
        __global__ void act_gradient_flat_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dz,
                                                 long long n, float alpha)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
                dz[i] = activation_gradient<ACT>(y[i], dy[i], alpha);
        }

        template <int ACT>
        int launch_act_bias_gradient(const nb200_conv_desc& d, float alpha, const float* y, const float* dy, float* dz, float* db,
                                     float* partial, cudaStream_t st)
        {
            const int HW = d.Ho * d.Wo;
            if (d.fmt != NB200_NCHW)
            {
                const long long n = (long long)d.N * d.K * HW;
                const long long blocks = (n + 255) / 256;
                NB200_CUDA_TRY(launch_kernel(act_gradient_flat_kernel<ACT>, dim3((unsigned)(blocks > 148 * 32 ? 148 * 32 : blocks)), dim3(256), 0, st, y, dy, dz, n, alpha));
                NB200_CUDA_TRY(cudaGetLastError());
                count_launch();
                return db ? bias_gradient(d, dz, db, st) : NB200_OK;
            }
            const int segs = ceil_div(HW, kAgSeg);
            if (d.K > 65535 || d.N > 65535)
                return fail(NB200_E_UNSUPPORTED, "more than 65535 filters or images");
            const dim3 grid((unsigned)segs, (unsigned)d.K, (unsigned)d.N);
            const bool vec = HW % 4 == 0 && (((uintptr_t)y | (uintptr_t)dy | (uintptr_t)dz) & 15) == 0;
            if (vec)
                NB200_CUDA_TRY(launch_kernel(act_bias_gradient_kernel<ACT, true>, dim3(grid), dim3(kAgThreads), 0, st, y, dy, dz, db ? partial : nullptr, HW, d.K, segs, alpha));
            else
                NB200_CUDA_TRY(launch_kernel(act_bias_gradient_kernel<ACT, false>, dim3(grid), dim3(kAgThreads), 0, st, y, dy, dz, db ? partial : nullptr, HW, d.K, segs, alpha));
            NB200_CUDA_TRY(cudaGetLastError());
            count_launch();
            if (db)
            {
                NB200_CUDA_TRY(launch_kernel(bias_partial_reduce_kernel, dim3(ceil_div(d.K, 8)), dim3(256), 0, st, partial, db, d.K, d.N * segs));
                NB200_CUDA_TRY(cudaGetLastError());
                count_launch();
            }
            return NB200_OK;
        }

        // y = act(x + bias[k]): the bias AddOp (Operations/AddOp.cpp:38-50) and the Activation layer's forward
        // (TensorOpCpu.cpp:807-864) in one pass, for convolutions whose epilogue cannot carry them (a BatchNormalization sits
        // between the convolution and its activation in the GAN / pix2pix stacks). Bound: HBM, 8 bytes per element.
        template <bool VEC>
        __global__ void bias_activation_kernel(const float* __restrict__ x, const float* __restrict__ bias, float* __restrict__ y, long long n,
                                               int K, int HW, int nchw, int act, float alpha)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const long long stride = (long long)gridDim.x * blockDim.x;
            if (VEC)
            {
                for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < n; i += stride)
                {
                    float4 q = __ldcs(reinterpret_cast<const float4*>(x) + i);
                    if (bias)
                    {
                        if (nchw)
                        {
                            const float b = __ldg(bias + (int)(((i * 4) / HW) % K)); // HW % 4 == 0: one channel per float4
                            q.x += b; q.y += b; q.z += b; q.w += b;
                        }
                        else
                        {
                            const int k = (int)((i * 4) % K);                          // K % 4 == 0
                            q.x += __ldg(bias + k); q.y += __ldg(bias + k + 1); q.z += __ldg(bias + k + 2); q.w += __ldg(bias + k + 3);
                        }
                    }
                    q.x = apply_activation(act, alpha, q.x); q.y = apply_activation(act, alpha, q.y);
                    q.z = apply_activation(act, alpha, q.z); q.w = apply_activation(act, alpha, q.w);
                    reinterpret_cast<float4*>(y)[i] = q;
                }
            }
            else
            {
                for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
                {
                    float v = x[i];
                    if (bias)
                        v += __ldg(bias + (nchw ? (int)((i / HW) % K) : (int)(i % K)));
                    y[i] = apply_activation(act, alpha, v);
                }
            }
        }

        // TensorOpCpu::AdamStep (TensorOpCpu.cpp:987-1003) with the gradient pre-scale folded in.
        __global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                         float* __restrict__ v, size_t n, float gs, float lr, float b1, float b2, float eps)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= n)
                return;
            // every step rounded separately (no FMA contraction) in the reference's operation order, so that with gs == 1 the
            // update is bit-identical to the reference's compiled loop (tests/test_conv_gpu.py)
            const float gi = __fmul_rn(gs, g[i]);
            const float mi = __fadd_rn(__fmul_rn(b1, m[i]), __fmul_rn(1.f - b1, gi));
            const float vi = __fadd_rn(__fmul_rn(v[i], b2), __fmul_rn(__fmul_rn(1.f - b2, gi), gi));
            m[i] = mi;
            v[i] = vi;
            p[i] = __fsub_rn(p[i], __fmul_rn(__fdiv_rn(mi, __fadd_rn(__fsqrt_rn(vi), eps)), lr));
        }

        // TensorOpCpu::SgdStep (TensorOpCpu.cpp:1006-1009)
        __global__ void sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, size_t n, float gs, float lr)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
            if (i < n)
                p[i] = __fadd_rn(p[i], __fmul_rn(-lr, __fmul_rn(gs, g[i]))); // Tensor::Add(1, -lr, gradient): 1*p + (-lr)*g
        }
    }

    int bias_gradient(const nb200_conv_desc& d, const float* dy, float* db, cudaStream_t st)
    {
        if (d.K == 0)
            return NB200_OK;
        const ActStrides ys = act_strides(d.fmt, d.K, d.Ho, d.Wo);
        NB200_CUDA_TRY(launch_kernel(bias_gradient_kernel, dim3(d.K), dim3(kReduceThreads), 0, st, dy, db, d.N, d.K, d.Ho * d.Wo, ys, d.fmt == NB200_NCHW));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    size_t bias_activation_gradient_workspace(const nb200_conv_desc& d)
    {
        if (d.fmt != NB200_NCHW)
            return 0;
        return (size_t)d.K * d.N * ceil_div((long long)d.Ho * d.Wo, kAgSeg) * sizeof(float);
    }

    int bias_activation_gradient(const nb200_conv_desc& d, int act, float alpha, const float* y, const float* dy, float* dz, float* db,
                                 void* ws, size_t wsBytes, cudaStream_t st)
    {
        if (db && d.fmt == NB200_NCHW && (!ws || wsBytes < bias_activation_gradient_workspace(d)))
            return fail(NB200_E_WORKSPACE, "bias/activation gradient needs %zu workspace bytes, got %zu", bias_activation_gradient_workspace(d), wsBytes);
        float* partial = (float*)ws;
        switch (act)
        {
        case NB200_ACT_SIGMOID: return launch_act_bias_gradient<NB200_ACT_SIGMOID>(d, alpha, y, dy, dz, db, partial, st);
        case NB200_ACT_RELU: return launch_act_bias_gradient<NB200_ACT_RELU>(d, alpha, y, dy, dz, db, partial, st);
        case NB200_ACT_TANH: return launch_act_bias_gradient<NB200_ACT_TANH>(d, alpha, y, dy, dz, db, partial, st);
        case NB200_ACT_ELU: return launch_act_bias_gradient<NB200_ACT_ELU>(d, alpha, y, dy, dz, db, partial, st);
        case NB200_ACT_LEAKY_RELU: return launch_act_bias_gradient<NB200_ACT_LEAKY_RELU>(d, alpha, y, dy, dz, db, partial, st);
        default: return launch_act_bias_gradient<NB200_ACT_IDENTITY>(d, alpha, y, dy, dz, db, partial, st);
        }
    }

    bool pool_act_bias_supported(const nb200_pool_desc& d)
    {
        return d.fmt == NB200_NCHW && d.mode == NB200_POOL_MAX && d.filter == 2 && d.stride == 2 && d.padX == 0 && d.padY == 0 && d.H % 2 == 0 &&
               d.W % 8 == 0 && d.Ho == d.H / 2 && d.Wo == d.W / 2 && d.C <= 65535 && d.N <= 65535;
    }

    size_t pool_act_bias_workspace(const nb200_pool_desc& d)
    {
        return (size_t)d.C * d.N * ceil_div((long long)d.Ho * d.Wo / 4, kPgSeg) * sizeof(float);
    }

    int pool_act_bias_gradient(const nb200_pool_desc& d, int act, float alpha, const float* y, const float* x, const float* dy, float* dz, float* db,
                               void* ws, size_t wsBytes, cudaStream_t st)
    {
        if (db && (!ws || wsBytes < pool_act_bias_workspace(d)))
            return fail(NB200_E_WORKSPACE, "pool + activation gradient needs %zu workspace bytes, got %zu", pool_act_bias_workspace(d), wsBytes);
        float* partial = (float*)ws;
        switch (act)
        {
        case NB200_ACT_SIGMOID: return launch_pool_act_bias<NB200_ACT_SIGMOID>(d, alpha, y, x, dy, dz, db, partial, st);
        case NB200_ACT_RELU: return launch_pool_act_bias<NB200_ACT_RELU>(d, alpha, y, x, dy, dz, db, partial, st);
        case NB200_ACT_TANH: return launch_pool_act_bias<NB200_ACT_TANH>(d, alpha, y, x, dy, dz, db, partial, st);
        case NB200_ACT_ELU: return launch_pool_act_bias<NB200_ACT_ELU>(d, alpha, y, x, dy, dz, db, partial, st);
        case NB200_ACT_LEAKY_RELU: return launch_pool_act_bias<NB200_ACT_LEAKY_RELU>(d, alpha, y, x, dy, dz, db, partial, st);
        default: return launch_pool_act_bias<NB200_ACT_IDENTITY>(d, alpha, y, x, dy, dz, db, partial, st);
        }
    }

    int bias_activation(const nb200_conv_desc& d, const float* x, const float* bias, int act, float alpha, float* y, cudaStream_t st)
    {
        const int HW = d.Ho * d.Wo;
        const long long n = (long long)d.N * d.K * HW;
        if (n == 0)
            return NB200_OK;
        const int nchw = d.fmt == NB200_NCHW;
        const bool vec = ((nchw ? HW : d.K) % 4 == 0) && ((((uintptr_t)x | (uintptr_t)y) & 15) == 0);
        const long long work = vec ? n / 4 : n;
        const long long blocks = (work + 255) / 256;
        const unsigned grid = (unsigned)(blocks > 148 * 32 ? 148 * 32 : blocks);
        if (vec)
            NB200_CUDA_TRY(launch_kernel(bias_activation_kernel<true>, dim3(grid), dim3(256), 0, st, x, bias, y, n, d.K, HW, nchw, act, alpha));
        else
            NB200_CUDA_TRY(launch_kernel(bias_activation_kernel<false>, dim3(grid), dim3(256), 0, st, x, bias, y, n, d.K, HW, nchw, act, alpha));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int adam_step(float* p, const float* g, float* m, float* v, size_t n, float gs, float lr, float b1, float b2, float eps,
                  cudaStream_t st)
    {
        if (n == 0)
            return NB200_OK;
        NB200_CUDA_TRY(launch_kernel(adam_step_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, p, g, m, v, n, gs, lr, b1, b2, eps));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int sgd_step(float* p, const float* g, size_t n, float gs, float lr, cudaStream_t st)
    {
        if (n == 0)
            return NB200_OK;
        NB200_CUDA_TRY(launch_kernel(sgd_step_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, p, g, n, gs, lr));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }
}
