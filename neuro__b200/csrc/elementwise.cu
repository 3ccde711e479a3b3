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

        // TensorOpCpu::AdamStep (TensorOpCpu.cpp:987-1003) with the gradient pre-scale folded in.
        __global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                         float* __restrict__ v, size_t n, float gs, float lr, float b1, float b2, float eps)
        {
            const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= n)
                return;
            const float gi = gs * g[i];
            const float mi = b1 * m[i] + (1.f - b1) * gi;
            const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
            m[i] = mi;
            v[i] = vi;
            p[i] = p[i] - mi / (sqrtf(vi) + eps) * lr;
        }

        // TensorOpCpu::SgdStep (TensorOpCpu.cpp:1006-1009)
        __global__ void sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, size_t n, float gs, float lr)
        {
            const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
            if (i < n)
                p[i] = p[i] - lr * (gs * g[i]);
        }
    }

    int bias_gradient(const nb200_conv_desc& d, const float* dy, float* db, cudaStream_t st)
    {
        if (d.K == 0)
            return NB200_OK;
        const ActStrides ys = act_strides(d.fmt, d.K, d.Ho, d.Wo);
        bias_gradient_kernel<<<d.K, kReduceThreads, 0, st>>>(dy, db, d.N, d.K, d.Ho * d.Wo, ys, d.fmt == NB200_NCHW);
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int adam_step(float* p, const float* g, float* m, float* v, size_t n, float gs, float lr, float b1, float b2, float eps,
                  cudaStream_t st)
    {
        if (n == 0)
            return NB200_OK;
        adam_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, g, m, v, n, gs, lr, b1, b2, eps);
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int sgd_step(float* p, const float* g, size_t n, float gs, float lr, cudaStream_t st)
    {
        if (n == 0)
            return NB200_OK;
        sgd_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, g, n, gs, lr);
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }
}
