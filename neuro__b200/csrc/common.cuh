// Shared device/host helpers for the Conv2D kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/neuro_b200.h"

namespace nb200
{
    // Thread-local last-error slot behind nb200_last_error().
    void set_error(const char* fmt, ...);
    void count_launch(int n = 1); // feeds nb200_kernel_launches()
    int fail(int code, const char* fmt, ...);

#define NB200_CUDA_TRY(expr)                                                                      \
    do                                                                                            \
    {                                                                                             \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return nb200::fail(NB200_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

    namespace ptx
    {
        // Programmatic dependent launch (launch attribute cudaLaunchAttributeProgrammaticStreamSerialization, launch_kernel below): the NEXT kernel of the stream may start its CTAs -- and run everything up to its own pdl_wait() --
        // once every CTA of this grid has executed pdl_launch_dependents() (or exited); pdl_wait() blocks until the previous
        // grid has completed and its memory is visible. Kernels call both right after their prologue (barrier init, TMEM
        // allocation, tensor-map prefetch), which therefore overlaps the previous kernel's tail; without the launch attribute
        // both instructions are no-ops.
        __device__ __forceinline__ void pdl_launch_dependents()
        {
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        }
        __device__ __forceinline__ void pdl_wait()
        {
            asm volatile("griddepcontrol.wait;" ::: "memory");
        }
    }

    // Launch with programmatic stream serialisation (see ptx::pdl_wait above). ONLY for kernels that execute pdl_wait()
    // before their first global-memory access; every other kernel keeps the <<< >>> launch and with it full stream order.
    // NB200_PDL=0 disables the attribute (then the device-side instructions are no-ops).
    bool pdl_enabled();
    template <typename... KArgs, typename... Args>
    cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smemBytes, cudaStream_t st, Args&&... args)
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smemBytes; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
        return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    }

    // Element strides of an activation tensor in either data format, so one kernel body serves
    // NCHW and NHWC (Neuro::Shape semantics, Neuro/src/Tensors/Shape.cpp:11-22).
    struct ActStrides
    {
        long long n, c, h, w;
    };

    __host__ __device__ inline ActStrides act_strides(int fmt, int C, int H, int W)
    {
        ActStrides s;
        if (fmt == NB200_NCHW)
        {
            s.w = 1; s.h = W; s.c = (long long)H * W; s.n = (long long)C * H * W;
        }
        else
        {
            s.c = 1; s.w = C; s.h = (long long)W * C; s.n = (long long)H * W * C;
        }
        return s;
    }

    // Activations as the reference evaluates them (Neuro/src/Tensors/TensorOpCpu.cpp:807-864).
    __device__ __forceinline__ float apply_activation(int act, float alpha, float v)
    {
        switch (act)
        {
        case NB200_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case NB200_ACT_RELU: return v > 0.f ? v : 0.f;
        case NB200_ACT_TANH: return 2.f / (1.f + expf(-2.f * v)) - 1.f;
        case NB200_ACT_ELU: return v >= 0.f ? v : alpha * (expf(v) - 1.f);
        case NB200_ACT_LEAKY_RELU: return v >= 0.f ? v : alpha * v;
        default: return v;
        }
    }

    inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

    // ---- kernel families (each returns an NB200_* status) ----

    // CUDA-core fp32 kernels: every shape, both formats, any stride/pad. conv_direct.cu
    int direct_forward(const nb200_conv_desc& d, const float* x, const float* w, const float* bias, int act, float alpha,
                       float* y, cudaStream_t st);
    int direct_input_gradient(const nb200_conv_desc& d, const float* dy, const float* w, float* dx, cudaStream_t st);
    int direct_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes,
                                cudaStream_t st);
    size_t direct_kernels_gradient_workspace(const nb200_conv_desc& d);

    // HBM-bound 3x3 stride-1 kernels for C <= 4 input channels (first layers). conv_smallc.cu
    bool smallc_supported(const nb200_conv_desc& d);
    size_t smallc_wgrad_workspace(const nb200_conv_desc& d);
    int smallc_forward(const nb200_conv_desc& d, const float* x, const float* w, const float* bias, int act, float alpha, float* y,
                       cudaStream_t st);
    int smallc_input_gradient(const nb200_conv_desc& d, const float* dy, const float* w, float* dx, cudaStream_t st);
    int smallc_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes,
                                cudaStream_t st);

    int smallc_input_gradient_epilogue(const nb200_conv_desc& d, const float* dy, const float* w, const float* bias, int act, float alpha,
                                       float* dx, cudaStream_t st);
    // few-filter layers (K <= 4) = the small-channel problem with its activation tensors exchanged. conv_smallc.cu
    bool smallk_supported(const nb200_conv_desc& d);
    nb200_conv_desc smallk_swapped(const nb200_conv_desc& d);
    int smallc_swap_filters(const float* in, float* out, int A, int B, cudaStream_t st); // out[b][a][2-r][2-s] = in[a][b][r][s]

    // stride-2 kernel gradient with 1-6 channels (first layers of the GAN discriminators / U-Net encoder). conv_smallc.cu
    bool strided_wgrad_supported(const nb200_conv_desc& d);
    size_t strided_wgrad_workspace(const nb200_conv_desc& d);
    int strided_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes, cudaStream_t st);

    // tcgen05/TMA implicit-GEMM kernels (NCHW). conv_tc.cu
    bool tc_forward_supported(const nb200_conv_desc& d);
    bool tc_input_gradient_supported(const nb200_conv_desc& d);
    bool tc_kernels_gradient_supported(const nb200_conv_desc& d);
    size_t tc_workspace_bytes(int op, const nb200_conv_desc& d);
    // C <= 4 kernel gradient as an SS-form tcgen05 GEMM over the dy stream (TF32 mode, large maps)
    bool tc_smallc_wgrad_supported(const nb200_conv_desc& d);
    size_t tc_smallc_wgrad_workspace(const nb200_conv_desc& d);
    int tc_smallc_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes, cudaStream_t st);
    // the same GEMM for strided / larger-filter first layers (C*R*S <= 96), im2col rows gathered from global memory
    bool tc_smallc_wgrad_gather_supported(const nb200_conv_desc& d);
    size_t tc_smallc_wgrad_gather_workspace(const nb200_conv_desc& d);
    int tc_smallc_gather_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes, cudaStream_t st);
    bool tc_uses_rowfold(const nb200_conv_desc& d);        // kernel gradient takes tc_wgrad_rowfold_kernel
    bool tc_uses_rowtap(int op, const nb200_conv_desc& d); // forward / stride-1 input gradient take tc_rowtap_kernel
    int tc_forward(const nb200_conv_desc& d, const float* x, const float* w, const float* bias, int act, float alpha, float* y,
                   void* ws, size_t wsBytes, cudaStream_t st);
    int tc_input_gradient(const nb200_conv_desc& d, const float* dy, const float* w, float* dx, void* ws, size_t wsBytes,
                          cudaStream_t st);
    int tc_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes,
                            cudaStream_t st);

    // What the forward / input-gradient launchers do about the repacked filters at the head of the workspace:
    // repack on every call (default), trust what nb200_conv2d_prepare_filters left there, or repack and return.
    enum { kFiltersRepack = 0, kFiltersReady = 1, kFiltersOnly = 2 };
    extern thread_local int g_tcFilterMode;

    // gathered-A tensor-core kernel: any stride / padding / map size, forward and input gradient. conv_tc.cu
    bool tc_gather_forward_supported(const nb200_conv_desc& d);
    bool tc_gather_input_gradient_supported(const nb200_conv_desc& d);
    size_t tc_gather_workspace_bytes(int op, const nb200_conv_desc& d);
    int tc_gather_forward(const nb200_conv_desc& d, const float* x, const float* w, const float* bias, int act, float alpha, float* y,
                          void* ws, size_t wsBytes, cudaStream_t st);
    int tc_gather_input_gradient(const nb200_conv_desc& d, const float* dy, const float* w, float* dx, void* ws, size_t wsBytes,
                                 cudaStream_t st);

    bool tc_gather_kernels_gradient_supported(const nb200_conv_desc& d);
    size_t tc_gather_kernels_gradient_workspace(const nb200_conv_desc& d);
    int tc_gather_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes,
                                   cudaStream_t st);

    // resample.cu: NHWC <-> NCHW layout pass (per image [HW][C] <-> [C][HW])
    int layout_transpose(const float* src, float* dst, int N, int C, int HW, bool toNchw, cudaStream_t st);

    // elementwise.cu
    int bias_gradient(const nb200_conv_desc& d, const float* dy, float* db, cudaStream_t st);
    // dz = act'(y) * dy and (optionally) db = sum over N,H,W of dz, one pass over HBM
    size_t bias_activation_gradient_workspace(const nb200_conv_desc& d);
    int bias_activation_gradient(const nb200_conv_desc& d, int act, float alpha, const float* y, const float* dy, float* dz, float* db,
                                 void* ws, size_t wsBytes, cudaStream_t st);
    // dz = act'(x) * maxpool2x2_gradient(y, x, dy), db = sum dz: backward of "fused conv layer -> 2x2 max pooling" in one pass
    bool pool_act_bias_supported(const nb200_pool_desc& d);
    size_t pool_act_bias_workspace(const nb200_pool_desc& d);
    int pool_act_bias_gradient(const nb200_pool_desc& d, int act, float alpha, const float* y, const float* x, const float* dy, float* dz, float* db,
                               void* ws, size_t wsBytes, cudaStream_t st);
    // y = act(x + bias): bias add + activation forward as one pass (layers whose conv epilogue cannot carry them)
    int bias_activation(const nb200_conv_desc& d, const float* x, const float* bias, int act, float alpha, float* y, cudaStream_t st);
    int adam_step(float* p, const float* g, float* m, float* v, size_t n, float gs, float lr, float b1, float b2, float eps,
                  cudaStream_t st);
    int sgd_step(float* p, const float* g, size_t n, float gs, float lr, cudaStream_t st);

    // batchnorm.cu -- statistics are split from the normalisation so that replicas can exchange them (2*G / 3*G floats)
    int bn_check(const nb200_bn_desc* d);
    int bn_groups(const nb200_bn_desc& d);
    long long bn_group_elements(const nb200_bn_desc& d);
    size_t bn_workspace_bytes(const nb200_bn_desc& d);
    int bn_moments(const nb200_bn_desc& d, const float* x, float* moments, void* ws, size_t wsBytes, cudaStream_t st);
    int bn_finalize(const nb200_bn_desc& d, const float* allMoments, int replicas, float momentum, float epsilon, float* runningMean,
                    float* runningVar, float* saveMean, float* saveInvVar, cudaStream_t st);
    int bn_apply(const nb200_bn_desc& d, bool inference, const float* x, const float* gamma, const float* beta, const float* mean,
                 const float* invOrVar, float epsilon, float* y, cudaStream_t st);
    int bn_gradient_sums(const nb200_bn_desc& d, const float* x, const float* dy, const float* saveMean, float* sums, void* ws, size_t wsBytes,
                         cudaStream_t st);
    int bn_gradient_apply(const nb200_bn_desc& d, int replicas, const float* x, const float* gamma, const float* dy, const float* saveMean,
                          const float* saveInvVar, const float* globalSums, const float* localSums, float* dgamma, float* dbeta, float* dx,
                          cudaStream_t st);
}
