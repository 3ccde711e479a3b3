/*
 * neuro_b200.h -- C ABI of the B200-native convolution backend for Neuro_.
 *
 * The reference has no plugin loader or FFI: a backend is a C++ subclass of Neuro::TensorOpCpu
 * (Neuro/include/Tensors/TensorOpCpu.h:7-78) returned by Tensor::GetOpFromMode
 * (Neuro/src/Tensors/Tensor.cpp:2701-2716). This header is the seam we define underneath such a
 * subclass (include/neuro_b200/TensorOpB200.h mirrors the virtual interface and unpacks Tensors
 * into these calls; INTEGRATION.md shows the binding a Neuro_ maintainer would add).
 *
 * Conventions shared by every entry point
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - all tensors are dense fp32. Activation pointers (x, y, dx, dy) are DEVICE pointers unless the
 *     function name ends in _host;
 *   - layouts follow Neuro::Shape (Neuro/src/Tensors/Shape.cpp:11-22):
 *       NB200_NCHW  x[((n*C+c)*H+h)*W+w]      (reference Shape(W,H,C,N))
 *       NB200_NHWC  x[((n*H+h)*W+w)*C+c]      (reference Shape(C,W,H,N))
 *       kernels     w[((k*C+c)*R+r)*S+s]      (reference Shape(S,R,C,K), both formats)
 *   - outputs are OVERWRITTEN (beta = 0), like the reference CPU ops, which zero or assign them
 *     (TensorOpCpu.cpp:1032,1076,1134); the caller owns and pre-sizes every tensor
 *     (Tensor.cpp:1767; Conv2DOp.cpp:18,27); nothing is allocated inside except through `workspace`;
 *   - calls are asynchronous on `stream`; no hidden device synchronisation;
 *   - return value: 0 on success, a negative NB200_E_* code otherwise; nb200_last_error() returns a
 *     thread-local message. (The reference's ops are void and assert in debug builds only,
 *     Neuro/include/Types.h:9-22; its CUDA_CHECK is a no-op in release, CudaErrorCheck.h:18-22.
 *     We always check.)
 *   - there is NO CPU fallback: without a usable CUDA device every compute call fails with
 *     NB200_E_NO_DEVICE.
 */
#ifndef NEURO_B200_H
#define NEURO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB200_API __attribute__((visibility("default")))

/* EDataFormat, Neuro/include/Types.h:94-98 */
enum { NB200_NCHW = 0, NB200_NHWC = 1 };

/* EActivation, Neuro/include/Types.h:83-92 (same numbering; _Softmax is not an epilogue) */
enum { NB200_ACT_IDENTITY = 0, NB200_ACT_SIGMOID = 1, NB200_ACT_RELU = 2, NB200_ACT_TANH = 3,
       NB200_ACT_ELU = 4, NB200_ACT_LEAKY_RELU = 5 };

/* Arithmetic used for the contraction (BASELINE.json north_star: TF32 <= 2e-3, 3xTF32 <= 1e-5
 * max-normalised error against the reference CPU ops). */
enum {
    NB200_MATH_TF32   = 0, /* tcgen05 kind::tf32, fp32 accumulate in TMEM (default)                    */
    NB200_MATH_3XTF32 = 1, /* split-operand TF32 (hi*hi + hi*lo + lo*hi), fp32 accumulate              */
    NB200_MATH_FP32   = 2  /* CUDA-core fp32 FMA kernels; also what HBM-bound small-channel layers use */
};

enum {
    NB200_OK = 0,
    NB200_E_INVALID = -1,     /* bad descriptor / null pointer / inconsistent sizes */
    NB200_E_NO_DEVICE = -2,   /* no CUDA device, or not an sm_100 part              */
    NB200_E_CUDA = -3,        /* a CUDA runtime/driver call or launch failed        */
    NB200_E_WORKSPACE = -4,   /* workspace smaller than nb200_conv2d_workspace_bytes */
    NB200_E_UNSUPPORTED = -5
};

enum { NB200_OP_FORWARD = 0, NB200_OP_INPUT_GRADIENT = 1, NB200_OP_KERNELS_GRADIENT = 2 };

/* One convolution problem. (N,C,H,W) is always the extent of the tensor on the INPUT side of the
 * forward op (x, or dx for the input gradient); (Ho,Wo) the extent on its OUTPUT side (y, or dy).
 * Both are supplied by the caller, exactly as the reference ops take them from the Tensor shapes:
 * the input gradient of a strided conv, and the forward of Conv2DTranspose
 * (Tensor.cpp:1806-1810, 2032-2051), rely on (H,W) not being derived from (Ho,Wo). */
typedef struct nb200_conv_desc
{
    int32_t N, C, H, W;
    int32_t K, R, S;
    int32_t Ho, Wo;
    int32_t stride;
    int32_t padX, padY;
    int32_t fmt;   /* NB200_NCHW | NB200_NHWC */
    int32_t math;  /* NB200_MATH_*            */
} nb200_conv_desc;

/* Library / device introspection. */
NB200_API const char* nb200_version(void);
NB200_API const char* nb200_last_error(void);
/* Number of CUDA kernels this library has launched in this process (for benchmark bookkeeping). */
NB200_API unsigned long long nb200_kernel_launches(void);
/* 0 if the current device can run the kernels (compute capability 10.x); fills optional outputs. */
NB200_API int nb200_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* hbm_bytes);

/* Shape helpers = Tensor::GetPadding / GetConvOutputShape / GetConvTransposeOutputShape
 * (Neuro/src/Tensors/Tensor.cpp:1966-1985, 2010-2029, 2032-2051). mode: 0 Valid, 1 Same, 2 Full. */
NB200_API int32_t nb200_padding(int32_t mode, int32_t filter);
NB200_API int32_t nb200_conv_out_size(int32_t in, int32_t filter, int32_t stride, int32_t pad);
NB200_API int32_t nb200_conv_transpose_out_size(int32_t in, int32_t filter, int32_t stride, int32_t pad);

/* Bytes of device scratch the op may use for this problem (0 is a valid answer). */
NB200_API size_t nb200_conv2d_workspace_bytes(int32_t op, const nb200_conv_desc* d);

/* Name of the kernel family the dispatcher picks for this problem; static storage. This is the advertised way to find out
 * whether a problem runs on the tensor cores ("tcgen05_*") or on a CUDA-core family -- nothing degrades silently:
 *
 *   math \ layout, shape        NCHW, C and K > 4                NCHW, C <= 4 or K <= 4 (HBM-bound)        NHWC
 *   NB200_MATH_TF32             tcgen05_{fprop,dgrad,wgrad}      smallc_* / smallk_* (fp32 FMA) forward    tcgen05_*_nhwc (the NCHW
 *                               tcgen05_rowtap_*, _rowfold_wgrad and input gradient; kernel gradient:    kernels between two layout
 *                               tcgen05_gather_* (strides,       tcgen05_smallc_wgrad,                    passes), else direct_*
 *                               small / odd maps)                tcgen05_smallc_gather_wgrad (strided)
 *   NB200_MATH_3XTF32           tcgen05_{fprop,dgrad},           as above (fp32 FMA kernels serve every   as above
 *                               tcgen05_gather_{fprop,dgrad,     math mode; small strided kernel
 *                               wgrad}                           gradients: strided_smallc_wgrad)
 *   NB200_MATH_FP32             direct_* (fp32 FMA)              smallc_* / smallk_* / strided_*          direct_*
 *
 * Grid-size heuristics assume the 148 SMs of a B200 (workspace sizes must be computable without a device); persistent
 * grids query the device's SM count at launch. */
NB200_API const char* nb200_conv2d_kernel_name(int32_t op, const nb200_conv_desc* d);

/* y = act(conv(x, w) + bias).
 * Replaces TensorOpCpu::Conv2D (TensorOpCpu.h:46, TensorOpCpu.cpp:1012) when bias == NULL and
 * act == NB200_ACT_IDENTITY, and TensorOpCpu::Conv2DBiasActivation (TensorOpCpu.h:47, .cpp:1055)
 * otherwise. bias: K floats (reference Shape(1,1,K)). alpha: ELU / LeakyReLU coefficient. */
NB200_API int nb200_conv2d_forward(const nb200_conv_desc* d, const float* x, const float* w,
                                   const float* bias, int32_t act, float alpha, float* y,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* dx = conv_input_gradient(dy, w); dx has extent (N,C,H,W), dy (N,K,Ho,Wo).
 * Replaces TensorOpCpu::Conv2DInputGradient (TensorOpCpu.h:49, TensorOpCpu.cpp:1071). Also the forward
 * of Conv2DTranspose and the image gradient of style transfer. Elements of dx that no output tap
 * reaches (ragged strides) are written as 0. */
NB200_API int nb200_conv2d_input_gradient(const nb200_conv_desc* d, const float* dy, const float* w,
                                          float* dx, void* workspace, size_t workspace_bytes, void* stream);

/* dw = conv_kernels_gradient(x, dy); optionally db[k] = sum_{n,ho,wo} dy in the same CALL (a separate pass over dy
 * inside it: use nb200_conv2d_bias_activation_gradient / nb200_pool2d_gradient_activation to get db for free from the pass
 * that produces dy). Replaces TensorOpCpu::Conv2DKernelsGradient (TensorOpCpu.h:50, TensorOpCpu.cpp:1129); db != NULL
 * additionally replaces Conv2DBiasGradient (TensorOpCpu.h:48). */
NB200_API int nb200_conv2d_kernels_gradient(const nb200_conv_desc* d, const float* x, const float* dy,
                                            float* dw, float* db, void* workspace, size_t workspace_bytes,
                                            void* stream);

/* db[k] = sum over N,Ho,Wo of dy. Replaces TensorOpCpu::Conv2DBiasGradient (TensorOpCpu.h:48,
 * TensorOpCpu.cpp:1065 = Sum over _013Axes). Only N,K,Ho,Wo,fmt of the descriptor are read. */
NB200_API int nb200_conv2d_bias_gradient(const nb200_conv_desc* d, const float* dy, float* db, void* stream);

/* Backward prologue of the fused bias+activation convolution, Conv2dBiasActivationOp::ComputeGradientInternal
 * (Neuro/src/ComputationalGraph/Operations/Conv2dBiasActivationOp.cpp:47-60), in one pass over HBM:
 *   dz = act'(y) * dy   -- Tensor::ActivationGradient -> TensorOpCpu::{Sigmoid,Tanh,ReLU,Elu,LeakyReLU}Gradient
 *                          (TensorOpCpu.cpp:813-864; the derivative is taken through the OUTPUT y, as there);
 *   db[k] = sum_{n,ho,wo} dz   -- TensorOpCpu::Conv2DBiasGradient (TensorOpCpu.cpp:1065-1068); db may be NULL.
 * dz is then the `gradient` argument of nb200_conv2d_input_gradient / nb200_conv2d_kernels_gradient.
 * y, dy, dz have extent (N,K,Ho,Wo) in d->fmt; only N,K,Ho,Wo,fmt of the descriptor are read. dz must not alias y/dy.
 * Workspace: nb200_conv2d_bias_activation_gradient_workspace_bytes (per-block partial sums, added in fixed order). */
NB200_API size_t nb200_conv2d_bias_activation_gradient_workspace_bytes(const nb200_conv_desc* d);
NB200_API int nb200_conv2d_bias_activation_gradient(const nb200_conv_desc* d, int32_t act, float alpha, const float* y,
                                                    const float* dy, float* dz, float* db, void* workspace,
                                                    size_t workspace_bytes, void* stream);

/* y = act(x + bias[k]) over an (N,K,Ho,Wo) tensor in d->fmt (only those fields of the descriptor are read); bias may be
 * NULL, y may alias x. One pass for the bias AddOp (Neuro/src/ComputationalGraph/Operations/AddOp.cpp:38-50) and the
 * Activation layer's forward (Tensor::Activation -> TensorOpCpu.cpp:807-864) of layers whose convolution cannot carry
 * them in its epilogue -- a BatchNormalization sits between the conv and the activation in the GAN / pix2pix stacks
 * (Neuro.Examples/src/DeepConvGAN.cpp:3-43, Pix2Pix.cpp:4-109). Its gradient is nb200_conv2d_bias_activation_gradient. */
NB200_API int nb200_bias_activation(const nb200_conv_desc* d, const float* x, const float* bias, int32_t act, float alpha,
                                    float* y, void* stream);

/* Filters that do not change between calls (inference; style transfer runs forward and input gradient against frozen
 * VGG weights, Neuro.Examples/include/NeuralStyleTransfer.h). The tensor-core kernels read the filters from a repacked
 * TF32 copy at the head of the workspace; nb200_conv2d_forward / _input_gradient rebuild it on every call because the
 * reference interface cannot tell them the filters are unchanged. nb200_conv2d_prepare_filters builds it once into a
 * workspace the caller dedicates to this (op, descriptor, w) -- size nb200_conv2d_workspace_bytes(op, d) -- and the
 * *_prepared calls skip the repack launch. op: NB200_OP_FORWARD or NB200_OP_INPUT_GRADIENT. Kernel families that read
 * w directly (first-layer and fp32 kernels) make prepare a no-op, so w must stay valid and unchanged either way. */
NB200_API int nb200_conv2d_prepare_filters(int32_t op, const nb200_conv_desc* d, const float* w, void* workspace,
                                           size_t workspace_bytes, void* stream);
NB200_API int nb200_conv2d_forward_prepared(const nb200_conv_desc* d, const float* x, const float* w, const float* bias,
                                            int32_t act, float alpha, float* y, void* workspace, size_t workspace_bytes,
                                            void* stream);
NB200_API int nb200_conv2d_input_gradient_prepared(const nb200_conv_desc* d, const float* dy, const float* w, float* dx,
                                                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- plans: one problem, fixed buffers, ONE launch call per run --------------------------------------------------------
 * The reference-shaped entry points above re-derive everything on every call, as the reference interface demands
 * (descriptor validation, kernel choice, cuTensorMapEncodeTiled of 2 tensor maps, the filter repack launch; the
 * reference's cuDNN path did the same with descriptors + algorithm search, TensorOpGpu.cpp:629-668). A PLAN binds one
 * (op, descriptor) to fixed device buffers and records the op's kernel launches -- tensor maps, grids, workspace layout
 * baked in -- as an instantiated CUDA graph; nb200_conv2d_plan_run replays it on any stream with no host-side work
 * beyond one cudaGraphLaunch. Results are bit-identical to the plain call (same kernels, same order).
 *   in / out by op:  FORWARD          a = x,  b = w,  out = y   (+ bias, act, alpha)
 *                    INPUT_GRADIENT   a = dy, b = w,  out = dx
 *                    KERNELS_GRADIENT a = x,  b = dy, out = dw  (+ db in `bias`, may be NULL)
 * filters_constant != 0 (FORWARD / INPUT_GRADIENT): w is repacked ONCE at plan creation (nb200_conv2d_prepare_filters)
 * and runs skip the repack launch -- inference, style transfer against frozen VGG weights; w must then stay unchanged.
 * The workspace (nb200_conv2d_workspace_bytes) is owned by the caller, dedicated to the plan and must outlive it.
 * A plan belongs to the device that was current when it was created. */
typedef struct nb200_conv_plan nb200_conv_plan;
NB200_API int nb200_conv2d_plan_create(int32_t op, const nb200_conv_desc* d, const float* a, const float* b, float* out,
                                       float* bias, int32_t act, float alpha, int32_t filters_constant,
                                       void* workspace, size_t workspace_bytes, nb200_conv_plan** plan);
NB200_API int nb200_conv2d_plan_run(const nb200_conv_plan* plan, void* stream);
/* Number of kernel nodes the plan replays (for reports). */
NB200_API int32_t nb200_conv2d_plan_kernels(const nb200_conv_plan* plan);
NB200_API void nb200_conv2d_plan_destroy(nb200_conv_plan* plan);

/* ---- spatial resamplers on either side of the convolutions; HBM-bound, bit-exact vs the reference ----
 * EPoolingMode, Neuro/include/Types.h:56-60 */
enum { NB200_POOL_MAX = 0, NB200_POOL_AVG = 1 };

/* (N,C,H,W) = input extent, (Ho,Wo) = pooled extent = Tensor::GetPooling2DOutputShape (Tensor.cpp:1988-2007). */
typedef struct nb200_pool_desc
{
    int32_t N, C, H, W;
    int32_t Ho, Wo;
    int32_t filter, stride;
    int32_t padX, padY;
    int32_t mode; /* NB200_POOL_* */
    int32_t fmt;  /* NB200_NCHW | NB200_NHWC */
} nb200_pool_desc;

/* y = pool(x). Replaces TensorOpCpu::Pool2D (TensorOpCpu.h:51, TensorOpCpu.cpp:1187-1246; Mt :336-404): padded taps read
 * -FLT_MAX (max) or 0 (avg); the average divides by filter*filter whatever the padding. */
NB200_API int nb200_pool2d(const nb200_pool_desc* d, const float* x, float* y, void* stream);
/* dx = pool gradient. Replaces TensorOpCpu::Pool2DGradient (TensorOpCpu.h:52, TensorOpCpu.cpp:1249-1338): max -> the first
 * window element (row-major) equal to the pooled value receives dy; avg -> every in-range element receives dy/filter^2;
 * overlapping windows add in (oh, ow) order. dx is overwritten (the reference zeroes it first). y, x may be NULL for avg. */
NB200_API int nb200_pool2d_gradient(const nb200_pool_desc* d, const float* y, const float* x, const float* dy, float* dx,
                                    void* stream);
/* Backward of "fused convolution layer -> 2x2 max pooling" (every VGG block boundary) in ONE pass over HBM:
 *   dz = act'(x) * Pool2DGradient(y, x, dy)      -- TensorOpCpu::Pool2DGradient (TensorOpCpu.cpp:1249-1338), then the
 *   db[c] = sum_{n,h,w} dz                           ActivationGradient + Conv2DBiasGradient of the layer that produced x
 *                                                    (Conv2dBiasActivationOp.cpp:47-60; x is that layer's activation output)
 * bit-identical to nb200_pool2d_gradient followed by nb200_conv2d_bias_activation_gradient, at 2.5 instead of 5.5 tensor
 * passes. Max pooling 2x2 stride 2 without padding, NCHW, H % 2 == 0, W % 8 == 0, 16-byte aligned tensors; anything else
 * returns NB200_E_UNSUPPORTED (nb200_pool2d_gradient_activation_supported tells) and the caller issues the two calls.
 * db may be NULL (then no workspace is needed). */
NB200_API int32_t nb200_pool2d_gradient_activation_supported(const nb200_pool_desc* d);
NB200_API size_t nb200_pool2d_gradient_activation_workspace_bytes(const nb200_pool_desc* d);
NB200_API int nb200_pool2d_gradient_activation(const nb200_pool_desc* d, int32_t act, float alpha, const float* y, const float* x,
                                               const float* dy, float* dz, float* db, void* workspace, size_t workspace_bytes,
                                               void* stream);

/* y[n,c,oh,ow] = x[n,c,oh/scale,ow/scale]; (N,C,H,W) = INPUT extent, NCHW planes. Replaces TensorOpCpu::UpSample2D
 * (TensorOpCpu.h:53, TensorOpCpu.cpp:1340-1354). */
NB200_API int nb200_upsample2d(int32_t N, int32_t C, int32_t H, int32_t W, int32_t scale, const float* x, float* y, void* stream);
/* dx[n,c,h,w] = sum of the scale x scale block of dy, added row by row. Replaces TensorOpCpu::UpSample2DGradient
 * (TensorOpCpu.h:54, TensorOpCpu.cpp:1357-1369). (N,C,H,W) = extent of dx. */
NB200_API int nb200_upsample2d_gradient(int32_t N, int32_t C, int32_t H, int32_t W, int32_t scale, const float* dy, float* dx,
                                        void* stream);
/* y = x framed by `value`; (N,C,H,W) = INPUT extent, output (H+top+bottom) x (W+left+right). Replaces
 * TensorOpCpu::ConstantPad2D (TensorOpCpu.cpp:528-546; PatchGAN's ZeroPadding2D). */
NB200_API int nb200_constant_pad2d(int32_t N, int32_t C, int32_t H, int32_t W, int32_t left, int32_t right, int32_t top,
                                   int32_t bottom, float value, const float* x, float* y, void* stream);

/* Optimiser updates that follow the gradient exchange in data-parallel Fit().
 * Replace TensorOpCpu::AdamStep / SgdStep (TensorOpCpu.h:75-76, TensorOpCpu.cpp:987-1009), with the
 * 1/replicas scaling of an all-reduced (summed) gradient folded in as grad_scale:
 *   g' = grad_scale*g;  m = b1*m + (1-b1)*g';  v = b2*v + (1-b2)*g'^2;  p -= lr * m / (sqrt(v) + eps). */
NB200_API int nb200_adam_step(float* param, const float* grad, float* m, float* v, size_t count,
                              float grad_scale, float lr, float beta1, float beta2, float epsilon, void* stream);
NB200_API int nb200_sgd_step(float* param, const float* grad, size_t count, float grad_scale, float lr, void* stream);

/* ---- batch normalisation around the convolutions; HBM-bound ----
 * EBatchNormMode, Neuro/include/Types.h:70-75 (same numbering). x is NCHW (N,C,H,W); statistics have G values:
 * PerActivation G = C*H*W (over N), Spatial G = C (over N,H,W), Instance G = N*C (over H,W). */
enum { NB200_BN_PER_ACTIVATION = 0, NB200_BN_SPATIAL = 1, NB200_BN_INSTANCE = 2 };
typedef struct nb200_bn_desc
{
    int32_t N, C, H, W;
    int32_t mode; /* NB200_BN_* */
} nb200_bn_desc;

/* Number of statistics (G) for this problem, and the device scratch every call below may use. */
NB200_API int32_t nb200_batch_norm_groups(const nb200_bn_desc* d);
NB200_API size_t nb200_batch_norm_workspace_bytes(const nb200_bn_desc* d);

/* y = (x - running_mean) / sqrt(running_var + eps) * gamma + beta. Replaces TensorOpCpu::BatchNormalization
 * (TensorOpCpu.h:55, TensorOpCpu.cpp:1371-1389). */
NB200_API int nb200_batch_norm(const nb200_bn_desc* d, const float* x, const float* gamma, const float* beta, float epsilon,
                               const float* running_mean, const float* running_var, float* y, void* stream);

/* Training forward. Replaces TensorOpCpu::BatchNormalizationTrain (TensorOpCpu.h:56, TensorOpCpu.cpp:1392-1434):
 * save_mean = mean(x), save_inv_var = 1/sqrt(var + eps) (biased variance), y = (x - mean) * inv * gamma + beta,
 * running_mean = (1-momentum)*running_mean + momentum*mean, running_var likewise with var*m/(m-1); running_* may be
 * NULL. One element per group (m == 1): y = x and nothing else is written, as in the reference. */
NB200_API int nb200_batch_norm_train(const nb200_bn_desc* d, const float* x, const float* gamma, const float* beta,
                                     float momentum, float epsilon, float* running_mean, float* running_var,
                                     float* save_mean, float* save_inv_var, float* y, void* workspace,
                                     size_t workspace_bytes, void* stream);

/* Gradient. Replaces TensorOpCpu::BatchNormalizationGradient (TensorOpCpu.h:57, TensorOpCpu.cpp:1437-1480; the
 * reference ignores `trainable` and `epsilon`, so they are not parameters). dgamma / dbeta may be NULL. */
NB200_API int nb200_batch_norm_gradient(const nb200_bn_desc* d, const float* x, const float* gamma, const float* dy,
                                        const float* save_mean, const float* save_inv_var, float* dgamma, float* dbeta,
                                        float* dx, void* workspace, size_t workspace_bytes, void* stream);

/* The same two ops for batch-sharded replicas (data-parallel Fit: every replica must normalise with the statistics
 * of the GLOBAL batch or the loss curve differs from the single-device run). The host inserts the exchange:
 *   forward : nb200_batch_norm_moments(x) -> moments[G][2] = (mean, sum of squared deviations) of the local shard;
 *             all-gather them (2*G floats per replica, rank order) -> all_moments[replicas][G][2];
 *             nb200_batch_norm_train_from_moments(all_moments, replicas, ...) combines them in rank order (Chan's
 *             formula; every replica derives bit-identical statistics) and normalises the local shard.
 *   backward: nb200_batch_norm_gradient_sums(x, dy, save_mean) -> sums[G][3] = (sum dy, sum dy*(x-mean), sum (x-mean));
 *             all-reduce(sum) a COPY -> global_sums;  nb200_batch_norm_gradient_from_sums(global_sums, local_sums, ...):
 *             dx uses the global sums and m = replicas * local elements per group; dgamma / dbeta are the LOCAL
 *             partial sums (they are reduced with the other parameter gradients).
 * All shards must have the same extent. replicas == 1 with all_moments = moments reproduces the single-device ops. */
NB200_API int nb200_batch_norm_moments(const nb200_bn_desc* d, const float* x, float* moments, void* workspace,
                                       size_t workspace_bytes, void* stream);
NB200_API int nb200_batch_norm_train_from_moments(const nb200_bn_desc* d, const float* all_moments, int32_t replicas,
                                                  const float* x, const float* gamma, const float* beta, float momentum,
                                                  float epsilon, float* running_mean, float* running_var, float* save_mean,
                                                  float* save_inv_var, float* y, void* stream);
NB200_API int nb200_batch_norm_gradient_sums(const nb200_bn_desc* d, const float* x, const float* dy, const float* save_mean,
                                             float* sums, void* workspace, size_t workspace_bytes, void* stream);
NB200_API int nb200_batch_norm_gradient_from_sums(const nb200_bn_desc* d, int32_t replicas, const float* global_sums,
                                                  const float* local_sums, const float* x, const float* gamma,
                                                  const float* dy, const float* save_mean, const float* save_inv_var,
                                                  float* dgamma, float* dbeta, float* dx, void* stream);

/* Host-buffer variants (pageable or pinned host memory in, host memory out): stage through device
 * buffers owned by the library on `stream`, run the op, copy the result back and wait for it. This is
 * what a caller holding host-resident Tensors (reference residency protocol, Storage.cpp:534-604)
 * pays per op; the TensorOpB200 C++ layer keeps intermediates device-resident instead. */
NB200_API int nb200_conv2d_forward_host(const nb200_conv_desc* d, const float* x, const float* w,
                                        const float* bias, int32_t act, float alpha, float* y, void* stream);
NB200_API int nb200_conv2d_input_gradient_host(const nb200_conv_desc* d, const float* dy, const float* w,
                                               float* dx, void* stream);
NB200_API int nb200_conv2d_kernels_gradient_host(const nb200_conv_desc* d, const float* x, const float* dy,
                                                 float* dw, float* db, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEURO_B200_H */
