"""TEST INFRASTRUCTURE ONLY -- ctypes front-end to the CPU oracle.

Two libraries sit behind this module:

* ``oracle/_build/libconv_oracle.so``  our plain-C restatement (``oracle/conv_oracle.c``), always buildable;
* ``oracle/_ref/libneuro_ref.so``      the reference's own TensorOpCpu / TensorOpCpuMt sources compiled
  unmodified (``make -C oracle ref``; needs ``/root/reference`` at build time only -- the built .so
  travels to the GPU box).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module. Nothing under ``neuro__b200/`` does.

Array conventions (numpy, float32, C-contiguous):
  NCHW: x (N,C,H,W)   y (N,K,Ho,Wo)      NHWC: x (N,H,W,C)   y (N,Ho,Wo,K)      kernels: (K,C,R,S) in both.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "_build", "libconv_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libneuro_ref.so")

NCHW, NHWC = 0, 1
# EActivation numbering, Neuro/include/Types.h:83-92
IDENTITY, SIGMOID, RELU, TANH, ELU, LEAKY_RELU = 0, 1, 2, 3, 4, 5


class ConvDims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("N", "C", "H", "W", "K", "R", "S", "Ho", "Wo", "stride", "padX", "padY", "fmt")]


def build(ref=True):
    """Compile the oracle (and, when /root/reference is present, the reference build)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    if ref and os.path.isdir("/root/reference/Neuro/src/Tensors"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_ORACLE_SO):
            build(ref=False)
        _lib = ctypes.CDLL(_ORACLE_SO)
    return _lib


def have_ref():
    return os.path.exists(_REF_SO)


def ref():
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(_REF_SO)
        _ref.neuro_ref_threads.restype = ctypes.c_int
    return _ref


def conv_out_size(size, f, stride, pad):
    """Tensor::GetConvOutputShape, Neuro/src/Tensors/Tensor.cpp:2010-2029."""
    return (size + 2 * pad - f) // stride + 1


def conv_transpose_out_size(size, f, stride, pad):
    """Tensor::GetConvTransposeOutputShape, Neuro/src/Tensors/Tensor.cpp:2032-2051."""
    return (size - 1) * stride + f - 2 * pad


def _fp(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _xshape(fmt, a):
    """(N,C,H,W) extents of an activation array in either format."""
    if fmt == NCHW:
        n, c, h, w = a.shape
    else:
        n, h, w, c = a.shape
    return n, c, h, w


def _mk(fmt, n, c, h, w, k):
    return (n, k, h, w) if fmt == NCHW else (n, h, w, k)


def _dims(fmt, N, C, H, W, K, R, S, Ho, Wo, stride, padX, padY):
    return ConvDims(N, C, H, W, K, R, S, Ho, Wo, stride, padX, padY, fmt)


# ---------------------------------------------------------------- restatement

def conv2d(x, w, stride, padX, padY=None, fmt=NCHW, f64=False):
    padY = padX if padY is None else padY
    N, C, H, W = _xshape(fmt, x)
    K, C2, R, S = w.shape
    assert C2 == C
    Ho, Wo = conv_out_size(H, R, stride, padY), conv_out_size(W, S, stride, padX)
    y = np.empty(_mk(fmt, N, C, Ho, Wo, K), np.float32)
    d = _dims(fmt, N, C, H, W, K, R, S, Ho, Wo, stride, padX, padY)
    fn = lib().oracle_conv2d_f64 if f64 else lib().oracle_conv2d
    fn(ctypes.byref(d), _fp(x), _fp(w), _fp(y))
    return y


def conv2d_input_gradient(dy, w, stride, padX, padY, in_hw, fmt=NCHW, f64=False):
    """in_hw = (H, W) of the input gradient; the caller supplies it, as in the reference."""
    N, K, Ho, Wo = _xshape(fmt, dy)
    K2, C, R, S = w.shape
    assert K2 == K
    H, W = in_hw
    dx = np.empty(_mk(fmt, N, K, H, W, C), np.float32)
    d = _dims(fmt, N, C, H, W, K, R, S, Ho, Wo, stride, padX, padY)
    fn = lib().oracle_conv2d_input_gradient_f64 if f64 else lib().oracle_conv2d_input_gradient
    fn(ctypes.byref(d), _fp(dy), _fp(w), _fp(dx))
    return dx


def conv2d_kernels_gradient(x, dy, stride, padX, padY, filt_rs, fmt=NCHW, f64=False):
    """filt_rs = (R, S) of the kernel gradient; the caller supplies it, as in the reference."""
    N, C, H, W = _xshape(fmt, x)
    N2, K, Ho, Wo = _xshape(fmt, dy)
    assert N2 == N
    R, S = filt_rs
    dw = np.empty((K, C, R, S), np.float32)
    d = _dims(fmt, N, C, H, W, K, R, S, Ho, Wo, stride, padX, padY)
    fn = lib().oracle_conv2d_kernels_gradient_f64 if f64 else lib().oracle_conv2d_kernels_gradient
    fn(ctypes.byref(d), _fp(x), _fp(dy), _fp(dw))
    return dw


def conv2d_bias_activation(x, w, bias, stride, pad, act, alpha=0.0, fmt=NCHW):
    N, C, H, W = _xshape(fmt, x)
    K, _, R, S = w.shape
    Ho, Wo = conv_out_size(H, R, stride, pad), conv_out_size(W, S, stride, pad)
    y = np.empty(_mk(fmt, N, C, Ho, Wo, K), np.float32)
    d = _dims(fmt, N, C, H, W, K, R, S, Ho, Wo, stride, pad, pad)
    bias = np.ascontiguousarray(bias.reshape(-1), np.float32)
    lib().oracle_conv2d_bias_activation(ctypes.byref(d), _fp(x), _fp(w), _fp(bias), int(act),
                                        ctypes.c_float(alpha), _fp(y))
    return y


def conv2d_bias_gradient(dy, fmt=NCHW):
    N, K, Ho, Wo = _xshape(fmt, dy)
    db = np.empty((K,), np.float32)
    d = _dims(fmt, N, 0, 0, 0, K, 0, 0, Ho, Wo, 1, 0, 0)
    lib().oracle_conv2d_bias_gradient(ctypes.byref(d), _fp(dy), _fp(db))
    return db


def activation_gradient(act, alpha, y, dy):
    """dz = act'(y) * dy, derivative taken through the output y (TensorOpCpu.cpp:813-864)."""
    dz = np.empty_like(dy)
    lib().oracle_activation_gradient(int(act), ctypes.c_float(alpha), _fp(y), _fp(dy), _fp(dz), ctypes.c_size_t(dy.size))
    return dz


def adam_step(p, g, m, v, lr, beta1, beta2, eps):
    """In place on p, m, v."""
    lib().oracle_adam_step(_fp(p), _fp(g), _fp(m), _fp(v), ctypes.c_size_t(p.size), ctypes.c_float(lr),
                           ctypes.c_float(beta1), ctypes.c_float(beta2), ctypes.c_float(eps))


def sgd_step(p, g, lr):
    lib().oracle_sgd_step(_fp(p), _fp(g), ctypes.c_size_t(p.size), ctypes.c_float(lr))


# EBatchNormMode, Neuro/include/Types.h
PER_ACTIVATION, SPATIAL, INSTANCE = 0, 1, 2


def bn_layout(mode, shape):
    """(Nn, G, S) view of an NCHW array for the three normalisation modes (see conv_oracle.c)."""
    N, C, H, W = shape
    if mode == SPATIAL:
        return N, C, H * W
    if mode == PER_ACTIVATION:
        return N, C * H * W, 1
    return 1, N * C, H * W


def batch_norm_train(mode, x, gamma, beta, momentum, eps, running_mean=None, running_var=None):
    """Returns (y, save_mean, save_inv_var); running_mean / running_var (flat, G values) are updated in place."""
    Nn, G, S = bn_layout(mode, x.shape)
    y = np.empty_like(x); sm = np.zeros(G, np.float32); sv = np.zeros(G, np.float32)
    lib().oracle_batch_norm_train(Nn, G, S, _fp(x), _fp(gamma), _fp(beta), ctypes.c_float(momentum), ctypes.c_float(eps),
                                  _fp(running_mean) if running_mean is not None else None,
                                  _fp(running_var) if running_var is not None else None, _fp(sm), _fp(sv), _fp(y))
    return y, sm, sv


def batch_norm(mode, x, gamma, beta, eps, running_mean, running_var):
    Nn, G, S = bn_layout(mode, x.shape)
    y = np.empty_like(x)
    lib().oracle_batch_norm(Nn, G, S, _fp(x), _fp(gamma), _fp(beta), ctypes.c_float(eps), _fp(running_mean), _fp(running_var), _fp(y))
    return y


def batch_norm_gradient(mode, x, gamma, dy, save_mean, save_inv_var):
    """Returns (dx, dgamma, dbeta)."""
    Nn, G, S = bn_layout(mode, x.shape)
    dx = np.empty_like(x); dg = np.zeros(G, np.float32); db = np.zeros(G, np.float32)
    lib().oracle_batch_norm_gradient(Nn, G, S, _fp(x), _fp(gamma), _fp(dy), _fp(save_mean), _fp(save_inv_var), _fp(dg), _fp(db), _fp(dx))
    return dx, dg, db


MAX_POOL, AVG_POOL = 0, 1


class PoolDims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("N", "C", "H", "W", "Ho", "Wo", "filter", "stride", "padX", "padY", "mode", "fmt")]


def pool_out_size(size, f, stride, pad):
    """Tensor::GetPooling2DOutputShape, Neuro/src/Tensors/Tensor.cpp:1988-2007."""
    return (size + 2 * pad - f) // stride + 1


def _pool_dims(x, f, stride, mode, padX, padY, fmt):
    N, C, H, W = _xshape(fmt, x)
    return PoolDims(N, C, H, W, pool_out_size(H, f, stride, padY), pool_out_size(W, f, stride, padX), f, stride, padX, padY, mode, fmt)


def pool2d(x, f, stride, mode, padX=0, padY=None, fmt=NCHW):
    padY = padX if padY is None else padY
    d = _pool_dims(x, f, stride, mode, padX, padY, fmt)
    y = np.empty(_mk(fmt, d.N, 0, d.Ho, d.Wo, d.C), np.float32)
    lib().oracle_pool2d(ctypes.byref(d), _fp(x), _fp(y))
    return y


def pool2d_gradient(y, x, dy, f, stride, mode, padX=0, padY=None, fmt=NCHW):
    padY = padX if padY is None else padY
    d = _pool_dims(x, f, stride, mode, padX, padY, fmt)
    dx = np.empty_like(x)
    lib().oracle_pool2d_gradient(ctypes.byref(d), _fp(y), _fp(x), _fp(dy), _fp(dx))
    return dx


def upsample2d(x, s):
    N, C, H, W = x.shape
    y = np.empty((N, C, H * s, W * s), np.float32)
    lib().oracle_upsample2d(N * C, H, W, s, _fp(x), _fp(y))
    return y


def upsample2d_gradient(dy, s):
    N, C, Ho, Wo = dy.shape
    dx = np.empty((N, C, Ho // s, Wo // s), np.float32)
    lib().oracle_upsample2d_gradient(N * C, Ho // s, Wo // s, s, _fp(dy), _fp(dx))
    return dx


def constant_pad2d(x, left, right, top, bottom, value):
    N, C, H, W = x.shape
    y = np.empty((N, C, H + top + bottom, W + left + right), np.float32)
    lib().oracle_constant_pad2d(N * C, H, W, left, right, top, bottom, ctypes.c_float(value), _fp(x), _fp(y))
    return y


# ---------------------------------------------------------------- the reference itself

def _shape4(fmt, a, kernels=False):
    """Reference Shape(d0..d3), fastest dimension first."""
    s = a.shape
    return (ctypes.c_uint32 * 4)(s[3], s[2], s[1], s[0])


def ref_threads():
    return ref().neuro_ref_threads()


def ref_set_threads(n=0):
    ref().neuro_ref_set_threads(int(n))


def ref_conv2d(x, w, stride, padX, padY=None, fmt=NCHW, mt=False):
    padY = padX if padY is None else padY
    N, C, H, W = _xshape(fmt, x)
    K, _, R, S = w.shape
    Ho, Wo = conv_out_size(H, R, stride, padY), conv_out_size(W, S, stride, padX)
    y = np.empty(_mk(fmt, N, C, Ho, Wo, K), np.float32)
    ref().neuro_ref_conv2d(int(mt), fmt, _fp(x), _shape4(fmt, x), _fp(w), _shape4(fmt, w), stride, padX, padY,
                           _fp(y), _shape4(fmt, y))
    return y


def ref_conv2d_input_gradient(dy, w, stride, padX, padY, in_hw, fmt=NCHW, mt=False):
    N, K, Ho, Wo = _xshape(fmt, dy)
    C = w.shape[1]
    dx = np.empty(_mk(fmt, N, K, in_hw[0], in_hw[1], C), np.float32)
    ref().neuro_ref_conv2d_input_gradient(int(mt), fmt, _fp(dy), _shape4(fmt, dy), _fp(w), _shape4(fmt, w),
                                          stride, padX, padY, _fp(dx), _shape4(fmt, dx))
    return dx


def ref_conv2d_kernels_gradient(x, dy, stride, padX, padY, filt_rs, fmt=NCHW, mt=False):
    N, C, H, W = _xshape(fmt, x)
    K = _xshape(fmt, dy)[1]
    dw = np.empty((K, C, filt_rs[0], filt_rs[1]), np.float32)
    ref().neuro_ref_conv2d_kernels_gradient(int(mt), fmt, _fp(x), _shape4(fmt, x), _fp(dy), _shape4(fmt, dy),
                                            stride, padX, padY, _fp(dw), _shape4(fmt, dw))
    return dw


def ref_activation_gradient(act, alpha, y, dy):
    dz = np.empty_like(dy)
    a = y.reshape(-1)
    dims = (ctypes.c_uint32 * 4)(a.size, 1, 1, 1)
    ref().neuro_ref_activation_gradient(int(act), ctypes.c_float(alpha), _fp(y), _fp(dy), _fp(dz), dims)
    return dz


def ref_pool2d(x, f, stride, mode, padX=0, padY=None, fmt=NCHW):
    padY = padX if padY is None else padY
    d = _pool_dims(x, f, stride, mode, padX, padY, fmt)
    y = np.empty(_mk(fmt, d.N, 0, d.Ho, d.Wo, d.C), np.float32)
    ref().neuro_ref_pool2d(fmt, _fp(x), _shape4(fmt, x), f, stride, mode, padX, padY, _fp(y), _shape4(fmt, y))
    return y


def ref_pool2d_gradient(y, x, dy, f, stride, mode, padX=0, padY=None, fmt=NCHW):
    padY = padX if padY is None else padY
    dx = np.empty_like(x)
    ref().neuro_ref_pool2d_gradient(fmt, _fp(y), _shape4(fmt, y), _fp(x), _shape4(fmt, x), _fp(dy), f, stride, mode, padX, padY, _fp(dx))
    return dx


def ref_upsample2d(x, s):
    N, C, H, W = x.shape
    y = np.empty((N, C, H * s, W * s), np.float32)
    ref().neuro_ref_upsample2d(_fp(x), _shape4(NCHW, x), s, _fp(y), _shape4(NCHW, y))
    return y


def ref_upsample2d_gradient(dy, s):
    N, C, Ho, Wo = dy.shape
    dx = np.empty((N, C, Ho // s, Wo // s), np.float32)
    ref().neuro_ref_upsample2d_gradient(_fp(dy), _shape4(NCHW, dy), s, _fp(dx), _shape4(NCHW, dx))
    return dx


def ref_constant_pad2d(x, left, right, top, bottom, value):
    N, C, H, W = x.shape
    y = np.empty((N, C, H + top + bottom, W + left + right), np.float32)
    ref().neuro_ref_constant_pad2d(_fp(x), _shape4(NCHW, x), left, right, top, bottom, ctypes.c_float(value), _fp(y), _shape4(NCHW, y))
    return y


def ref_adam_step(p, g, m, v, lr, beta1, beta2, eps):
    """TensorOpCpu::AdamStep itself (TensorOpCpu.cpp:987-1003), in place on p, m, v."""
    ref().neuro_ref_adam_step(_fp(p), _fp(g), _fp(m), _fp(v), ctypes.c_uint32(p.size), ctypes.c_float(lr), ctypes.c_float(beta1),
                              ctypes.c_float(beta2), ctypes.c_float(eps))


def ref_sgd_step(p, g, lr):
    ref().neuro_ref_sgd_step(_fp(p), _fp(g), ctypes.c_uint32(p.size), ctypes.c_float(lr))


def _bn_param_dims(mode, shape):
    """Reference Shape of gamma / beta / statistics for an NCHW input (BatchNormalization.cpp Build; InstanceNormalize op)."""
    N, C, H, W = shape
    if mode == SPATIAL:
        return (ctypes.c_uint32 * 4)(1, 1, C, 1)
    if mode == PER_ACTIVATION:
        return (ctypes.c_uint32 * 4)(W, H, C, 1)
    return (ctypes.c_uint32 * 4)(1, 1, C, N)


def ref_batch_norm_train(mode, x, gamma, beta, momentum, eps, running_mean=None, running_var=None):
    G = gamma.size
    y = np.empty_like(x); sm = np.zeros(G, np.float32); sv = np.zeros(G, np.float32)
    ref().neuro_ref_batch_norm_train(int(mode), _fp(x), _shape4(NCHW, x), _fp(gamma), _fp(beta), _bn_param_dims(mode, x.shape),
                                     ctypes.c_float(momentum), ctypes.c_float(eps),
                                     _fp(running_mean) if running_mean is not None else None,
                                     _fp(running_var) if running_var is not None else None, _fp(sm), _fp(sv), _fp(y))
    return y, sm, sv


def ref_batch_norm(mode, x, gamma, beta, eps, running_mean, running_var):
    y = np.empty_like(x)
    ref().neuro_ref_batch_norm(int(mode), _fp(x), _shape4(NCHW, x), _fp(gamma), _fp(beta), _bn_param_dims(mode, x.shape), ctypes.c_float(eps),
                               _fp(running_mean), _fp(running_var), _fp(y))
    return y


def ref_batch_norm_gradient(mode, x, gamma, eps, dy, save_mean, save_inv_var):
    G = gamma.size
    dx = np.empty_like(x); dg = np.zeros(G, np.float32); db = np.zeros(G, np.float32)
    ref().neuro_ref_batch_norm_gradient(int(mode), _fp(x), _shape4(NCHW, x), _fp(gamma), _bn_param_dims(mode, x.shape), ctypes.c_float(eps),
                                        _fp(dy), _fp(save_mean), _fp(save_inv_var), _fp(dg), _fp(db), 1, _fp(dx))
    return dx, dg, db
