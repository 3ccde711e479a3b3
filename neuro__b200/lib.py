"""ctypes binding of the C ABI in include/neuro_b200.h (libneuro_b200.so, built in-tree by build.py).

There is no fallback: if the shared library is missing, or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libneuro_b200.so")

NCHW, NHWC = 0, 1
ACT_IDENTITY, ACT_SIGMOID, ACT_RELU, ACT_TANH, ACT_ELU, ACT_LEAKY_RELU = 0, 1, 2, 3, 4, 5
MATH_TF32, MATH_3XTF32, MATH_FP32 = 0, 1, 2
OP_FORWARD, OP_INPUT_GRADIENT, OP_KERNELS_GRADIENT = 0, 1, 2

# every symbol include/neuro_b200.h declares (tests check the .so exports exactly these)
EXPORTS = [
    "nb200_version", "nb200_last_error", "nb200_kernel_launches", "nb200_device_info", "nb200_padding", "nb200_conv_out_size",
    "nb200_conv_transpose_out_size", "nb200_conv2d_workspace_bytes", "nb200_conv2d_kernel_name",
    "nb200_conv2d_forward", "nb200_conv2d_input_gradient", "nb200_conv2d_kernels_gradient",
    "nb200_conv2d_bias_gradient", "nb200_adam_step", "nb200_sgd_step", "nb200_conv2d_forward_host",
    "nb200_conv2d_input_gradient_host", "nb200_conv2d_kernels_gradient_host",
    "nb200_conv2d_bias_activation_gradient_workspace_bytes", "nb200_conv2d_bias_activation_gradient",
    "nb200_conv2d_prepare_filters", "nb200_conv2d_forward_prepared", "nb200_conv2d_input_gradient_prepared",
    "nb200_pool2d", "nb200_pool2d_gradient", "nb200_upsample2d", "nb200_upsample2d_gradient", "nb200_constant_pad2d",
    "nb200_conv2d_plan_create", "nb200_conv2d_plan_run", "nb200_conv2d_plan_kernels", "nb200_conv2d_plan_destroy",
    "nb200_bias_activation", "nb200_pool2d_gradient_activation_supported", "nb200_pool2d_gradient_activation_workspace_bytes",
    "nb200_pool2d_gradient_activation", "nb200_batch_norm_groups", "nb200_batch_norm_workspace_bytes", "nb200_batch_norm", "nb200_batch_norm_train",
    "nb200_batch_norm_gradient", "nb200_batch_norm_moments", "nb200_batch_norm_train_from_moments",
    "nb200_batch_norm_gradient_sums", "nb200_batch_norm_gradient_from_sums",
]


class ConvDesc(ctypes.Structure):
    """struct nb200_conv_desc"""
    _fields_ = [(n, ctypes.c_int32) for n in
                ("N", "C", "H", "W", "K", "R", "S", "Ho", "Wo", "stride", "padX", "padY", "fmt", "math")]

    def flops(self):
        """2*N*K*Ho*Wo*C*R*S -- the dense FLOP convention shared by fwd, dgrad and wgrad."""
        return 2.0 * self.N * self.K * self.Ho * self.Wo * self.C * self.R * self.S

    def bytes(self):
        """4*(|x|+|w|+|y|): each tensor touched once."""
        return 4.0 * (self.N * self.C * self.H * self.W + self.K * self.C * self.R * self.S
                      + self.N * self.K * self.Ho * self.Wo)


POOL_MAX, POOL_AVG = 0, 1


class PoolDesc(ctypes.Structure):
    """struct nb200_pool_desc"""
    _fields_ = [(n, ctypes.c_int32) for n in ("N", "C", "H", "W", "Ho", "Wo", "filter", "stride", "padX", "padY", "mode", "fmt")]


BN_PER_ACTIVATION, BN_SPATIAL, BN_INSTANCE = 0, 1, 2


class BnDesc(ctypes.Structure):
    """struct nb200_bn_desc"""
    _fields_ = [(n, ctypes.c_int32) for n in ("N", "C", "H", "W", "mode")]


class NeuroB200Error(RuntimeError):
    pass


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise NeuroB200Error(
            "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)" % SO_PATH)
    L = ctypes.CDLL(SO_PATH)
    c_p, c_f, c_i, c_sz = ctypes.c_void_p, ctypes.c_float, ctypes.c_int32, ctypes.c_size_t
    dp = ctypes.POINTER(ConvDesc)
    L.nb200_version.restype = ctypes.c_char_p
    L.nb200_last_error.restype = ctypes.c_char_p
    L.nb200_kernel_launches.restype = ctypes.c_ulonglong
    L.nb200_device_info.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                    ctypes.POINTER(ctypes.c_int), ctypes.POINTER(c_sz)]
    for f in (L.nb200_padding,):
        f.argtypes = [c_i, c_i]; f.restype = c_i
    for f in (L.nb200_conv_out_size, L.nb200_conv_transpose_out_size):
        f.argtypes = [c_i, c_i, c_i, c_i]; f.restype = c_i
    L.nb200_conv2d_workspace_bytes.argtypes = [c_i, dp]; L.nb200_conv2d_workspace_bytes.restype = c_sz
    L.nb200_conv2d_kernel_name.argtypes = [c_i, dp]; L.nb200_conv2d_kernel_name.restype = ctypes.c_char_p
    L.nb200_conv2d_forward.argtypes = [dp, c_p, c_p, c_p, c_i, c_f, c_p, c_p, c_sz, c_p]
    L.nb200_conv2d_input_gradient.argtypes = [dp, c_p, c_p, c_p, c_p, c_sz, c_p]
    L.nb200_conv2d_kernels_gradient.argtypes = [dp, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]
    L.nb200_conv2d_bias_gradient.argtypes = [dp, c_p, c_p, c_p]
    L.nb200_adam_step.argtypes = [c_p, c_p, c_p, c_p, c_sz, c_f, c_f, c_f, c_f, c_f, c_p]
    L.nb200_sgd_step.argtypes = [c_p, c_p, c_sz, c_f, c_f, c_p]
    L.nb200_conv2d_forward_host.argtypes = [dp, c_p, c_p, c_p, c_i, c_f, c_p, c_p]
    L.nb200_conv2d_input_gradient_host.argtypes = [dp, c_p, c_p, c_p, c_p]
    L.nb200_conv2d_kernels_gradient_host.argtypes = [dp, c_p, c_p, c_p, c_p, c_p]
    L.nb200_conv2d_bias_activation_gradient_workspace_bytes.argtypes = [dp]
    L.nb200_conv2d_bias_activation_gradient_workspace_bytes.restype = c_sz
    L.nb200_conv2d_bias_activation_gradient.argtypes = [dp, c_i, c_f, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]
    L.nb200_conv2d_plan_create.argtypes = [c_i, dp, c_p, c_p, c_p, c_p, c_i, c_f, c_i, c_p, c_sz, ctypes.POINTER(c_p)]
    L.nb200_conv2d_plan_run.argtypes = [c_p, c_p]
    L.nb200_conv2d_plan_kernels.argtypes = [c_p]; L.nb200_conv2d_plan_kernels.restype = c_i
    L.nb200_conv2d_plan_destroy.argtypes = [c_p]; L.nb200_conv2d_plan_destroy.restype = None
    L.nb200_bias_activation.argtypes = [dp, c_p, c_p, c_i, c_f, c_p, c_p]
    L.nb200_conv2d_prepare_filters.argtypes = [c_i, dp, c_p, c_p, c_sz, c_p]
    L.nb200_conv2d_forward_prepared.argtypes = L.nb200_conv2d_forward.argtypes
    L.nb200_conv2d_input_gradient_prepared.argtypes = L.nb200_conv2d_input_gradient.argtypes
    pp = ctypes.POINTER(PoolDesc)
    L.nb200_pool2d.argtypes = [pp, c_p, c_p, c_p]
    L.nb200_pool2d_gradient.argtypes = [pp, c_p, c_p, c_p, c_p, c_p]
    L.nb200_pool2d_gradient_activation_supported.argtypes = [pp]; L.nb200_pool2d_gradient_activation_supported.restype = c_i
    L.nb200_pool2d_gradient_activation_workspace_bytes.argtypes = [pp]; L.nb200_pool2d_gradient_activation_workspace_bytes.restype = c_sz
    L.nb200_pool2d_gradient_activation.argtypes = [pp, c_i, c_f, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]
    L.nb200_upsample2d.argtypes = [c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]
    L.nb200_upsample2d_gradient.argtypes = [c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]
    L.nb200_constant_pad2d.argtypes = [c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p, c_p, c_p]
    bp = ctypes.POINTER(BnDesc)
    L.nb200_batch_norm_groups.argtypes = [bp]; L.nb200_batch_norm_groups.restype = c_i
    L.nb200_batch_norm_workspace_bytes.argtypes = [bp]; L.nb200_batch_norm_workspace_bytes.restype = c_sz
    L.nb200_batch_norm.argtypes = [bp, c_p, c_p, c_p, c_f, c_p, c_p, c_p, c_p]
    L.nb200_batch_norm_train.argtypes = [bp, c_p, c_p, c_p, c_f, c_f, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]
    L.nb200_batch_norm_gradient.argtypes = [bp, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]
    L.nb200_batch_norm_moments.argtypes = [bp, c_p, c_p, c_p, c_sz, c_p]
    L.nb200_batch_norm_train_from_moments.argtypes = [bp, c_p, c_i, c_p, c_p, c_p, c_f, c_f, c_p, c_p, c_p, c_p, c_p, c_p]
    L.nb200_batch_norm_gradient_sums.argtypes = [bp, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]
    L.nb200_batch_norm_gradient_from_sums.argtypes = [bp, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise NeuroB200Error("nb200 call failed (%d): %s" % (rc, load().nb200_last_error().decode()))
