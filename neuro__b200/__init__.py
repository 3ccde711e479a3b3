"""neuro__b200 -- B200-native Conv2D backend for Neuro_ (hot path only).

Layout:  csrc/        CUDA kernels + the C ABI (include/neuro_b200.h) -> libneuro_b200.so
         lib.py       ctypes binding of the C ABI
         tensor_op.py TensorOpB200: host-side mirror of the reference's conv op interface
         synth.py     counter-based synthetic inputs
"""
from . import lib, synth  # noqa: F401
