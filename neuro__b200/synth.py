"""Counter-based synthetic data.

The reference fills test tensors with ``std::uniform_real_distribution`` over a ``mt19937``
(``Tensor::FillWithRand``, Neuro/src/Tensors/Tensor.cpp:239-247), whose stream differs between MSVC and
libstdc++. We therefore own the generator: element ``i`` of stream ``seed`` is a pure function of
``(seed, i)`` (splitmix64 finaliser), so the same host buffers can be rebuilt bit-for-bit on any box,
in any order, by any rank.
"""
import math

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def uniform(seed, shape, lo=-1.0, hi=1.0, offset=0):
    """float32 array of ``shape`` with element i = lo + (hi-lo) * u24(seed, offset+i) / 2**24."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        idx = np.arange(offset, offset + n, dtype=np.uint64)
        key = _splitmix64(np.uint64(seed) * np.uint64(0xD1342543DE82EF95) + np.uint64(0x2545F4914F6CDD1D))
        bits = _splitmix64(idx ^ key)
    u = (bits >> np.uint64(40)).astype(np.float64) * (1.0 / (1 << 24))
    return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


def glorot_uniform(seed, K, C, R, S):
    """Kernel init of the reference's Conv2D layer: GlorotUniform = VarianceScaling(1, fan_avg, uniform),
    fan_in = C*R*S, fan_out = K*R*S (Neuro/src/Initializers/VarianceScaling.cpp:40,59-65)."""
    limit = math.sqrt(6.0 / (C * R * S + K * R * S))
    return uniform(seed, (K, C, R, S), -limit, limit)


# Seeds mirror the reference's equivalence tests (Neuro.Tests/src/TensorOpGpuTests.cpp:1307-1310).
SEED_X, SEED_W, SEED_DY, SEED_BIAS, SEED_MODEL = 11, 12, 13, 14, 1337
