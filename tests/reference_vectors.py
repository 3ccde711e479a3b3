"""The literal known-answer vectors of the reference's forward-convolution tests.

Source: Neuro.Tests/src/TensorTests.cpp:352-425 (Conv2D_Valid_1Kernel_1Batch, Conv2D_Valid_3Kernels_1Batch,
Conv2D_Valid_2Kernels_2Batches, Conv2D_Same_1Kernel_1Batch, Conv2D_Full_1Kernel_1Batch). Inputs are
FillWithRange(0) = 0,1,2,... (Tensor.cpp:261-267): 6x6x2 input (x N), 3x3x2 kernels (x K), stride 1.
Each entry: (name, N, K, pad, expected outputs flattened in NCHW order).
"""
import numpy as np

_V1 = [5511, 5664, 5817, 5970, 6429, 6582, 6735, 6888, 7347, 7500, 7653, 7806, 8265, 8418, 8571, 8724]
_V3 = _V1 + [13611, 14088, 14565, 15042, 16473, 16950, 17427, 17904, 19335, 19812, 20289, 20766, 22197, 22674, 23151,
             23628, 21711, 22512, 23313, 24114, 26517, 27318, 28119, 28920, 31323, 32124, 32925, 33726, 36129, 36930,
             37731, 38532]
_V22 = _V3[:32] + [16527, 16680, 16833, 16986, 17445, 17598, 17751, 17904, 18363, 18516, 18669, 18822, 19281, 19434,
                   19587, 19740, 47955, 48432, 48909, 49386, 50817, 51294, 51771, 52248, 53679, 54156, 54633, 55110,
                   56541, 57018, 57495, 57972]
_SAME = [2492, 3674, 3794, 3914, 4034, 2624, 3765, 5511, 5664, 5817, 5970, 3855, 4413, 6429, 6582, 6735, 6888, 4431,
         5061, 7347, 7500, 7653, 7806, 5007, 5709, 8265, 8418, 8571, 8724, 5583, 3416, 4898, 4982, 5066, 5150, 3260]
_FULL = [612, 1213, 1801, 1870, 1939, 2008, 1315, 645, 1266, 2492, 3674, 3794, 3914, 4034, 2624, 1278, 1926, 3765,
         5511, 5664, 5817, 5970, 3855, 1863, 2268, 4413, 6429, 6582, 6735, 6888, 4431, 2133, 2610, 5061, 7347, 7500,
         7653, 7806, 5007, 2403, 2952, 5709, 8265, 8418, 8571, 8724, 5583, 2673, 1782, 3416, 4898, 4982, 5066, 5150,
         3260, 1542, 786, 1489, 2107, 2140, 2173, 2206, 1375, 639]

KNOWN_ANSWERS = [
    ("valid_1kernel_1batch", 1, 1, 0, _V1),
    ("valid_3kernels_1batch", 1, 3, 0, _V3),
    ("valid_2kernels_2batches", 2, 2, 0, _V22),
    ("same_1kernel_1batch", 1, 1, 1, _SAME),
    ("full_1kernel_1batch", 1, 1, 2, _FULL),
]


def known_answer_inputs(N, K):
    x = np.arange(N * 2 * 6 * 6, dtype=np.float32).reshape(N, 2, 6, 6)
    w = np.arange(K * 2 * 3 * 3, dtype=np.float32).reshape(K, 2, 3, 3)
    return x, w
