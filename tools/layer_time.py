"""Diagnostic: time one op of one layer with CUDA events. usage: layer_time.py op N C HW K [iters]"""
import sys
import torch
sys.path.insert(0, ".")
from neuro__b200 import lib  # noqa: E402
from neuro__b200.tensor_op import TensorOpB200  # noqa: E402
opn, N, C, HW, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 10
op = TensorOpB200(lib.MATH_TF32)
x = torch.randn(N, C, HW, HW, device="cuda"); w = torch.randn(K, C, 3, 3, device="cuda") * 0.05
y = torch.empty(N, K, HW, HW, device="cuda"); dy = torch.randn_like(y); dx = torch.empty_like(x); dw = torch.empty_like(w)
b = torch.zeros(K, device="cuda")
def run():
    if opn == "fwd":
        op.Conv2DBiasActivation(x, w, 1, 1, 1, b, lib.ACT_RELU, 0.0, y)
    elif opn == "dgrad":
        op.Conv2DInputGradient(dy, w, 1, 1, 1, lib.NCHW, dx)
    else:
        op.Conv2DKernelsGradient(x, dy, 1, 1, 1, lib.NCHW, dw)
for _ in range(3):
    run()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(iters):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("%s N=%d C=%d HW=%d K=%d: %.3f ms  %.1f TF/s" % (opn, N, C, HW, K, ms, 2.0 * N * C * K * 9 * HW * HW / ms * 1e-9))
