import sys, torch
sys.path.insert(0, ".")
from neuro__b200 import lib
from neuro__b200.tensor_op import TensorOpB200
op = TensorOpB200(lib.MATH_TF32)
for N in (8, 1):
    x = torch.randn(N, 3, 512, 512, device="cuda"); w = torch.randn(64, 3, 3, 3, device="cuda") * 0.05
    dy = torch.randn(N, 64, 512, 512, device="cuda"); dx = torch.empty_like(x)
    fn = lambda: op.Conv2DInputGradient(dy, w, 1, 1, 1, lib.NCHW, dx)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print("N=%d smallc_dgrad %.3f ms" % (N, e0.elapsed_time(e1) / 20))
    ref = dx.clone()
print(float(ref.abs().sum()))
