"""Diagnostic: where the per-layer time floor of the gathered / halo-tile kernels comes from.

Times forward, input gradient and kernel gradient of synthetic layers whose ITERATION COUNT per CTA (channel blocks x taps)
varies while the grid stays the same, replayed as a CUDA graph of 20 calls (GPU time only). A straight-line fit over the
iteration count separates launch + prologue + epilogue (intercept) from the per-iteration chain (slope).

  python tools/floor_probe.py            # the table
  NB200_GATHER_DEBUG=<flags> python tools/floor_probe.py ablate
"""
import os
import sys
import torch
sys.path.insert(0, ".")
from neuro__b200 import lib  # noqa: E402
from neuro__b200.tensor_op import TensorOpB200  # noqa: E402

op = TensorOpB200(lib.MATH_TF32)
ITERS = 20


def graph_ms(fn):
    fn(); fn(); torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(ITERS):
            fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / ITERS)
    return best


def probe(tag, N, C, H, K, F, st, p):
    Ho = (H + 2 * p - F) // st + 1
    x = torch.randn(N, C, H, H, device="cuda"); w = torch.randn(K, C, F, F, device="cuda") * 0.05
    y = torch.empty(N, K, Ho, Ho, device="cuda"); dy = torch.randn_like(y); dx = torch.empty_like(x); dw = torch.empty_like(w)
    d = lib.ConvDesc(N, C, H, H, K, F, F, Ho, Ho, st, p, p, lib.NCHW, lib.MATH_TF32)
    pf = op.PrepareKernels(lib.OP_FORWARD, x, w, y, st, p, p)
    pg = op.PrepareKernels(lib.OP_INPUT_GRADIENT, dx, w, dy, st, p, p)
    t = [graph_ms(lambda: op.Conv2D(x, w, st, p, p, lib.NCHW, y, prepared=pf)),
         graph_ms(lambda: op.Conv2DInputGradient(dy, w, st, p, p, lib.NCHW, dx, prepared=pg)),
         graph_ms(lambda: op.Conv2DKernelsGradient(x, dy, st, p, p, lib.NCHW, dw))]
    iters = ((C + 31) // 32) * F * F
    print("%-34s iters/CTA %4d  %5.2f GF | %-22s %7.1f us | %-22s %7.1f us | %-22s %7.1f us" %
          (tag, iters, d.flops() / 1e9, op.kernel_name(0, d), t[0] * 1e3, op.kernel_name(1, d), t[1] * 1e3, op.kernel_name(2, d), t[2] * 1e3), flush=True)


def run_table(which):
    # empty-kernel floor of a graph node on this box
    a = torch.zeros(1024, device="cuda")
    print("graph node floor (torch fill_ of 4 KB): %.1f us" % (graph_ms(lambda: a.fill_(1.0)) * 1e3))
    if which == "ablate":
        # attribute the gathered kernel's fixed cost: run once per NB200_GATHER_DEBUG value (read once per process):
        # 0 full | 1 no gather loads | 2 no output stores | 3 neither | 4 empty body (launch + barrier init + TMEM alloc/free) | 12 no TMEM either
        print("NB200_GATHER_DEBUG=%s" % os.environ.get("NB200_GATHER_DEBUG", "0"))
        probe("gather s2 @8 N128 C32 K128 1x1", 128, 32, 8, 128, 1, 2, 0)
        probe("gather s2 @8 N128 C128 K128 3x3", 128, 128, 8, 128, 3, 2, 1)
        probe("gather s2 @32 N128 C128 K128 3x3", 128, 128, 32, 128, 3, 2, 1)
        return
    # gathered kernel, fixed grid (M = 128*4*4 = 2048 pixels = 16 tiles), growing reduction
    for C in (32, 64, 128, 256, 512):
        probe("gather s2 @8 N128 C%d K128 3x3" % C, 128, C, 8, 128, 3, 2, 1)
    for C in (32, 128, 512):
        probe("gather s2 @8 N128 C%d K128 1x1" % C, 128, C, 8, 128, 1, 2, 0)
    # gathered kernel, a full wave (M = 128*16*16 = 32768 pixels = 256 tiles)
    for C in (32, 128, 512):
        probe("gather s2 @32 N128 C%d K128 3x3" % C, 128, C, 32, 128, 3, 2, 1)
    # halo-tile kernel at batch 1 (style transfer): 64x64 map = 32 tiles per filter tile
    for C in (32, 128, 512):
        probe("halo s1 @64 N1 C%d K256 3x3" % C, 1, C, 64, 256, 3, 1, 1)
    for C in (32, 128, 512):
        probe("halo s1 @64 N1 C%d K256 1x1" % C, 1, C, 64, 256, 1, 1, 0)
    # halo-tile kernel, full waves
    for C in (32, 128, 512):
        probe("halo s1 @128 N4 C%d K256 3x3" % C, 4, C, 128, 256, 3, 1, 1)


if __name__ == "__main__":
    run_table(sys.argv[1] if len(sys.argv) > 1 else "all")
