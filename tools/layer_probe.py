"""Diagnostic: per-layer times of the three conv ops on the layer shapes of a BASELINE config.

  python tools/layer_probe.py dcgan|pix2pix|vgg1|autoenc [--prepared] [--graph]

Rows: conv y = conv(x[N,C,H,H], w[K,C,F,F], stride, pad). Conv2DTranspose layers are listed as the conv whose input
gradient is their forward. --prepared times forward / input gradient with filters prepared once (constant weights)."""
import sys
import torch
sys.path.insert(0, ".")
from neuro__b200 import lib  # noqa: E402
from neuro__b200.tensor_op import TensorOpB200  # noqa: E402

CONFIGS = {
    # CifarGAN, batch 128 (CifarGAN.cpp:11-36)
    "dcgan": (128, [("D conv1 3->64 s2", 3, 32, 64, 3, 2, 1), ("D conv2 64->128 s2", 64, 16, 128, 3, 2, 1),
                    ("D conv3 128->128 s2", 128, 8, 128, 3, 2, 1), ("D conv4 128->256 s1", 128, 4, 256, 3, 1, 1),
                    ("G deconv1 (256x4x4 -> 128x8x8)", 128, 8, 256, 4, 2, 1), ("G deconv2 (128x8x8 -> 128x16x16)", 128, 16, 128, 4, 2, 1),
                    ("G deconv3 (128x16x16 -> 128x32x32)", 128, 32, 128, 4, 2, 1), ("G conv out 128->3 s1", 128, 32, 3, 3, 1, 1)]),
    # pix2pix U-Net generator + PatchGAN discriminator, 256x256, batch 8 (Pix2Pix.cpp:4-109)
    "pix2pix": (8, [("G enc1 3->64 s2", 3, 256, 64, 3, 2, 1), ("G enc2 64->128 s2", 64, 128, 128, 3, 2, 1), ("G enc3 128->256 s2", 128, 64, 256, 3, 2, 1),
                    ("G enc4 256->512 s2", 256, 32, 512, 3, 2, 1), ("G enc5 512->512 s2 @16", 512, 16, 512, 3, 2, 1), ("G enc6 @8", 512, 8, 512, 3, 2, 1),
                    ("G enc7 @4", 512, 4, 512, 3, 2, 1), ("G enc8 @2", 512, 2, 512, 3, 2, 1),
                    ("G dec1 512->512 @2", 512, 2, 512, 3, 1, 1), ("G dec2 1024->512 @4", 1024, 4, 512, 3, 1, 1), ("G dec3 1024->512 @8", 1024, 8, 512, 3, 1, 1),
                    ("G dec4 1024->256 @16", 1024, 16, 256, 3, 1, 1), ("G dec5 768->128 @32", 768, 32, 128, 3, 1, 1), ("G dec6 384->64 @64", 384, 64, 64, 3, 1, 1),
                    ("G dec7 192->64 @128", 192, 128, 64, 3, 1, 1), ("G last 128->3 @256", 128, 256, 3, 3, 1, 1),
                    ("D 6->64 s2 @259", 6, 259, 64, 4, 2, 0), ("D 64->128 s2 @131", 64, 131, 128, 4, 2, 0), ("D 128->256 s2 @67", 128, 67, 256, 4, 2, 0),
                    ("D 256->512 s1 @34", 256, 34, 512, 4, 1, 0), ("D 512->1 s1 @33", 512, 33, 1, 4, 1, 0)]),
    # VGG16 @512, batch 1 (style transfer)
    "vgg1": (1, [("conv %d->%d @%d" % (c, k, h), c, h, k, 3, 1, 1) for (c, k, h) in
                 [(3, 64, 512), (64, 64, 512), (64, 128, 256), (128, 128, 256), (128, 256, 128), (256, 256, 128), (256, 512, 64), (512, 512, 64), (512, 512, 32)]]),
    # conv autoencoder, batch 256 (ConvAutoencoderNetwork.h:25-35)
    "autoenc": (256, [("enc1 1->16 @28", 1, 28, 16, 3, 1, 1), ("enc2 16->8 @14", 16, 14, 8, 3, 1, 1), ("dec1 8->8 @7", 8, 7, 8, 3, 1, 1),
                      ("dec2 8->16 @14", 8, 14, 16, 3, 1, 1), ("dec3 16->1 @28", 16, 28, 1, 3, 1, 1)]),
}
name = sys.argv[1] if len(sys.argv) > 1 else "dcgan"
prepared = "--prepared" in sys.argv
graph = "--graph" in sys.argv
N, LAYERS = CONFIGS[name]
op = TensorOpB200(lib.MATH_TF32)
tot = [0.0, 0.0, 0.0]; totfl = 0.0
for (lname, C, H, K, F, st, p) in LAYERS:
    Ho = (H + 2 * p - F) // st + 1
    x = torch.randn(N, C, H, H, device="cuda"); w = torch.randn(K, C, F, F, device="cuda") * 0.05
    y = torch.empty(N, K, Ho, Ho, device="cuda"); dy = torch.randn_like(y); dx = torch.empty_like(x); dw = torch.empty_like(w)
    d = lib.ConvDesc(N, C, H, H, K, F, F, Ho, Ho, st, p, p, lib.NCHW, lib.MATH_TF32)
    pf = op.PrepareKernels(lib.OP_FORWARD, x, w, y, st, p, p) if prepared else None
    pg = op.PrepareKernels(lib.OP_INPUT_GRADIENT, dx, w, dy, st, p, p) if prepared else None
    fns = [lambda: op.Conv2D(x, w, st, p, p, lib.NCHW, y, prepared=pf), lambda: op.Conv2DInputGradient(dy, w, st, p, p, lib.NCHW, dx, prepared=pg),
           lambda: op.Conv2DKernelsGradient(x, dy, st, p, p, lib.NCHW, dw)]
    out = []
    for i, fn in enumerate(fns):
        fn(); fn(); torch.cuda.synchronize()
        run = lambda: [fn() for _ in range(10)]
        if graph:   # GPU time without the host's per-call cost (ctypes, tensor-map encode, launches): 10 calls replayed as one CUDA graph
            side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                run()
            run = gr.replay
            run(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tot[i] += ms
        out.append("%s %7.3f ms %6.1f TF/s %6.0f GB/s" % (op.kernel_name(i, d)[:20], ms, d.flops() / ms / 1e9, d.bytes() / ms / 1e6))
    totfl += d.flops()
    print("%-36s %6.2f GF | " % (lname, d.flops() / 1e9) + " | ".join(out), flush=True)
print("TOTAL %s N=%d prepared=%d graph=%d: fwd %.3f ms dgrad %.3f ms wgrad %.3f ms; %.1f GF per op -> %.1f / %.1f / %.1f TF/s" %
      (name, N, prepared, graph, tot[0], tot[1], tot[2], totfl / 1e9, totfl / tot[0] / 1e9, totfl / tot[1] / 1e9, totfl / tot[2] / 1e9))
