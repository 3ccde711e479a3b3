"""Batch-sharded data-parallel training of a convolution stack -- the Fit() slice of the hot path.

Mirrors what the reference does per training step, restricted to what lies on the path:

    ModelBase::Fit / TrainStep        Neuro/src/Models/ModelBase.cpp:685-892, 1035-1043   -> ConvStackTrainer.fit / train_step
    Conv2D layer (kernels + bias)     Neuro/src/Layers/Conv2D.cpp:59-89                    -> ConvLayerSpec / parameters
    conv ops fwd + both gradients     Conv2DOp.cpp:22-42, Conv2dBiasActivationOp.cpp:23-66 -> op.Conv2DBiasActivation / ...InputGradient / ...KernelsGradient
    mean loss over the GLOBAL batch   ModelBase.cpp:344,374 (mean over GlobalAxis)
    Adam::MinimizationOperation       Neuro/src/Optimizers/Adam.cpp:66-111                 -> bias-corrected lr, per-bucket AdamStep
    SGD::MinimizationOperation        Neuro/src/Optimizers/SGD.cpp:45-52

What is new relative to the reference (which is single-device): every rank (one process per GPU) holds a full replica,
takes a contiguous slice of each global batch, and the kernel/bias gradients are SUM-all-reduced (NCCL over NVLink;
gloo in the CPU tests) between the backward pass and the optimiser step. Because each rank already divides its loss by
the GLOBAL element count, the summed gradient equals the single-device full-batch gradient and every replica applies
the identical update -- replicas stay in lock-step without a broadcast. All-reduces are issued per layer as soon as that
layer's kernel gradient is ready (backward visits layers last to first), so communication overlaps the remaining
input-gradient / kernel-gradient kernels.

`op` is any object with the TensorOpB200 method set (neuro__b200/tensor_op.py); the loss (MSE against a target, not on
the path) uses torch on the same device. Nothing here falls back to a CPU convolution.
"""
import math
from dataclasses import dataclass

import torch

from . import lib
from .lib import NCHW


@dataclass
class ConvLayerSpec:
    """Conv2D(inputShape, filtersNum, filterSize, stride, padding, activation) -- Neuro/include/Layers/Conv2D.h:16"""
    filters: int
    filter_size: int
    stride: int = 1
    padding: int = 0
    activation: int = lib.ACT_RELU
    alpha: float = 0.0


class ConvStackTrainer:
    """Sequential stack of Conv2D layers trained with MSE against a target, data-parallel over `group`."""

    def __init__(self, op, in_shape, layers, device, optimizer="adam", lr=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8,
                 seed=1337, group=None, world_size=1, rank=0):
        self.op, self.device, self.layers = op, device, layers
        self.optimizer, self.lr, self.beta1, self.beta2, self.eps = optimizer, lr, beta1, beta2, epsilon
        self.group, self.world, self.rank = group, world_size, rank
        self.iteration = 0
        C, H, W = in_shape
        # one flat bucket for all parameters / gradients / Adam moments; per-layer views into it
        shapes = []
        for l in layers:
            shapes.append((l.filters, C, l.filter_size, l.filter_size))
            H = (H + 2 * l.padding - l.filter_size) // l.stride + 1
            W = (W + 2 * l.padding - l.filter_size) // l.stride + 1
            C = l.filters
        self.out_shape = (C, H, W)
        total = sum(s[0] * s[1] * s[2] * s[3] + s[0] for s in shapes)
        self.params = torch.zeros(total, device=device)
        self.grads = torch.zeros(total, device=device)
        self.m = torch.zeros(total, device=device)
        self.v = torch.zeros(total, device=device)
        self.views = []
        gen = torch.Generator(device="cpu"); gen.manual_seed(seed)   # identical initial replicas on every rank
        off = 0
        for (K, Cin, R, S) in shapes:
            nw = K * Cin * R * S
            limit = math.sqrt(6.0 / (Cin * R * S + K * R * S))        # GlorotUniform, VarianceScaling.cpp:59-65
            w0 = ((torch.rand(K, Cin, R, S, generator=gen) * 2 - 1) * limit).to(device)
            view = dict(w=self.params[off:off + nw].view(K, Cin, R, S), dw=self.grads[off:off + nw].view(K, Cin, R, S),
                        b=self.params[off + nw:off + nw + K], db=self.grads[off + nw:off + nw + K], lo=off, hi=off + nw + K)
            view["w"].copy_(w0)                                         # bias init = zeros (Conv2D.h:46)
            self.views.append(view)
            off += nw + K

    # -- one step on this rank's shard; returns the GLOBAL mean loss
    def train_step(self, x_shard, target_shard, global_batch):
        import torch.distributed as dist
        op = self.op
        acts = [x_shard]
        for l, v in zip(self.layers, self.views):
            xin = acts[-1]
            N, _, H, W = xin.shape
            Ho = (H + 2 * l.padding - l.filter_size) // l.stride + 1
            Wo = (W + 2 * l.padding - l.filter_size) // l.stride + 1
            y = torch.empty((N, l.filters, Ho, Wo), device=self.device)
            op.Conv2DBiasActivation(xin, v["w"], l.stride, l.padding, l.padding, v["b"], l.activation, l.alpha, y)
            acts.append(y)
        out = acts[-1]
        # MSE, mean over the GLOBAL batch (every rank divides by the global element count)
        count = global_batch * out[0].numel()
        diff = out - target_shard
        loss_local = (diff * diff).sum() / count
        grad = diff * (2.0 / count)

        works = []
        for i in reversed(range(len(self.layers))):
            l, v = self.layers[i], self.views[i]
            # backward of Conv2dBiasActivationOp (Conv2dBiasActivationOp.cpp:47-60): activation gradient and bias gradient
            # in one pass over the output gradient, then kernel gradient and input gradient of the result
            dz = torch.empty_like(grad)
            op.Conv2DBiasActivationGradient(acts[i + 1], grad.contiguous(), l.activation, l.alpha, dz, v["db"])
            grad = dz
            op.Conv2DKernelsGradient(acts[i], grad, l.stride, l.padding, l.padding, NCHW, v["dw"])
            if self.world > 1:
                works.append(dist.all_reduce(self.grads[v["lo"]:v["hi"]], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            if i > 0:
                dx = torch.empty_like(acts[i])
                op.Conv2DInputGradient(grad, v["w"], l.stride, l.padding, l.padding, NCHW, dx)
                grad = dx
        for wk in works:
            wk.wait()

        self.iteration += 1
        if self.optimizer == "adam":
            # bias-corrected step size, Adam.cpp:90
            lr_t = self.lr * math.sqrt(1.0 - self.beta2 ** self.iteration) / (1.0 - self.beta1 ** self.iteration)
            op.AdamStep(self.params, self.grads, self.m, self.v, lr_t, self.beta1, self.beta2, self.eps)
        else:
            op.SgdStep(self.params, self.grads, self.lr)

        if self.world > 1:
            dist.all_reduce(loss_local, op=dist.ReduceOp.SUM, group=self.group)
        return float(loss_local)

    # -- ModelBase::Fit: epochs x batches, each global batch split into contiguous per-rank slices
    def fit(self, inputs, targets, batch_size, epochs=1, shuffle=False, seed=0):
        """inputs/targets: HOST tensors holding the full dataset on every rank (as the reference's Fit receives them)."""
        assert batch_size % self.world == 0, "batch size must divide evenly over the replicas"
        n = inputs.shape[0]
        per = batch_size // self.world
        losses = []
        gen = torch.Generator(); gen.manual_seed(seed)
        for _ in range(epochs):
            order = torch.randperm(n, generator=gen) if shuffle else torch.arange(n)
            for b0 in range(0, n - batch_size + 1, batch_size):
                idx = order[b0 + self.rank * per: b0 + (self.rank + 1) * per]
                xs = inputs[idx].to(self.device, non_blocking=True)
                ts = targets[idx].to(self.device, non_blocking=True)
                losses.append(self.train_step(xs.contiguous(), ts.contiguous(), batch_size))
        return losses
