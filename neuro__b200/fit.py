"""Batch-sharded data-parallel training of a convolution stack -- the Fit() slice of the hot path.

Mirrors what the reference does per training step, restricted to what lies on the path:

    ModelBase::Fit / TrainStep        Neuro/src/Models/ModelBase.cpp:685-892, 1035-1043   -> ConvStackTrainer.fit / train_step
    Conv2D layer (kernels + bias)     Neuro/src/Layers/Conv2D.cpp:59-89                    -> ConvLayerSpec / parameters
    Conv2DTranspose layer             Neuro/src/Layers/Conv2DTranspose.cpp                 -> ConvLayerSpec(transposed=True)
    BatchNormalization layer          Neuro/src/Layers/BatchNormalization.cpp              -> ConvLayerSpec(batch_norm=True)
    Pooling2D / UpSampling2D layers   Neuro/src/Layers/Pooling2D.cpp, UpSampling2D.cpp     -> PoolSpec / UpSampleSpec
    conv ops fwd + both gradients     Conv2DOp.cpp:22-42, Conv2dBiasActivationOp.cpp:23-66, Conv2dTransposeOp.cpp:22-41
    mean loss over the GLOBAL batch   ModelBase.cpp:344,374 (mean over GlobalAxis)
    Adam::MinimizationOperation       Neuro/src/Optimizers/Adam.cpp:66-111                 -> bias-corrected lr, one AdamStep over the flat bucket
    SGD::MinimizationOperation        Neuro/src/Optimizers/SGD.cpp:45-52

What is new relative to the reference (which is single-device): every rank (one process per GPU) holds a full replica,
takes a contiguous slice of each global batch, and the parameter gradients are SUM-all-reduced (NCCL over NVLink; gloo
in the CPU tests) between the backward pass and the optimiser step. Because each rank already divides its loss by the
GLOBAL element count, the summed gradient equals the single-device full-batch gradient and every replica applies the
identical update -- replicas stay in lock-step without a broadcast.

Exchange: the flat gradient buffer is cut into BUCKETS of >= `bucket_bytes` (backward fills it from the end, so a bucket
is complete when the kernel gradient of its first layer is enqueued); each bucket is one all-reduce issued on the spot,
underneath the remaining backward kernels. Small models (<= one bucket) exchange once: latency, not bandwidth, is what
NVLink 5 / NVSwitch charges for messages of this size. Batch-norm layers normalise with the statistics of the GLOBAL
batch (`sync_bn`): 2*C floats per layer are all-gathered in the forward pass and 3*C floats all-reduced in the backward
pass (include/neuro_b200.h: nb200_batch_norm_moments ...), so the loss curve of N replicas is that of one device.

Activations, gradients and scratch live in STATIC device buffers planned once per shard shape: a step allocates nothing
and does not synchronise with the host (the loss is read back only on request), which also lets the whole forward +
backward + exchange be captured as ONE CUDA graph (`use_graph=True`) -- per-call host cost (ctypes, tensor-map encode,
launch) disappears from small-layer models, and the tensor maps encoded at capture time are the plan that is replayed.

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
    """Conv2D(inputShape, filtersNum, filterSize, stride, padding, activation) -- Neuro/include/Layers/Conv2D.h:16;
    transposed=True: Conv2DTranspose(outputDepth, filterSize, stride, padding, activation) -- Layers/Conv2DTranspose.h;
    batch_norm=True: a BatchNormalization layer (Spatial mode) between the convolution (+bias) and its activation, as the GAN /
    pix2pix stacks build it (DeepConvGAN.cpp:3-43, Pix2Pix.cpp:4-109)."""
    filters: int
    filter_size: int
    stride: int = 1
    padding: int = 0
    activation: int = lib.ACT_RELU
    alpha: float = 0.0
    batch_norm: bool = False
    transposed: bool = False


@dataclass
class PoolSpec:
    """Pooling2D(filterSize, stride, padding, mode) -- Neuro/include/Layers/Pooling2D.h"""
    filter_size: int = 2
    stride: int = 2
    padding: int = 0
    mode: int = lib.POOL_MAX


@dataclass
class UpSampleSpec:
    """UpSampling2D(scaleFactor) -- Neuro/include/Layers/UpSampling2D.h"""
    scale: int = 2


BN_MOMENTUM, BN_EPSILON = 0.99, 0.001   # BatchNormalization.h defaults


class ConvStackTrainer:
    """Sequential stack of Conv2D / Conv2DTranspose (+BatchNormalization) / Pooling2D / UpSampling2D layers trained with MSE
    against a target, data-parallel over `group`."""

    def __init__(self, op, in_shape, layers, device, optimizer="adam", lr=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8,
                 seed=1337, group=None, world_size=1, rank=0, bucket_bytes=24 << 20, sync_bn=True, input_gradient=False,
                 use_graph=False):
        self.op, self.device, self.layers = op, device, layers
        self.optimizer, self.lr, self.beta1, self.beta2, self.eps = optimizer, lr, beta1, beta2, epsilon
        self.group, self.world, self.rank = group, world_size, rank
        self.sync_bn, self.input_gradient, self.use_graph = sync_bn, input_gradient, use_graph
        self.iteration = 0
        self._plan_n = None
        self._graph = None
        self.graph_kernel_launches = 0
        C, H, W = in_shape
        self.in_shape = in_shape
        # one flat buffer for all parameters / gradients / Adam moments; per-layer views into it
        self.shapes = []   # per layer: (C_in, H_in, W_in, C_out, H_out, W_out)
        sizes = []
        for l in layers:
            if isinstance(l, ConvLayerSpec):
                if l.transposed:
                    Ho = (H - 1) * l.stride + l.filter_size - 2 * l.padding      # Tensor::GetConvTransposeOutputShape
                    Wo = (W - 1) * l.stride + l.filter_size - 2 * l.padding
                    wshape = (C, l.filters, l.filter_size, l.filter_size)         # (inDepth, outDepth, F, F), Conv2DTranspose.cpp
                else:
                    Ho = (H + 2 * l.padding - l.filter_size) // l.stride + 1
                    Wo = (W + 2 * l.padding - l.filter_size) // l.stride + 1
                    wshape = (l.filters, C, l.filter_size, l.filter_size)
                self.shapes.append((C, H, W, l.filters, Ho, Wo))
                sizes.append((wshape, l.filters, l.filters if l.batch_norm else 0))
                C, H, W = l.filters, Ho, Wo
            elif isinstance(l, PoolSpec):
                Ho = (H + 2 * l.padding - l.filter_size) // l.stride + 1
                Wo = (W + 2 * l.padding - l.filter_size) // l.stride + 1
                self.shapes.append((C, H, W, C, Ho, Wo)); sizes.append(None)
                H, W = Ho, Wo
            else:
                self.shapes.append((C, H, W, C, H * l.scale, W * l.scale)); sizes.append(None)
                H, W = H * l.scale, W * l.scale
        self.out_shape = (C, H, W)
        total = sum((s[0][0] * s[0][1] * s[0][2] * s[0][3] + s[1] + 2 * s[2]) for s in sizes if s)
        self.params = torch.zeros(total, device=device)
        self.grads = torch.zeros(total, device=device)
        self.m = torch.zeros(total, device=device)
        self.v = torch.zeros(total, device=device)
        self.views = []
        gen = torch.Generator(device="cpu"); gen.manual_seed(seed)   # identical initial replicas on every rank
        off = 0
        for l, s in zip(layers, sizes):
            if s is None:
                self.views.append(None)
                continue
            wshape, K, nbn = s
            nw = wshape[0] * wshape[1] * wshape[2] * wshape[3]
            fan = wshape[2] * wshape[3]
            limit = math.sqrt(6.0 / (wshape[1] * fan + wshape[0] * fan))   # GlorotUniform, VarianceScaling.cpp:59-65
            w0 = ((torch.rand(*wshape, generator=gen) * 2 - 1) * limit).to(device)
            view = dict(w=self.params[off:off + nw].view(*wshape), dw=self.grads[off:off + nw].view(*wshape),
                        b=self.params[off + nw:off + nw + K], db=self.grads[off + nw:off + nw + K], lo=off)
            view["w"].copy_(w0)                                         # bias init = zeros (Conv2D.h:46)
            off += nw + K
            if nbn:
                view["gamma"] = self.params[off:off + K]; view["dgamma"] = self.grads[off:off + K]
                view["beta"] = self.params[off + K:off + 2 * K]; view["dbeta"] = self.grads[off + K:off + 2 * K]
                view["gamma"].fill_(1.0)                                # BatchNormalization.cpp: gamma ones, beta zeros
                view["rmean"] = torch.zeros(K, device=device); view["rvar"] = torch.ones(K, device=device)
                off += 2 * K
            view["hi"] = off
            self.views.append(view)
        # all-reduce buckets over the flat gradient buffer, filled from the end: (lo, hi, index of the layer that completes it)
        self.buckets = []
        hi, acc_lo, first = total, total, None
        for i in reversed(range(len(layers))):
            v = self.views[i]
            if v is None:
                continue
            acc_lo, first = v["lo"], i
            if (hi - acc_lo) * 4 >= bucket_bytes:
                self.buckets.append((acc_lo, hi, i)); hi = acc_lo; first = None
        if first is not None and hi > acc_lo:
            self.buckets.append((acc_lo, hi, first))
        # The bucket that ends the backward pass cannot start before the FIRST layer's kernel gradient and the optimiser waits for it:
        # its all-reduce is the exposed tail of the exchange. Keep that one small (<= tail_bytes): split the leading layers off so
        # that everything else is already in flight underneath the first layers' (largest-activation, slowest) backward kernels.
        tail_bytes = 1 << 20
        lo, hi, _ = self.buckets[-1] if self.buckets else (0, 0, 0)
        if (hi - lo) * 4 > tail_bytes:
            cut = None
            for i, v in enumerate(self.views):
                if v is None:
                    continue
                if v["lo"] >= hi or (v["hi"] - lo) * 4 > tail_bytes:
                    break
                cut = (v["hi"], i)
            if cut is not None and cut[0] < hi:
                nxt = min(i for i, v in enumerate(self.views) if v is not None and v["lo"] >= cut[0])
                self.buckets[-1] = (cut[0], hi, nxt)
                self.buckets.append((lo, cut[0], self.first_param_layer_index()))
        self.first_param_layer = self.first_param_layer_index()

    def first_param_layer_index(self):
        return min((i for i, v in enumerate(self.views) if v is not None), default=0)

    # ---- static buffers for one shard shape
    def _plan(self, n):
        dev = self.device
        e = lambda *s: torch.empty(s, device=dev)
        self.x_in = e(n, *self.in_shape)
        self.target = e(n, *self.out_shape)
        self.acts, self.pre, self.dact, self.dpre, self.bn = [self.x_in], [], [], [], []
        for l, (C, H, W, K, Ho, Wo) in zip(self.layers, self.shapes):
            self.acts.append(e(n, K, Ho, Wo))
            self.dact.append(e(n, K, Ho, Wo))                  # gradient w.r.t. this layer's output
            if isinstance(l, ConvLayerSpec):
                self.dpre.append(e(n, K, Ho, Wo))              # ... w.r.t. the convolution's output (after activation / BN gradients)
                if l.batch_norm:
                    self.pre.append(e(n, K, Ho, Wo))           # convolution (+bias) output = batch-norm input
                    self.bn.append(dict(save_mean=e(K), save_inv=e(K), moments=e(K, 2), all_moments=e(self.world, K, 2),
                                        sums=e(K, 3), gsums=e(K, 3), dbn=e(n, K, Ho, Wo)))
                else:
                    self.pre.append(None); self.bn.append(None)
            else:
                self.dpre.append(None); self.pre.append(None); self.bn.append(None)
        self.dx_in = e(n, *self.in_shape) if self.input_gradient else None
        self.loss_buf = torch.zeros((), device=dev)
        self._plan_n = n

    def _tag(self, i):
        if hasattr(self.op, "set_tag"):
            self.op.set_tag(i)

    # ---- forward over the static buffers
    def _forward(self):
        import torch.distributed as dist
        op = self.op
        for i, (l, v) in enumerate(zip(self.layers, self.views)):
            xin, y = self.acts[i], self.acts[i + 1]
            self._tag(i)
            if isinstance(l, PoolSpec):
                op.Pool2D(xin, l.filter_size, l.stride, l.mode, l.padding, l.padding, NCHW, y)
            elif isinstance(l, UpSampleSpec):
                op.UpSample2D(xin, l.scale, y)
            elif not l.batch_norm and not l.transposed:
                op.Conv2DBiasActivation(xin, v["w"], l.stride, l.padding, l.padding, v["b"], l.activation, l.alpha, y)
            else:
                z = self.pre[i] if l.batch_norm else y
                if l.transposed:
                    op.Conv2DTransposed(xin, v["w"], l.stride, l.padding, NCHW, z)
                else:
                    op.Conv2D(xin, v["w"], l.stride, l.padding, l.padding, NCHW, z)
                if not l.batch_norm:
                    op.BiasActivation(z, v["b"], l.activation, l.alpha, y)
                    continue
                op.BiasActivation(z, v["b"], lib.ACT_IDENTITY, 0.0, z)       # conv -> add bias -> BN -> activation (Pix2Pix.cpp)
                b = self.bn[i]
                if self.world > 1 and self.sync_bn:
                    op.BatchNormalizationMoments(z, lib.BN_SPATIAL, b["moments"])
                    dist.all_gather_into_tensor(b["all_moments"].view(-1, 2), b["moments"], group=self.group)   # rank-major concatenation
                    op.BatchNormalizationTrainFromMoments(b["all_moments"], self.world, z, lib.BN_SPATIAL, v["gamma"], v["beta"], BN_MOMENTUM,
                                                          BN_EPSILON, v["rmean"], v["rvar"], b["save_mean"], b["save_inv"], y)
                else:
                    op.BatchNormalizationTrain(z, lib.BN_SPATIAL, v["gamma"], v["beta"], BN_MOMENTUM, BN_EPSILON, v["rmean"], v["rvar"],
                                               b["save_mean"], b["save_inv"], y)
                op.BiasActivation(y, None, l.activation, l.alpha, y)

    # ---- backward + exchange; grad w.r.t. the last output is in self.dact[-1]
    def _backward(self, weight_gradients=True):
        import torch.distributed as dist
        op = self.op
        works = []
        pending = {b[2]: b for b in self.buckets}
        fused_act = set()   # conv layers whose activation / bias gradient was produced by the pooling layer above them
        for i in reversed(range(len(self.layers))):
            l, v = self.layers[i], self.views[i]
            xin, y, dy = self.acts[i], self.acts[i + 1], self.dact[i]
            need_dx = i > 0 or self.input_gradient
            dx = (self.dact[i - 1] if i > 0 else self.dx_in) if need_dx else None
            self._tag(i)
            if isinstance(l, PoolSpec):
                prev = self.layers[i - 1] if i > 0 else None
                if (isinstance(prev, ConvLayerSpec) and not prev.batch_norm and hasattr(op, "Pool2DGradientActivation")
                        and op.Pool2DGradientActivationSupported(xin, l.filter_size, l.stride, l.mode, l.padding, l.padding, NCHW, y)):
                    # "fused conv layer -> max pooling": pooling gradient, the conv's activation gradient and its bias gradient in one pass
                    op.Pool2DGradientActivation(y, xin, dy, l.filter_size, l.stride, l.mode, l.padding, l.padding, NCHW, prev.activation, prev.alpha,
                                                self.dpre[i - 1], self.views[i - 1]["db"] if weight_gradients else None)
                    fused_act.add(i - 1)
                elif need_dx:
                    op.Pool2DGradient(y, xin, dy, l.filter_size, l.stride, l.mode, l.padding, l.padding, NCHW, dx)
                continue
            if isinstance(l, UpSampleSpec):
                if need_dx:
                    op.UpSample2DGradient(dy, l.scale, dx)
                continue
            dz = self.dpre[i]
            if l.batch_norm:
                b = self.bn[i]
                z = self.pre[i]
                op.Conv2DBiasActivationGradient(y, dy, l.activation, l.alpha, b["dbn"], None)          # through the activation
                if self.world > 1 and self.sync_bn:
                    op.BatchNormalizationGradientSums(z, lib.BN_SPATIAL, b["dbn"], b["save_mean"], b["sums"])
                    b["gsums"].copy_(b["sums"])
                    dist.all_reduce(b["gsums"], op=dist.ReduceOp.SUM, group=self.group)
                    op.BatchNormalizationGradientFromSums(self.world, b["gsums"], b["sums"], z, lib.BN_SPATIAL, v["gamma"], b["dbn"],
                                                          b["save_mean"], b["save_inv"], v["dgamma"] if weight_gradients else None,
                                                          v["dbeta"] if weight_gradients else None, dz)
                else:
                    op.BatchNormalizationGradient(z, lib.BN_SPATIAL, v["gamma"], BN_EPSILON, b["dbn"], b["save_mean"], b["save_inv"],
                                                  v["dgamma"], v["dbeta"], True, dz)
                if weight_gradients:
                    op.Conv2DBiasGradient(dz, v["db"])
            elif i not in fused_act:
                # backward of Conv2dBiasActivationOp (Conv2dBiasActivationOp.cpp:47-60): activation gradient and bias gradient
                # in one pass over the output gradient, then kernel gradient and input gradient of the result
                op.Conv2DBiasActivationGradient(y, dy, l.activation, l.alpha, dz, v["db"] if weight_gradients else None)
            if weight_gradients:
                if l.transposed:
                    op.Conv2DTransposedKernelsGradient(xin, dz, l.stride, l.padding, NCHW, v["dw"])
                else:
                    op.Conv2DKernelsGradient(xin, dz, l.stride, l.padding, l.padding, NCHW, v["dw"])
                if self.world > 1 and i in pending:
                    lo, hi, _ = pending[i]
                    works.append(dist.all_reduce(self.grads[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            if need_dx:
                if l.transposed:
                    op.Conv2DTransposedInputsGradient(dz, v["w"], l.stride, l.padding, NCHW, dx)
                else:
                    op.Conv2DInputGradient(dz, v["w"], l.stride, l.padding, l.padding, NCHW, dx)
        for wk in works:
            wk.wait()

    def _loss_and_seed(self, global_batch):
        # MSE, mean over the GLOBAL batch (every rank divides by the global element count)
        out = self.acts[-1]
        count = global_batch * out[0].numel()
        torch.sub(out, self.target, out=self.dact[-1])
        self.loss_buf.copy_((self.dact[-1] * self.dact[-1]).sum() / count)
        self.dact[-1].mul_(2.0 / count)

    def _fwd_bwd(self, global_batch):
        self._forward()
        self._loss_and_seed(global_batch)
        self._backward()

    def _optimizer_step(self):
        self.iteration += 1
        self._tag(-1)
        if self.optimizer == "adam":
            # bias-corrected step size, Adam.cpp:90
            lr_t = self.lr * math.sqrt(1.0 - self.beta2 ** self.iteration) / (1.0 - self.beta1 ** self.iteration)
            self.op.AdamStep(self.params, self.grads, self.m, self.v, lr_t, self.beta1, self.beta2, self.eps)
        else:
            self.op.SgdStep(self.params, self.grads, self.lr)

    def load_batch(self, x_shard, target_shard):
        """Copies one shard into the static input buffers (host or device source; pinned host memory copies asynchronously)."""
        if self._plan_n != x_shard.shape[0]:
            self._plan(x_shard.shape[0]); self._graph = None
        self.x_in.copy_(x_shard, non_blocking=True)
        self.target.copy_(target_shard, non_blocking=True)

    def run_step(self, global_batch):
        """One training step on the batch already in the static buffers. Asynchronous: nothing is read back."""
        if self.use_graph and self.x_in.is_cuda:
            if self._graph is None:
                self._capture(global_batch)
            self._graph.replay()
        else:
            self._fwd_bwd(global_batch)
        self._optimizer_step()

    def _capture(self, global_batch):
        # warm-up on a side stream (workspaces grow, NCCL channels open), then capture forward + backward + exchange
        params, m, v, it = self.params.clone(), self.m.clone(), self.v.clone(), self.iteration
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._fwd_bwd(global_batch)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        counter = lib.load().nb200_kernel_launches
        n0 = counter()
        with torch.cuda.graph(g):
            self._fwd_bwd(global_batch)
        self._graph = g
        self.graph_kernel_launches = int(counter() - n0)   # library kernels one replay launches (bookkeeping for benchmarks)
        # batch-norm running statistics advanced during warm-up / capture are part of the model state; parameters are not touched
        self.params.copy_(params); self.m.copy_(m); self.v.copy_(v); self.iteration = it

    # -- one step on this rank's shard; returns the GLOBAL mean loss
    def train_step(self, x_shard, target_shard, global_batch):
        import torch.distributed as dist
        self.load_batch(x_shard, target_shard)
        self.run_step(global_batch)
        loss = self.loss_buf.clone()
        if self.world > 1:
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        return float(loss)

    # -- style-transfer form of the path: frozen weights, forward + gradient back to the IMAGE (no kernel gradients, no update)
    def image_gradient_step(self, global_batch):
        assert self.input_gradient
        self._forward()
        self._loss_and_seed(global_batch)
        self._backward(weight_gradients=False)
        return self.dx_in

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
                losses.append(self.train_step(inputs[idx].contiguous(), targets[idx].contiguous(), batch_size))
        return losses
