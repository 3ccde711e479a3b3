"""Loss curves of the CUDA path track the reference over N training steps (BASELINE.json north_star), single GPU:
the same ConvStackTrainer is run once with the oracle-backed op on CPU and once with TensorOpB200 on cuda:0."""
import numpy as np
import pytest
import torch

from neuro__b200 import lib, synth
from neuro__b200.fit import ConvLayerSpec, ConvStackTrainer
from neuro__b200.tensor_op import TensorOpB200
from tests.oracle_op import OracleOp

pytestmark = pytest.mark.gpu

LAYERS = [ConvLayerSpec(16, 3, 1, 1, lib.ACT_RELU), ConvLayerSpec(16, 3, 1, 1, lib.ACT_LEAKY_RELU, 0.2), ConvLayerSpec(8, 3, 1, 1, lib.ACT_TANH)]
IN_SHAPE = (8, 32, 32)
N, BATCH, EPOCHS = 8, 4, 5


def _run(op, device, optimizer):
    tr = ConvStackTrainer(op, IN_SHAPE, LAYERS, device, optimizer=optimizer, lr=0.01)
    x = torch.from_numpy(synth.uniform(synth.SEED_X, (N,) + IN_SHAPE))
    t = torch.from_numpy(synth.uniform(synth.SEED_DY, (N,) + tr.out_shape, -0.5, 0.5))
    return tr.fit(x, t, BATCH, epochs=EPOCHS), tr.params.detach().cpu().numpy()


@pytest.mark.parametrize("optimizer", ["adam", "sgd"])
@pytest.mark.parametrize("math,tol", [(lib.MATH_TF32, 5e-3), (lib.MATH_FP32, 1e-4)], ids=["tf32", "fp32"])
def test_loss_curve_tracks_reference(math, tol, optimizer):
    ref_losses, ref_params = _run(OracleOp(), torch.device("cpu"), optimizer)
    op = TensorOpB200(math)
    d = lib.ConvDesc(BATCH, 8, 32, 32, 16, 3, 3, 32, 32, 1, 1, 1, lib.NCHW, math)
    if math == lib.MATH_TF32:
        assert op.kernel_name(lib.OP_FORWARD, d) == "tcgen05_fprop" and op.kernel_name(lib.OP_KERNELS_GRADIENT, d) == "tcgen05_rowfold_wgrad"
    losses, params = _run(op, torch.device("cuda", 0), optimizer)
    assert len(losses) == len(ref_losses) == (N // BATCH) * EPOCHS
    rel = np.abs(np.array(losses) - np.array(ref_losses)) / np.abs(np.array(ref_losses))
    assert rel.max() <= tol, (losses, ref_losses)
    assert ref_losses[-1] < ref_losses[0]
    # Adam divides by sqrt(v): where the gradient is ~0 a TF32-sized perturbation can flip the sign of an lr-sized step,
    # so parameters are compared loosely there; SGD parameters track tightly.
    ptol = {("adam", lib.MATH_TF32): 5e-2, ("adam", lib.MATH_FP32): 1e-3, ("sgd", lib.MATH_TF32): 5e-3, ("sgd", lib.MATH_FP32): 1e-5}
    assert np.abs(params - ref_params).max() <= ptol[(optimizer, math)]


# DCGAN-discriminator-like stack (BASELINE configs[1]): RGB input, stride-2 3x3 convolutions with LeakyReLU(0.2) -- the first layer's
# kernel gradient runs on the strided few-channel kernel, the others on the gathered tensor-core kernels, every backward step goes
# through the fused activation-gradient + bias-gradient pass.
GAN_LAYERS = [ConvLayerSpec(16, 3, 2, 1, lib.ACT_LEAKY_RELU, 0.2), ConvLayerSpec(32, 3, 2, 1, lib.ACT_LEAKY_RELU, 0.2),
              ConvLayerSpec(8, 3, 1, 1, lib.ACT_TANH)]
GAN_IN = (3, 32, 32)


def _run_gan(op, device):
    tr = ConvStackTrainer(op, GAN_IN, GAN_LAYERS, device, optimizer="sgd", lr=0.05)
    x = torch.from_numpy(synth.uniform(synth.SEED_X, (N,) + GAN_IN))
    t = torch.from_numpy(synth.uniform(synth.SEED_DY, (N,) + tr.out_shape, -0.5, 0.5))
    return tr.fit(x, t, BATCH, epochs=EPOCHS), tr.params.detach().cpu().numpy()


@pytest.mark.parametrize("math,tol", [(lib.MATH_TF32, 5e-3), (lib.MATH_FP32, 1e-4)], ids=["tf32", "fp32"])
def test_strided_stack_loss_curve_tracks_reference(math, tol):
    ref_losses, ref_params = _run_gan(OracleOp(), torch.device("cpu"))
    op = TensorOpB200(math)
    d0 = lib.ConvDesc(BATCH, 3, 32, 32, 16, 3, 3, 16, 16, 2, 1, 1, lib.NCHW, math)
    assert op.kernel_name(lib.OP_KERNELS_GRADIENT, d0) == "strided_smallc_wgrad"
    if math == lib.MATH_TF32:
        d1 = lib.ConvDesc(BATCH, 16, 16, 16, 32, 3, 3, 8, 8, 2, 1, 1, lib.NCHW, math)
        assert [op.kernel_name(o, d1) for o in (0, 1, 2)] == ["tcgen05_gather_fprop", "tcgen05_gather_dgrad", "tcgen05_gather_wgrad"]
    losses, params = _run_gan(op, torch.device("cuda", 0))
    rel = np.abs(np.array(losses) - np.array(ref_losses)) / np.abs(np.array(ref_losses))
    assert rel.max() <= tol, (losses, ref_losses)
    assert ref_losses[-1] < ref_losses[0]
    assert np.abs(params - ref_params).max() <= (5e-3 if math == lib.MATH_TF32 else 1e-5)


# DeepConvGAN-like stack (Neuro.Examples/src/DeepConvGAN.cpp:3-43; BASELINE configs[1]): strided convolutions with a
# BatchNormalization between conv(+bias) and LeakyReLU, up-sampling + convolution and a transposed convolution on the generator
# side, max pooling -- every layer kind ConvStackTrainer knows, through the CUDA kernels.
from neuro__b200.fit import PoolSpec, UpSampleSpec  # noqa: E402

BN_LAYERS = [ConvLayerSpec(16, 3, 2, 1, lib.ACT_LEAKY_RELU, 0.2, batch_norm=True),            # 1x28x28 -> 16x14x14
             ConvLayerSpec(32, 3, 1, 1, lib.ACT_RELU, batch_norm=True),                       # 32x14x14
             PoolSpec(2, 2, 0, lib.POOL_MAX),                                                  # 32x7x7
             UpSampleSpec(2),                                                                  # 32x14x14
             ConvLayerSpec(16, 4, 2, 1, lib.ACT_RELU, batch_norm=True, transposed=True),       # 16x28x28
             ConvLayerSpec(1, 3, 1, 1, lib.ACT_TANH)]                                          # 1x28x28
BN_IN = (1, 28, 28)


def _run_bn(op, device, use_graph=False):
    tr = ConvStackTrainer(op, BN_IN, BN_LAYERS, device, optimizer="adam", lr=0.002, use_graph=use_graph)
    assert tr.out_shape == (1, 28, 28)
    x = torch.from_numpy(synth.uniform(synth.SEED_X, (16,) + BN_IN))
    t = torch.from_numpy(synth.uniform(synth.SEED_DY, (16,) + tr.out_shape, -0.5, 0.5))
    losses = tr.fit(x, t, 8, epochs=4)
    return losses, tr.params.detach().cpu().numpy(), [v["rvar"].cpu().numpy() for v in tr.views if v and "rvar" in v]


@pytest.mark.parametrize("math,tol", [(lib.MATH_TF32, 5e-3), (lib.MATH_FP32, 2e-4)], ids=["tf32", "fp32"])
def test_batch_norm_gan_stack_loss_curve_tracks_reference(math, tol):
    ref_losses, ref_params, ref_rvar = _run_bn(OracleOp(), torch.device("cpu"))
    losses, params, rvar = _run_bn(TensorOpB200(math), torch.device("cuda", 0))
    rel = np.abs(np.array(losses) - np.array(ref_losses)) / np.abs(np.array(ref_losses))
    assert rel.max() <= tol, (losses, ref_losses)
    assert ref_losses[-1] < ref_losses[0]
    # running statistics follow the reference's update rule; with momentum 0.99 they are essentially the LAST step's batch variance,
    # i.e. a function of parameters that 8 Adam steps (1/sqrt(v) amplifies TF32-sized gradient differences, see above) moved apart
    rtol = 3e-2 if math == lib.MATH_TF32 else 1e-3
    for a, b in zip(rvar, ref_rvar):
        assert np.abs(a - b).max() <= rtol * max(1.0, np.abs(b).max())


def test_cuda_graph_step_is_the_eager_step():
    """use_graph=True replays forward + backward as one CUDA graph over the static buffers: same kernels, same order, so the
    loss curve and the parameters are bit-identical to issuing the calls one by one."""
    eager = _run_bn(TensorOpB200(lib.MATH_TF32), torch.device("cuda", 0), use_graph=False)
    graph = _run_bn(TensorOpB200(lib.MATH_TF32), torch.device("cuda", 0), use_graph=True)
    assert eager[0] == graph[0]
    assert np.array_equal(eager[1], graph[1])
    # running statistics: the graph path advanced them three extra times during warm-up / capture (documented in fit.py)
    assert all(np.isfinite(a).all() for a in graph[2])
