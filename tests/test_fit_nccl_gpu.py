"""fit.py ITSELF under NCCL on real GPUs (the NCCL path had no hardware test): world_size 2, one process per GPU,
TensorOpB200 underneath. Batch sharding + bucketed gradient all-reduce (kernel, bias, gamma, beta gradients) + batch-norm
statistics over the global batch must reproduce the single-GPU full-batch run, replicas must stay bit-identical, and the
CUDA-graph step (collectives captured inside) must equal the eager step. Skipped on a one-GPU box (run with gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuro__b200 import lib, synth
from neuro__b200.fit import ConvLayerSpec, ConvStackTrainer, PoolSpec, UpSampleSpec
from neuro__b200.tensor_op import TensorOpB200

pytestmark = pytest.mark.gpu

LAYERS = [ConvLayerSpec(32, 3, 2, 1, lib.ACT_LEAKY_RELU, 0.2, batch_norm=True), ConvLayerSpec(64, 3, 1, 1, lib.ACT_RELU), PoolSpec(2, 2, 0, lib.POOL_MAX),
          UpSampleSpec(2), ConvLayerSpec(32, 4, 2, 1, lib.ACT_RELU, batch_norm=True, transposed=True), ConvLayerSpec(3, 3, 1, 1, lib.ACT_TANH)]
IN_SHAPE = (3, 32, 32)
N, BATCH, EPOCHS = 32, 16, 3


def _data(out_shape):
    x = torch.from_numpy(synth.uniform(synth.SEED_X, (N,) + IN_SHAPE))
    t = torch.from_numpy(synth.uniform(synth.SEED_DY, (N,) + out_shape, -0.5, 0.5))
    return x, t


def _train(world, rank, use_graph, group=None):
    tr = ConvStackTrainer(TensorOpB200(lib.MATH_FP32), IN_SHAPE, LAYERS, torch.device("cuda", rank), optimizer="adam", lr=0.002,
                          group=group, world_size=world, rank=rank, bucket_bytes=64 << 10, use_graph=use_graph)
    assert len(tr.buckets) >= 2
    x, t = _data(tr.out_shape)
    losses = tr.fit(x, t, BATCH, epochs=EPOCHS)
    torch.cuda.synchronize()
    return tr.params.detach().cpu(), losses


def _worker(rank, world, port, use_graph, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        params, losses = _train(world, rank, use_graph)
        gathered = [torch.zeros_like(params).cuda() for _ in range(world)]
        dist.all_gather(gathered, params.cuda())
        if rank == 0:
            torch.save({"params": [g.cpu() for g in gathered], "losses": losses}, out)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("use_graph", [False, True], ids=["eager", "cuda_graph"])
def test_two_gpus_match_one_gpu(tmp_path, use_graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    torch.cuda.set_device(0)
    ref_params, ref_losses = _train(1, 0, False)
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), use_graph, out), nprocs=2, join=True)
    got = torch.load(out)
    p0, p1 = got["params"]
    assert torch.equal(p0, p1), "replicas diverged"
    # fp32 kernels: shard-summed gradients and merged batch-norm moments equal the full-batch values up to addition order
    assert np.allclose(got["losses"], ref_losses, rtol=2e-4, atol=1e-7), (got["losses"], ref_losses)
    assert float((p0 - ref_params).abs().max()) <= 2e-3     # Adam's 1/sqrt(v) amplifies last-bit gradient differences near zero
    assert ref_losses[-1] < ref_losses[0]
