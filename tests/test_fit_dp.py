"""Data-parallel Fit() host logic, world_size 2 over gloo on CPU: batch sharding + per-layer gradient all-reduce + one
optimiser step must reproduce single-process full-batch training, and replicas must stay bit-identical."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neuro__b200 import lib, synth  # noqa: E402
from neuro__b200.fit import ConvLayerSpec, ConvStackTrainer, PoolSpec, UpSampleSpec  # noqa: E402
from tests.oracle_op import OracleOp  # noqa: E402

LAYERS = [ConvLayerSpec(6, 3, 1, 1, lib.ACT_RELU), ConvLayerSpec(4, 3, 2, 1, lib.ACT_LEAKY_RELU, 0.2), ConvLayerSpec(2, 3, 1, 1, lib.ACT_TANH)]
IN_SHAPE = (3, 10, 10)
N, BATCH, EPOCHS = 8, 4, 2


def _data():
    x = torch.from_numpy(synth.uniform(synth.SEED_X, (N,) + IN_SHAPE))
    t = torch.from_numpy(synth.uniform(synth.SEED_DY, (N, 2, 5, 5), 0.0, 1.0))
    return x, t


# every layer kind + batch norm whose statistics must span both shards (DeepConvGAN-like; ModelBase::Fit over replicas)
BN_LAYERS = [ConvLayerSpec(6, 3, 2, 1, lib.ACT_LEAKY_RELU, 0.2, batch_norm=True), PoolSpec(1, 1, 0, lib.POOL_MAX), UpSampleSpec(2),
             ConvLayerSpec(4, 4, 2, 1, lib.ACT_RELU, batch_norm=True, transposed=True), ConvLayerSpec(2, 3, 4, 1, lib.ACT_TANH)]


def _train(world, rank, optimizer, group=None):
    bn = optimizer.endswith("+bn")
    tr = ConvStackTrainer(OracleOp(), IN_SHAPE, BN_LAYERS if bn else LAYERS, torch.device("cpu"), optimizer=optimizer.split("+")[0],
                          lr=0.01, group=group, world_size=world, rank=rank, bucket_bytes=256 if bn else 24 << 20)
    assert tr.out_shape == (2, 5, 5)
    if bn:
        assert len(tr.buckets) >= 2 and tr.buckets[0][1] == tr.grads.numel() and tr.buckets[-1][0] == 0   # several buckets, end first
    x, t = _data()
    losses = tr.fit(x, t, BATCH, epochs=EPOCHS)
    return tr.params.clone(), losses


def _worker(rank, world, port, optimizer, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params, losses = _train(world, rank, optimizer)
        gathered = [torch.zeros_like(params) for _ in range(world)]
        dist.all_gather(gathered, params)
        if rank == 0:
            torch.save({"params": [g.clone() for g in gathered], "losses": losses}, out)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("optimizer", ["adam", "sgd", "sgd+bn"])
def test_two_replicas_match_single_process(tmp_path, optimizer):
    ref_params, ref_losses = _train(1, 0, optimizer)
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), optimizer, out), nprocs=2, join=True)
    got = torch.load(out)
    p0, p1 = got["params"]
    assert torch.equal(p0, p1), "replicas diverged"                      # identical update on every replica
    # summed shard gradients == full-batch gradient (up to fp32 addition order)
    # (batch norm: the single process sums a group sequentially in fp32 like the reference, the replicas merge per-shard moments)
    assert float((p0 - ref_params).abs().max()) <= (2e-5 if optimizer.endswith("+bn") else 2e-6)
    assert np.allclose(got["losses"], ref_losses, rtol=2e-5 if optimizer.endswith("+bn") else 1e-5, atol=1e-7)
    assert ref_losses[-1] < ref_losses[0]                               # and it actually trains


def test_uneven_batch_is_rejected():
    tr = ConvStackTrainer(OracleOp(), IN_SHAPE, LAYERS, torch.device("cpu"), world_size=3, rank=0)
    x, t = _data()
    with pytest.raises(AssertionError):
        tr.fit(x, t, BATCH)
