"""Host logic of the data-parallel path on CPU with the gloo backend, world_size 2: sharding by image, grad_scale = 1/R and a
SUM all-reduce of the flat gradient buffer reproduce the single-process gradient and parameters."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tf_vqa_regat_b200.dp import DataParallelTrainer, allreduce_flat_, shard_batch, shard_range


class ToyEngine:
    """Same duck-typed surface as HotPathEngine, linear-quadratic loss on CPU: loss = mean_b ||x_b W||^2 / 2."""

    def __init__(self, W):
        self.params = W.clone().reshape(-1)
        self.grads = torch.zeros_like(self.params)
        self.shape = W.shape

    def fwd_bwd(self, features, boxes, q_att, q_last, target, grad_scale=1.0):
        W = self.params.view(self.shape)
        y = features @ W
        self.grads.copy_((features.t() @ y / features.shape[0] * grad_scale).reshape(-1))
        return {"loss": 0.5 * (y ** 2).sum(1).mean()}

    def update(self, lr, step):
        self.params -= lr * self.grads


def _worker(rank, world, port, W, X, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = ToyEngine(W)
    tr = DataParallelTrainer(eng, bucket_elems=7)
    if rank == 1:
        eng.params.add_(1.0)                       # diverged replica: broadcast must repair it
    tr.broadcast_params(0)
    batch = shard_batch({"features": X}, rank, world)
    for _ in range(3):
        tr.step(batch["features"], None, None, None, None, lr=0.1)
    out[rank] = eng.params.clone()
    dist.destroy_process_group()


def test_two_rank_gloo_equals_single_process():
    torch.manual_seed(0)
    W = torch.randn(6, 5, dtype=torch.float64)
    X = torch.randn(8, 6, dtype=torch.float64)
    ref = ToyEngine(W)
    for _ in range(3):
        ref.fwd_bwd(X, None, None, None, None)
        ref.update(0.1, 0)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29517, W, X, out), nprocs=2, join=True)
    for r in range(2):
        torch.testing.assert_close(out[r], ref.params, rtol=1e-12, atol=1e-12)


def test_shard_helpers():
    assert shard_range(256, 3, 8) == (96, 128)
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)
    b = {"a": np.arange(12).reshape(6, 2), "b": np.arange(6)}
    s = shard_batch(b, 1, 3)
    assert s["a"].tolist() == [[4, 5], [6, 7]] and s["b"].tolist() == [2, 3]
    t = torch.ones(5)
    assert allreduce_flat_(t) is t               # no process group: identity
