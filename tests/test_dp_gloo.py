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


class OracleEngine:
    """HotPathEngine's duck-typed surface with the CPU oracle behind it (tiny widths): lets the 2-rank gloo run exercise the
    data-parallel contract of SURVEY 8e on the REAL path -- graphs are independent and the loss is a mean over graphs, so
    sum-all-reducing grad_scale = 1/R gradients of equal shards reproduces the single-process step, clip + Adamax included."""

    def __init__(self, cfg, flat):
        from tf_vqa_regat_b200.config import param_layout
        self.cfg, (self.entries, total) = cfg, param_layout(cfg)
        self.params = torch.tensor(np.asarray(flat, dtype=np.float64))
        self.grads = torch.zeros_like(self.params)
        self.m, self.u = np.zeros(total), np.zeros(total)

    def _named(self, buf):
        return {e.name: buf[e.offset:e.offset + e.numel].reshape(e.shape) for e in self.entries}

    def fwd_bwd(self, features, boxes, q_att, q_last, target, grad_scale=1.0):
        from oracle import regat_torch as ot
        inp = dict(features=np.asarray(features), boxes=np.asarray(boxes), q_att=np.asarray(q_att), q_last=np.asarray(q_last), target=np.asarray(target))
        loss, grads, _, _, _ = ot.loss_and_grads(self._named(self.params.numpy().copy()), self.cfg, inp)
        self.grads.zero_()
        for e in self.entries:
            self.grads[e.offset:e.offset + e.numel] = torch.tensor(grads[e.name].reshape(-1) * grad_scale)
        return {"loss": torch.tensor(loss)}

    def update(self, lr, step):
        from oracle import regat_torch as ot
        p, g = self.params.numpy(), self.grads.numpy()
        for e in self.entries:
            sl = slice(e.offset, e.offset + e.numel)
            w, self.m[sl], self.u[sl] = ot.adamax_step(p[sl], ot.clip_by_norm(g[sl], self.cfg.grad_clip), self.m[sl], self.u[sl], step, lr)
            p[sl] = w


def _oracle_setup():
    from tf_vqa_regat_b200 import synthetic as syn
    from tf_vqa_regat_b200.config import HotPathConfig
    cfg = HotPathConfig(v_dim=48, q_dim=24, rel_dim=32, num_heads=2, nongt_dim=5, num_answers=19)
    return cfg, syn.make_params(cfg, seed=7, trained_like=True), syn.make_inputs(cfg, 4, 7, seed=1000, adaptive=True)


def _oracle_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg, flat, inp = _oracle_setup()
    eng = OracleEngine(cfg, flat)
    tr = DataParallelTrainer(eng, bucket_elems=1000)
    tr.broadcast_params(0)
    b = shard_batch({k: v for k, v in inp.items() if k != "n_obj"}, rank, world)
    for step in range(2):
        tr.step(b["features"], b["boxes"], b["q_att"], b["q_last"], b["target"], lr=1e-3)
    out[rank] = eng.params.clone()
    dist.destroy_process_group()


def test_two_rank_gloo_on_the_oracle_path_equals_single_process():
    cfg, flat, inp = _oracle_setup()
    ref = OracleEngine(cfg, flat)
    for step in (1, 2):
        ref.fwd_bwd(inp["features"], inp["boxes"], inp["q_att"], inp["q_last"], inp["target"])
        ref.update(1e-3, step)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_oracle_worker, args=(2, 29519, out), nprocs=2, join=True)
    moved = (ref.params - torch.tensor(np.asarray(flat, dtype=np.float64))).abs().max().item()
    assert moved > 5e-4
    for r in range(2):
        # exactly-zero gradient directions (softmax shift invariance) take +-lr steps from rounding noise on either side: bound by
        # 2 steps x 2 lr there, demand agreement to a tiny fraction of a step everywhere else
        diff = (out[r] - ref.params).abs()
        assert diff.max().item() <= 4.1e-3
        assert (diff > 1e-7).float().mean().item() < 0.02
    torch.testing.assert_close(out[0], out[1], rtol=0, atol=0)          # both replicas hold the same parameters, bit for bit
