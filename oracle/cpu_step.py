"""The reference's train step on the CPU, timed by bench.py as the reported baseline:
host NumPy position embedding (train.py:97) -> forward in the reference formulation -> autograd
(train.py:103-111) -> per-tensor clip_by_norm (train.py:112) -> Keras Adamax (train.py:113),
fp32, torch-CPU with all host threads.  TensorFlow is not installable here, so this restatement IS the
"reference CPU path" (BASELINE.md section 3).  TEST INFRASTRUCTURE / reported baseline only."""
import os
import time

import numpy as np
import torch

from . import position_emb as pe
from . import regat_torch as ot


class CpuTrainer:
    def __init__(self, cfg, flat_params, unflatten, threads=None):
        self.cfg = cfg
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        named = unflatten(cfg, np.array(flat_params, dtype=np.float32))
        self.p = {k: torch.tensor(np.array(v), dtype=torch.float32, requires_grad=True) for k, v in named.items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.u = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.t = 0

    def step(self, inp, lr=1e-3):
        cfg = self.cfg
        pos_emb = pe.prepare_graph_variables("implicit", inp["boxes"], None, None, inp["features"].shape[1], cfg.nongt_dim,
                                             cfg.pos_emb_dim, 11, 15)[0]
        out = ot.forward(self.p, cfg, torch.from_numpy(inp["features"]), inp["boxes"], torch.from_numpy(inp["q_att"]),
                         torch.from_numpy(inp["q_last"]), torch.from_numpy(inp["target"]), pos_emb=torch.from_numpy(pos_emb))
        for v in self.p.values():
            v.grad = None
        out["loss"].backward()
        self.t_fb = time.perf_counter()
        self.t += 1
        with torch.no_grad():
            for k, w in self.p.items():
                g = w.grad if w.grad is not None else torch.zeros_like(w)
                g = g * (cfg.grad_clip / max(float(g.norm()), cfg.grad_clip))
                self.m[k].mul_(cfg.beta1).add_(g, alpha=1 - cfg.beta1)
                torch.maximum(self.u[k] * cfg.beta2, g.abs(), out=self.u[k])
                w.sub_((lr / (1 - cfg.beta1 ** self.t)) * self.m[k] / (self.u[k] + cfg.eps))
        return float(out["loss"].detach())


def time_cpu_train(cfg, make_inputs, make_params, unflatten, sample_batch, n_rois, full_batch=256, steps=3, warmup=1,
                   budget_s=40.0):
    """Times the CPU step on `sample_batch` graphs and extrapolates to a full step of `full_batch` graphs:
        t_step(full) = t_fwd_bwd(sample) * full/sample + t_clip_adamax      (the optimizer does not scale with the batch)
    Returns (graphs_per_s, seconds_per_full_step, threads, steps_timed)."""
    tr = CpuTrainer(cfg, make_params(cfg, seed=7, trained_like=True), unflatten)
    inp = make_inputs(cfg, sample_batch, n_rois, seed=1001)
    t_begin = time.perf_counter()
    for _ in range(warmup):
        tr.step(inp)
    best = None
    n = 0
    for _ in range(steps):
        t0 = time.perf_counter()
        tr.step(inp)
        t1 = time.perf_counter()
        full = (tr.t_fb - t0) * full_batch / sample_batch + (t1 - tr.t_fb)
        best = full if best is None else min(best, full)
        n += 1
        if time.perf_counter() - t_begin > budget_s:
            break
    return full_batch / best, best, tr.threads, n


def time_cpu_train_full(cfg, make_inputs, make_params, unflatten, batch, n_rois, steps=5, warmup=2, budget_s=120.0, adaptive=False):
    """The WHOLE step on `batch` graphs (no sampling, no extrapolation): `warmup` untimed steps, then best of up to `steps`
    (stops early once `budget_s` of wall clock is spent, but always times at least one).  Returns
    (graphs_per_s, seconds_per_step, threads, steps_timed, warmups_run)."""
    tr = CpuTrainer(cfg, make_params(cfg, seed=7, trained_like=True), unflatten)
    inp = make_inputs(cfg, batch, n_rois, seed=1001, adaptive=adaptive)
    t_begin = time.perf_counter()
    w = 0
    for _ in range(warmup):
        tr.step(inp)
        w += 1
        if time.perf_counter() - t_begin > budget_s / 2:
            break
    best, n = None, 0
    for _ in range(steps):
        t0 = time.perf_counter()
        tr.step(inp)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        n += 1
        if time.perf_counter() - t_begin > budget_s:
            break
    return batch / best, best, tr.threads, n, w


def time_cpu_forward(cfg, make_inputs, make_params, unflatten, batch, n_rois, steps=5, warmup=2, adaptive=False):
    """Forward only (host position embedding + reference-formulation forward, no autograd): BASELINE.json configs[0] is this at
    batch 4, K=36.  Best of `steps` after `warmup`.  Returns (graphs_per_s, seconds, threads)."""
    tr = CpuTrainer(cfg, make_params(cfg, seed=7, trained_like=True), unflatten)
    inp = make_inputs(cfg, batch, n_rois, seed=1001, adaptive=adaptive)
    f, q1, q2, tg = (torch.from_numpy(inp[k]) for k in ("features", "q_att", "q_last", "target"))

    def fwd():
        pos_emb = pe.prepare_graph_variables("implicit", inp["boxes"], None, None, n_rois, cfg.nongt_dim, cfg.pos_emb_dim, 11, 15)[0]
        with torch.no_grad():
            return ot.forward(tr.p, cfg, f, inp["boxes"], q1, q2, tg, pos_emb=torch.from_numpy(pos_emb))["logits"]

    for _ in range(warmup):
        fwd()
    best = None
    for _ in range(steps):
        t0 = time.perf_counter()
        fwd()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return batch / best, best, tr.threads
