"""Seeded synthetic inputs and weights of the hot path's shapes.

Host-side NumPy only.  Shared by tests, bench.py and smoke(); the same arrays feed the
CUDA path and the oracle, so a run is its own parity check.
"""
import numpy as np

from .config import HotPathConfig, param_layout


def make_inputs(cfg: HotPathConfig, batch: int, n_rois: int, seed: int = 1001, adaptive: bool = False):
    """features [B,N,V] >= 0, boxes [B,N,4] abs pixels (x1,y1,x2,y2) in 640x480,
    q_att/q_last [B,Q], targets [B,A] soft scores; adaptive => per-graph object counts
    in 10..n_rois with zero post-padding of features AND boxes (dataset.py:329-355)."""
    rng = np.random.default_rng(seed)
    B, N = batch, n_rois
    feat = np.maximum(rng.standard_normal((B, N, cfg.v_dim), dtype=np.float32), 0.0)
    x1 = rng.uniform(0, 560, (B, N)); y1 = rng.uniform(0, 400, (B, N))
    w = rng.uniform(8, 320, (B, N)); h = rng.uniform(8, 240, (B, N))
    boxes = np.stack([x1, y1, np.minimum(x1 + w, 639), np.minimum(y1 + h, 479)], -1).astype(np.float32)
    q_att = (0.5 * rng.standard_normal((B, cfg.q_dim))).astype(np.float32)
    q_last = (0.5 * rng.standard_normal((B, cfg.q_dim))).astype(np.float32)
    target = np.zeros((B, cfg.num_answers), dtype=np.float32)
    scores = np.array([0.3, 0.6, 0.9, 1.0], dtype=np.float32)     # dataset.py:314-318
    for b in range(B):
        k = int(rng.integers(1, 11))
        idx = rng.choice(cfg.num_answers, size=k, replace=False)
        target[b, idx] = scores[rng.integers(0, 4, size=k)]
    n_obj = np.full((B,), N, dtype=np.int32)
    if adaptive:
        lo = 10 if N >= 20 else max(1, N // 2)
        n_obj = rng.integers(lo, N + 1, size=B).astype(np.int32)
        n_obj[int(rng.integers(0, B))] = N                         # someone defines the padded width
        for b in range(B):
            feat[b, n_obj[b]:] = 0.0
            boxes[b, n_obj[b]:] = 0.0
    return dict(features=feat, boxes=boxes, q_att=q_att, q_last=q_last, target=target, n_obj=n_obj)


def make_params(cfg: HotPathConfig, seed: int = 7, trained_like: bool = False) -> np.ndarray:
    """Flat fp32 parameter buffer in config.param_layout order.  Keras defaults: v
    Glorot-uniform, bias zeros, g = ||v||_F (weight_norm.py:25,35-37).  trained_like
    perturbs g and draws biases ~0.1*N(0,1) so that padded rows stop being exactly zero
    (flips the relation_encoder.py:20-21 mask) and W != v."""
    rng = np.random.default_rng(seed)
    entries, total = param_layout(cfg)
    flat = np.zeros((total,), dtype=np.float32)
    norms = {}
    for e in entries:
        if e.kind == "v":
            if len(e.shape) == 4:                       # Conv2D kernel [1,1,cin/groups,cout]
                fan_in, fan_out = e.shape[2], e.shape[3]
            else:
                fan_in, fan_out = e.shape[0], e.shape[1]
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            v = rng.uniform(-lim, lim, e.numel).astype(np.float32)
            if e.numel == 1:
                v = np.abs(v) + 0.5                     # keep the 1x1 label FC away from ||v||=0
            flat[e.offset:e.offset + e.numel] = v
            norms[e.layer] = float(np.sqrt(np.sum(v.astype(np.float64) ** 2)))
        elif e.kind == "g":
            g = norms[e.layer]
            if trained_like:
                g *= float(rng.uniform(0.7, 1.4))
            flat[e.offset] = g
        elif trained_like:
            flat[e.offset:e.offset + e.numel] = 0.1 * rng.standard_normal(e.numel).astype(np.float32)
    return flat


def unflatten(cfg: HotPathConfig, flat: np.ndarray) -> dict:
    """name -> array views (g as 0-d)."""
    out = {}
    for e in param_layout(cfg)[0]:
        out[e.name] = flat[e.offset:e.offset + e.numel].reshape(e.shape)
    return out


# ---- explicit relation encoders: seeded inputs and weights shared by the golden generator, the oracle and the tests
def make_explicit_inputs(v_dim, q_dim, label_num, B, N, seed):
    """visual >= 0 with zero rows for padded objects, question, and a one-hot labelled adjacency [B,N,N,L] in which padded objects
    have no edges at all (rows AND columns empty: their softmax rows are fully masked) and ~40 % of the real pairs are linked."""
    rng = np.random.default_rng(seed)
    visual = np.maximum(rng.standard_normal((B, N, v_dim)), 0.0)
    question = 0.5 * rng.standard_normal((B, q_dim))
    n_obj = rng.integers(max(2, N // 2), N + 1, size=B)
    n_obj[0] = N
    adj = np.zeros((B, N, N, label_num))
    for b in range(B):
        visual[b, n_obj[b]:] = 0.0
        for i in range(n_obj[b]):
            for j in range(n_obj[b]):
                if i == j or rng.random() < 0.4:
                    adj[b, i, j, rng.integers(0, label_num)] = 1.0
    return visual, question, adj, n_obj


def explicit_param_values(shapes, seed):
    """Trained-like values for a list of variable shapes in Keras order (v, g, [bias] per WeightNorm): v Glorot-uniform, g = ||v|| times
    U(0.7, 1.4), bias 0.1 N(0,1).  float64; the same call rebuilds the parameters the golden generator assigned."""
    rng = np.random.default_rng(seed)
    out, last_norm = [], 1.0
    for shp in shapes:
        shp = tuple(int(x) for x in shp)
        if len(shp) >= 2:
            fan_in, fan_out = (shp[2], shp[3]) if len(shp) == 4 else (shp[0], shp[1])
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            v = rng.uniform(-lim, lim, shp)
            last_norm = float(np.sqrt((v * v).sum()))
            out.append(v)
        elif len(shp) == 0:
            out.append(np.asarray(last_norm * rng.uniform(0.7, 1.4)))
        else:
            out.append(0.1 * rng.standard_normal(shp))
    return out
