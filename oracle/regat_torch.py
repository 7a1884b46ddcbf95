"""Stages 2-3 oracle on torch-CPU: same reference formulation as oracle/regat_numpy.py
but written against torch ops so that (a) autograd supplies the gradients the
reference gets from tf.GradientTape (train.py:103-111), (b) it is an independent second
transcription to cross-check the NumPy one, and (c) bench.py can time it on all host
cores as the "reference-formulation CPU restatement" baseline.

Also restates the train-step glue: per-tensor clip_by_norm (train.py:112) and Keras
Adamax (train.py:48,113).

Pinned by tests/golden/refexec_*.npz (the reference's own files executed over oracle/tf_shim;
see oracle/__init__.py for what that does and does not establish).  TEST INFRASTRUCTURE.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import position_emb as pe


def _wn(v, g):
    # weight_norm.py:41: l2_normalize(v, axis=None) * g
    return v * torch.rsqrt(torch.clamp(v.pow(2).sum(), min=1e-12)) * g


def _fc(x, p, name, relu=False):
    y = torch.matmul(x, _wn(p[name + "/v"], p[name + "/g"]))
    b = p.get(name + "/bias")
    if b is not None:
        y = y + b
    return torch.relu(y) if relu else y


def _att_layer(p, pre, roi, adj, pos_emb, label_att, cfg):
    # graph_att_layer.py:39-121
    B, N, D = roi.shape
    H, dh = cfg.num_heads, cfg.head_dim
    M = min(cfg.nongt_dim, N)
    trunc = roi[:, :M, :]
    q = _fc(roi, p, pre + ".query").view(B, N, H, dh).permute(0, 2, 1, 3)
    k = _fc(trunc, p, pre + ".key").view(B, M, H, dh).permute(0, 2, 1, 3)
    aff = torch.matmul(q, k.transpose(-1, -2)) * (1.0 / float(np.sqrt(np.float32(dh))))
    waff = aff.permute(0, 2, 1, 3)                                    # [B,N,H,M]
    if pos_emb is not None and cfg.pos_emb_dim > 0:
        e = pos_emb.reshape(B, -1, cfg.pos_emb_dim)
        pw = torch.relu(_fc(e, p, pre + ".pair_pos_fc")).reshape(B, -1, M, H).permute(0, 1, 3, 2)
        waff = waff + torch.log(torch.clamp(pw, min=1e-6))
    if adj is not None:
        at = waff.permute(0, 1, 3, 2)
        at = torch.where(adj.unsqueeze(-1) > 0, at, torch.full_like(at, -9e15))
        waff = (at + label_att.unsqueeze(3)).permute(0, 1, 3, 2)
    prob = torch.softmax(waff, dim=3)
    att = torch.matmul(prob.reshape(B, N * H, M), trunc)              # un-projected values
    conv_in = att.reshape(B * N, H * D, 1, 1)                         # NCHW view of [B*N,1,1,H*D]
    w = _wn(p[pre + ".linear_out_/v"], p[pre + ".linear_out_/g"])     # [1,1,D,D] = [kh,kw,cin/g,cout]
    w_oihw = w.permute(3, 2, 0, 1).contiguous()                       # [cout, cin/groups, 1, 1]
    out = F.conv2d(conv_in, w_oihw, p[pre + ".linear_out_/bias"], groups=H)
    return out.reshape(B, N, D)


def encoder(p, cfg, features, pos_emb, q_att, num_steps=1):
    # relation_encoder.py:65-93 + graph_att_net.py:40-83; num_steps > 1 repeats :82-91 on the running `visual`, same variables
    B, N, _ = features.shape
    adj = torch.ones(B, N, N, 1, dtype=features.dtype)
    visual = _fc(features, p, "v_relation.v2out", relu=True) if cfg.v_dim != cfg.rel_dim else features
    pre = "v_relation.implicit_relation"
    adjs = [adj, adj.permute(0, 2, 1, 3)]
    for _ in range(num_steps):
        mask = (visual.sum(-1) != 0).to(visual.dtype).unsqueeze(-1)
        x = torch.cat([visual, q_att.unsqueeze(1).expand(B, N, -1) * mask], dim=-1)
        s = _fc(x, p, pre + ".self_weights")
        out = s
        for d in range(cfg.dir_num):
            a = adjs[d][:, :, :cfg.nongt_dim, :]
            lab = _fc(a, p, pre + ".bias").squeeze(-1)
            out = out + _att_layer(p, f"{pre}.neighbor_net.{d}", s, a.sum(-1), pos_emb, lab, cfg)
        imp = torch.relu(out)
        visual = (visual + imp) if cfg.residual else imp
    return visual


def head(p, v1, q_last):
    # fusion.py:22-54, classifier.py:14-25
    t = _fc(v1, p, "joint_emb.v2attention")
    u = _fc(q_last, p, "joint_emb.q2attention")
    w = torch.softmax(_fc(t * u.unsqueeze(1), p, "joint_emb.linear"), dim=1)
    pooled = (w * v1).sum(1)
    joint = _fc(pooled, p, "joint_emb.visual_embed") * _fc(q_last, p, "joint_emb.question_embed")
    logits = _fc(_fc(joint, p, "classifier.layers.0", relu=True), p, "classifier.layers.3")
    return logits, joint, w


def loss_fn(logits, target):
    # train.py:23,107-108
    return F.binary_cross_entropy_with_logits(logits, target, reduction="mean") * target.shape[1]


def to_torch_params(named_np, dtype=torch.float64, requires_grad=True):
    return {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=requires_grad) for k, v in named_np.items()}


def forward(p, cfg, features, boxes, q_att, q_last, target=None, pos_emb=None):
    """features/boxes/... are NumPy or torch; boxes go through the stage-1 NumPy oracle in fp32
    (as the reference does on the host, train.py:97) and are then cast to the compute dtype."""
    dt = next(iter(p.values())).dtype
    tt = lambda a: a if isinstance(a, torch.Tensor) else torch.tensor(np.asarray(a), dtype=dt)
    if pos_emb is None:
        pos_emb = pe.prepare_graph_variables("implicit", np.asarray(boxes, dtype=np.float32), None, None,
                                             features.shape[1], cfg.nongt_dim, cfg.pos_emb_dim, 11, 15)[0]
    features, q_att, q_last, pos_emb = tt(features), tt(q_att), tt(q_last), tt(pos_emb)
    v1 = encoder(p, cfg, features, pos_emb, q_att)
    logits, joint, w = head(p, v1, q_last)
    out = dict(v1=v1, joint=joint, att_weights=w, logits=logits)
    if target is not None:
        out["loss"] = loss_fn(logits, tt(target))
    return out


def loss_and_grads(named_np, cfg, inputs, dtype=torch.float64):
    """One GradientTape-equivalent: returns (loss, {name: grad}, dq_att, dq_last, outputs)."""
    p = to_torch_params(named_np, dtype)
    q_att = torch.tensor(inputs["q_att"], dtype=dtype, requires_grad=True)
    q_last = torch.tensor(inputs["q_last"], dtype=dtype, requires_grad=True)
    out = forward(p, cfg, inputs["features"], inputs["boxes"], q_att, q_last, inputs["target"])
    out["loss"].backward()
    grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in p.items()}
    return float(out["loss"].detach()), grads, q_att.grad.numpy(), q_last.grad.numpy(), \
        {k: v.detach().numpy() for k, v in out.items()}


def clip_by_norm(g, clip):
    """tf.clip_by_norm, train.py:112: g * clip / max(||g||_2, clip)."""
    n = np.sqrt(np.sum(np.square(g.astype(np.float64))))
    return g * (clip / max(n, clip))


def adamax_step(w, g, m, u, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """Keras Adamax (train.py:48): m=b1*m+(1-b1)g; u=max(b2*u,|g|); w -= lr/(1-b1^t) * m/(u+eps)."""
    m = beta1 * m + (1.0 - beta1) * g
    u = np.maximum(beta2 * u, np.abs(g))
    w = w - (lr / (1.0 - beta1 ** step)) * m / (u + eps)
    return w, m, u
