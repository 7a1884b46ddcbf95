"""TEST INFRASTRUCTURE -- CPU restatement (torch, autograd-able, float64 by default) of the explicit relation encoder, in the
re-associated form the kernels use.  Follows, under /root/reference/model:
    relation_encoder.py:95-143   ExplicitRelationEncoder.call: v2out (no dropout, relu), concat_visual_question, residual
    relation_encoder.py:13-37    concat_visual_question: mask = (row sum != 0)
    graph_att_net.py:53-83       self_weights, adj / adj^T per direction, [:, :, :nongt] slice, label FC, relu(s + sum_d o_d)
    graph_att_layer.py:39-121    Q, K on the first M objects, QK^T/sqrt(dh), where(adj > 0, aff, -9e15) + label_att, softmax over M,
                                 aggregation of the un-projected values, grouped 1x1 conv (re-associated: V' = s[:M] Kc + bc)
    weight_norm.py:41            W = g * v / ||v||_F
Pinned by tests/golden/refexec_explicit_*.npz (the reference's own files executed over oracle/tf_shim): tests/test_explicit.py."""
import torch


def _wn(p, name):
    v, g = p[name + "/v"], p[name + "/g"]
    return v * (g / torch.sqrt(torch.clamp((v * v).sum(), min=1e-12)))


def _fc(p, name, x):
    w = _wn(p, name)
    y = x @ w.reshape(-1, w.shape[-1])
    return y + p[name + "/bias"] if (name + "/bias") in p else y


def pair_bias(p, prefix, adj, nongt, dirs):
    """[B, dirs, N, M]: label FC of the labelled adjacency where an edge exists, -9e15 where none does (graph_att_layer.py:90-100)."""
    out = []
    for d in range(dirs):
        a = adj if d == 0 else adj.transpose(1, 2)                   # graph_att_net.py:56
        a = a[:, :, :nongt, :]                                       # :65
        lab = _fc(p, prefix + ".bias", a).squeeze(-1)                # :71
        out.append(torch.where(a.sum(-1) > 0, lab, torch.full_like(lab, -9e15)))
    return torch.stack(out, 1)


def forward(p, cfg, visual, adj, question):
    """p: name -> tensor with the golden files' variable names; cfg: dict(v_dim, q_dim, out_dim, dir_num, label_num, nongt_dim,
    num_heads, residual, label_bias[, num_steps = 1]).  Returns the encoder output [B, N, out_dim]."""
    D, H, dirs = cfg["out_dim"], cfg["num_heads"], cfg["dir_num"]
    dh = D // H
    v = torch.relu(_fc(p, "v2out", visual)) if cfg["v_dim"] != cfg["out_dim"] else visual
    B, N, _ = v.shape
    M = min(cfg["nongt_dim"], N)
    pre = "explicit_relation"
    pb = pair_bias(p, pre, adj, M, dirs)
    for _ in range(cfg.get("num_steps", 1)):                         # relation_encoder.py:134-141: same variables every step
        mask = (v.sum(-1) != 0).to(v.dtype)
        x = torch.cat([v, mask[..., None] * question[:, None, :]], -1)
        s = _fc(p, pre + ".self_weights", x)
        total = s
        for d in range(dirs):
            ln = f"{pre}.neighbor_net.{d}"
            q = _fc(p, ln + ".query", s).view(B, N, H, dh).transpose(1, 2)
            k = _fc(p, ln + ".key", s[:, :M]).view(B, M, H, dh).transpose(1, 2)
            kc = _wn(p, ln + ".linear_out_").reshape(D, D)           # Conv2D kernel [1,1,D,D]: head h owns output block h
            vp = (s[:, :M] @ kc + p[ln + ".linear_out_/bias"]).view(B, M, H, dh).transpose(1, 2)
            aff = q @ k.transpose(-1, -2) / dh ** 0.5
            live = pb[:, d][:, None] > -1e15
            logits = torch.where(live, aff + pb[:, d][:, None], pb[:, d][:, None].expand_as(aff))   # -9e15 absorbs aff and label in fp32/fp64 alike
            att = torch.softmax(logits, -1)
            total = total + (att @ vp).transpose(1, 2).reshape(B, N, D)
        out = torch.relu(total)
        v = v + out if cfg["residual"] else out
    return v
