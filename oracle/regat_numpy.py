"""Stages 2-3 oracle, NumPy, in the REFERENCE FORMULATION (materialised pos_emb, raw
reshape scramble, un-projected 1024-d values, [B*N, H*D] grouped 1x1 conv, [v || mask*q]
concat).  Forward only; gradients come from oracle/regat_torch.py.

Follows, op for op:
  weight_norm.py:35-41, fc.py:17-48, relation_encoder.py:13-37,65-93,
  graph_att_net.py:40-83, graph_att_layer.py:39-121, fusion.py:22-54,
  classifier.py:14-25, train.py:20-26,107-108.
TensorFlow semantics restated (TF itself is absent):
  Dense on rank-3 = matmul over the last axis + bias;
  Conv2D NHWC groups=G, kernel [1,1,Cin/G,Cout]: group g maps input channels
    [g*Cin/G,(g+1)*Cin/G) to output channels [g*Cout/G,(g+1)*Cout/G);
  l2_normalize(x, axis=None) = x * rsqrt(max(sum(x^2), 1e-12));
  sigmoid_cross_entropy_with_logits = max(x,0) - x*z + log1p(exp(-|x|)).
Dropout layers are identities (train.py:104 never passes training=True).

Pinned by tests/golden/refexec_*.npz (the reference's own files executed over oracle/tf_shim;
see oracle/__init__.py for what that does and does not establish).  TEST INFRASTRUCTURE.
"""
import numpy as np

from . import position_emb as pe


def weight_norm(v, g):
    """weight_norm.py:41 -- whole-tensor (Frobenius) norm, scalar g."""
    ss = np.sum(np.square(v))
    return v * (1.0 / np.sqrt(np.maximum(ss, np.asarray(1e-12, dtype=v.dtype)))) * g


def wn_dense(x, p, name, act=None):
    """WeightNorm(Dense) (+ optional 'relu'); fc.py:36-43."""
    w = weight_norm(p[name + "/v"], p[name + "/g"])
    y = x @ w
    if (name + "/bias") in p:
        y = y + p[name + "/bias"]
    if act == "relu":
        y = np.maximum(y, 0)
    return y


def softmax(x, axis):
    x = x - np.max(x, axis=axis, keepdims=True)
    e = np.exp(x)
    return e / np.sum(e, axis=axis, keepdims=True)


def concat_visual_question(q, v):
    """relation_encoder.py:13-37: mask = (sum_d v != 0); x = [v || mask*q]."""
    B, N, _ = v.shape
    qb = np.broadcast_to(q[:, None, :], (B, N, q.shape[1]))
    mask = (np.sum(v, axis=-1) != 0).astype(v.dtype)[..., None]
    return np.concatenate([v, qb * mask], axis=-1), mask[..., 0]


def graph_self_attention_layer(p, pre, roi, adj, pos_emb, label_att, cfg):
    """graph_att_layer.py:39-121."""
    B, N, D = roi.shape
    H, dh = cfg.num_heads, cfg.head_dim
    M = cfg.nongt_dim if cfg.nongt_dim < N else N
    trunc = roi[:, :M]
    q = wn_dense(roi, p, pre + ".query").reshape(B, N, H, dh).transpose(0, 2, 1, 3)
    k = wn_dense(trunc, p, pre + ".key").reshape(B, M, H, dh).transpose(0, 2, 1, 3)
    value = trunc
    scale = np.asarray(1.0 / np.sqrt(np.float32(dh)), dtype=roi.dtype)
    aff = scale * (q @ k.transpose(0, 1, 3, 2))                       # [B,H,N,M]
    waff = aff.transpose(0, 2, 1, 3)                                  # [B,N,H,M]
    aux = {}
    if pos_emb is not None and cfg.pos_emb_dim > 0:
        e = pos_emb.reshape(B, -1, cfg.pos_emb_dim)                   # [B, M*N, E]
        z = wn_dense(e, p, pre + ".pair_pos_fc")                      # activation=None then tf.nn.relu
        pw = np.maximum(z, 0).reshape(B, -1, M, H).transpose(0, 1, 3, 2)   # raw reshape -> [B,N,H,M]
        pw = np.maximum(pw, np.asarray(1e-6, dtype=roi.dtype))
        waff = waff + np.log(pw)
        aux["z"] = z.reshape(B, -1, M, H).transpose(0, 1, 3, 2)
    if adj is not None:
        at = waff.transpose(0, 1, 3, 2)                               # [B,N,M,H]
        at = np.where(adj[..., None] > 0, at, np.asarray(-9e15, dtype=roi.dtype))
        at = at + label_att[..., None]
        waff = at.transpose(0, 1, 3, 2)
    prob = softmax(waff, axis=3)                                      # over M
    aux["prob"] = prob
    att = prob.reshape(B, N * H, M) @ value                           # [B, N*H, D]
    conv_in = att.reshape(B * N, H, D)                                # channels h*D + c
    w = weight_norm(p[pre + ".linear_out_/v"], p[pre + ".linear_out_/g"])[0, 0]   # [D, D]
    out = np.einsum("rhc,hco->rho", conv_in, w.reshape(D, H, dh).transpose(1, 0, 2))
    out = out.reshape(B * N, D) + p[pre + ".linear_out_/bias"]
    return out.reshape(B, N, D), aux


def graph_attention_network(p, x, adj_mat, pos_emb, cfg):
    """graph_att_net.py:40-83 with label_num = 1."""
    pre = "v_relation.implicit_relation"
    if cfg.pos_emb_dim > 0 and pos_emb is None:
        raise ValueError("position embedding is None with pos_emb_dim > 0")
    if cfg.pos_emb_dim < 0 and pos_emb is not None:
        raise ValueError("position embedding is NOT None with pos_emb_dim < 0")
    s = wn_dense(x, p, pre + ".self_weights")
    out = s
    adj_list = [adj_mat, adj_mat.transpose(0, 2, 1, 3)]
    auxes = []
    for d in range(cfg.dir_num):
        a = adj_list[d][:, :, :cfg.nongt_dim, :]
        cond = np.sum(a, axis=-1)
        lab = wn_dense(a, p, pre + ".bias")[..., 0]
        o, aux = graph_self_attention_layer(p, f"{pre}.neighbor_net.{d}", s, cond, pos_emb, lab, cfg)
        auxes.append(aux)
        out = out + o
    return np.maximum(out, 0), s, auxes


def implicit_relation_encoder(p, visual, pos_emb, question, cfg):
    """relation_encoder.py:65-93, num_steps = 1."""
    B, N, _ = visual.shape
    adj = np.ones((B, N, N, 1), dtype=visual.dtype)
    if cfg.v_dim != cfg.rel_dim:
        visual = wn_dense(visual, p, "v_relation.v2out", act="relu")
    x, mask = concat_visual_question(question, visual)
    imp, s, auxes = graph_attention_network(p, x, adj, pos_emb, cfg)
    v1 = visual + imp if cfg.residual else imp
    return v1, dict(v0=visual, mask=mask, s=s, att=auxes)


def butd(p, visual, question):
    """fusion.py:22-54 -- all five FCs are plain linear."""
    t = wn_dense(visual, p, "joint_emb.v2attention")
    u = wn_dense(question, p, "joint_emb.q2attention")
    logit = wn_dense(t * u[:, None, :], p, "joint_emb.linear")        # [B,N,1]
    w = softmax(logit, axis=1)
    pooled = np.sum(w * visual, axis=1)
    joint = wn_dense(pooled, p, "joint_emb.visual_embed") * wn_dense(question, p, "joint_emb.question_embed")
    return joint, w


def simple_classifier(p, x):
    """classifier.py:14-25."""
    return wn_dense(wn_dense(x, p, "classifier.layers.0", act="relu"), p, "classifier.layers.3")


def bce_loss(logits, target):
    """train.py:23,107-108: mean over B*A of BCE-with-logits, times A."""
    l = np.maximum(logits, 0) - logits * target + np.log1p(np.exp(-np.abs(logits)))
    return np.mean(l) * np.asarray(target.shape[1], dtype=logits.dtype)


def forward(p, cfg, features, boxes, q_att, q_last, target=None, pos_emb=None):
    """Hot path of rel_graph_net.py:53-62 (+ loss).  boxes are absolute-pixel `bb`."""
    if pos_emb is None:
        pos_emb = pe.prepare_graph_variables("implicit", boxes, None, None, features.shape[1],
                                             cfg.nongt_dim, cfg.pos_emb_dim, 11, 15)[0]
    pos_emb = pos_emb.astype(features.dtype)
    v1, aux = implicit_relation_encoder(p, features, pos_emb, q_att, cfg)
    joint, w = butd(p, v1, q_last)
    logits = simple_classifier(p, joint)
    out = dict(v1=v1, joint=joint, att_weights=w, logits=logits, pos_emb=pos_emb, **aux)
    if target is not None:
        out["loss"] = bce_loss(logits, target)
    return out
