"""The SAME math as oracle/regat_numpy.py, re-associated the way the CUDA kernels compute
it, exposing every intermediate a kernel produces so unit tests can
check kernels one at a time.  tests/test_oracle.py proves this file equals the
reference-formulation oracle; it is not itself the oracle of record.

Re-associations (all exact in real arithmetic):
  * W = g*v/||v||  ==>  x@W = alpha*(x@v), alpha = g/||v||            (weight_norm.py:41)
  * [v0 || mask*q] @ Ws = v0@Ws[:D] + mask*(q@Ws[D:])                  (relation_encoder.py:31-35)
  * (p @ s[:M]) per head through the grouped conv == p @ (s[:M] @ Kc[:, head block]);
    the conv bias folds into V' because softmax rows sum to 1          (graph_att_layer.py:110-117)
  * bias[i,j] uses box pair (f//N, f%N), f = i*M+j                     (graph_att_layer.py:74,81)
  * adj == 1 => tf.where is a no-op; label bias is one constant c      (graph_att_net.py:69-71)

TEST INFRASTRUCTURE.
"""
import numpy as np

from . import position_emb as pe


def alpha(p, name):
    v = p[name + "/v"]
    return p[name + "/g"] / np.sqrt(max(float(np.sum(np.square(v.astype(np.float64)))), 1e-12))


def pair_geometry(boxes, n_keys, feat_dim=64):
    """Emb[b, i, j, :] for the SCRAMBLED pair the attention layer uses: [B,N,M,E] (fp32 op order)."""
    B, N, _ = boxes.shape
    pos_emb = pe.prepare_graph_variables("implicit", boxes, None, None, N, n_keys, feat_dim, 11, 15)[0]
    return pos_emb.reshape(B, -1, feat_dim).reshape(B, N, min(n_keys, N), feat_dim)


def forward(p, cfg, features, boxes, q_att, q_last, target=None):
    f = features.dtype
    B, N, _ = features.shape
    D, H, dh = cfg.rel_dim, cfg.num_heads, cfg.head_dim
    M = min(cfg.nongt_dim, N)
    bias_of = lambda n: p.get(n + "/bias", 0.0)
    it = {}
    if cfg.v_dim != D:
        n = "v_relation.v2out"
        v0 = np.maximum(alpha(p, n) * (features @ p[n + "/v"]) + bias_of(n), 0)
    else:
        v0 = features
    mask = (np.sum(v0, -1) != 0).astype(f)
    pre = "v_relation.implicit_relation"
    n = pre + ".self_weights"
    ws = p[n + "/v"]
    qs = q_att @ ws[D:]                                                  # raw, alpha applied below
    s = alpha(p, n) * (v0 @ ws[:D] + mask[..., None] * qs[:, None, :]) + bias_of(n)
    n = pre + ".bias"
    c = alpha(p, n) * p[n + "/v"].reshape(()) + (p[n + "/bias"].reshape(()) if (n + "/bias") in p else 0.0)
    emb = pair_geometry(np.asarray(boxes, dtype=np.float32), cfg.nongt_dim, cfg.pos_emb_dim).astype(f)  # [B,N,M,E]
    acc = s.copy()
    it.update(v0=v0, mask=mask, qs=qs, s=s, c=c, emb=emb, dirs=[])
    for d in range(cfg.dir_num):
        ln = f"{pre}.neighbor_net.{d}"
        Q = alpha(p, ln + ".query") * (s @ p[ln + ".query/v"]) + bias_of(ln + ".query")
        K = alpha(p, ln + ".key") * (s[:, :M] @ p[ln + ".key/v"]) + bias_of(ln + ".key")
        Vp = alpha(p, ln + ".linear_out_") * (s[:, :M] @ p[ln + ".linear_out_/v"][0, 0]) + bias_of(ln + ".linear_out_")
        z = alpha(p, ln + ".pair_pos_fc") * (emb @ p[ln + ".pair_pos_fc/v"]) + bias_of(ln + ".pair_pos_fc")  # [B,N,M,H]
        gbias = np.log(np.maximum(np.maximum(z, 0), 1e-6)).transpose(0, 3, 1, 2)                           # [B,H,N,M]
        Qh = Q.reshape(B, N, H, dh).transpose(0, 2, 1, 3)
        Kh = K.reshape(B, M, H, dh).transpose(0, 2, 1, 3)
        Vh = Vp.reshape(B, M, H, dh).transpose(0, 2, 1, 3)
        L = (Qh @ Kh.transpose(0, 1, 3, 2)) * (1.0 / np.sqrt(dh)) + gbias + c
        L = L - L.max(-1, keepdims=True)
        P = np.exp(L); P = P / P.sum(-1, keepdims=True)                                                    # [B,H,N,M]
        O = (P @ Vh).transpose(0, 2, 1, 3).reshape(B, N, D)
        acc = acc + O
        it["dirs"].append(dict(Q=Q, K=K, Vp=Vp, z=z.transpose(0, 3, 1, 2), P=P, O=O))
    imp = np.maximum(acc, 0)
    v1 = v0 + imp if cfg.residual else imp
    n = "joint_emb"
    t = alpha(p, n + ".v2attention") * (v1 @ p[n + ".v2attention/v"]) + bias_of(n + ".v2attention")
    u = alpha(p, n + ".q2attention") * (q_last @ p[n + ".q2attention/v"]) + bias_of(n + ".q2attention")
    wl = alpha(p, n + ".linear") * p[n + ".linear/v"][:, 0]
    lg = np.einsum("bnc,bc,c->bn", t, u, wl) + bias_of(n + ".linear")
    lg = lg - lg.max(1, keepdims=True)
    aw = np.exp(lg); aw = aw / aw.sum(1, keepdims=True)
    pooled = np.einsum("bn,bnd->bd", aw, v1)
    pv = alpha(p, n + ".visual_embed") * (pooled @ p[n + ".visual_embed/v"]) + bias_of(n + ".visual_embed")
    qe = alpha(p, n + ".question_embed") * (q_last @ p[n + ".question_embed/v"]) + bias_of(n + ".question_embed")
    joint = pv * qe
    n = "classifier.layers"
    hid = np.maximum(alpha(p, n + ".0") * (joint @ p[n + ".0/v"]) + bias_of(n + ".0"), 0)
    logits = alpha(p, n + ".3") * (hid @ p[n + ".3/v"]) + bias_of(n + ".3")
    it.update(v1=v1, t=t, u=u, att_weights=aw[..., None], pooled=pooled, pv=pv, qe=qe, joint=joint, hid=hid,
              logits=logits)
    if target is not None:
        l = np.maximum(logits, 0) - logits * target + np.log1p(np.exp(-np.abs(logits)))
        it["loss"] = np.mean(l) * target.shape[1]
        it["dlogits"] = (1.0 / (1.0 + np.exp(-logits)) - target) / B
    return it
