"""Question front-end oracle: WordEmbedding -> GRU -> QuestionSelfAttention, the two tensors the hot path
consumes (q_emb_self_att, q_emb = last GRU state) and, through autograd, what the front-end does with the dq_att / dq_last
gradients the hot path returns.  torch-CPU restatement of /root/reference/model/language_model.py.

Pinned by tests/golden/refexec_question_*.npz: the reference's own language_model.py + rel_graph_net.py executed over
oracle/tf_shim (oracle/make_golden_ref_question.py).  TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference semantics kept, including the odd ones:
  * language_model.py:33-40   embedding rows of tokens equal to padding_idx (= n_token) are zeroed by a mask, not skipped;
  * :82-92                    op 'c': a second table `emb_` is concatenated on the feature axis (300 -> 600);
  * :105                      GRU dropout forced to 0; Keras GRU defaults otherwise (reset_after, gate order z|r|h, zero initial
                              state) and NO masking: padded positions are run through the GRU like any other input, so
                              call_last (:114-120, output[:, -1]) is the state after the trailing padding;
  * rel_graph_net.py:44,57    the GRU runs twice per step on the same input (q_emb(w_emb) and q_emb.call_last(w_emb));
  * :151-167                  the attention softmax is taken over the BATCH axis: logits [B,T,1] are squeezed, transposed to
                              [T,B], softmax(axis=1), and the [T,B] result is RAW-RESHAPED to [B,1,T] -- so sample b's weights are
                              the flat elements b*T .. b*T+T-1 of the [T,B] matrix, a mix of positions and samples.  (tf.squeeze
                              with B == 1 would also drop the batch axis; batch 1 is not supported here.)
"""
import numpy as np
import torch


def _wn(v, g):
    return v * torch.rsqrt(torch.clamp(v.pow(2).sum(), min=1e-12)) * g           # weight_norm.py:41


def _fc(x, p, name):
    y = torch.matmul(x, _wn(p[name + "/v"], p[name + "/g"]))
    b = p.get(name + "/bias")
    return y if b is None else y + b


def word_embedding(p, tokens, n_token, op="c"):
    """language_model.py:23-40,82-92.  tokens int64 [B,T]; tables [n_token+1, E]."""
    tok = torch.as_tensor(np.asarray(tokens), dtype=torch.long)
    mask = (tok != n_token).unsqueeze(-1)
    emb = p["w_emb.emb/emb"][tok] * mask
    if "c" in op:
        emb = torch.cat([emb, p["w_emb.emb_/emb_"][tok] * mask], dim=2)
    return emb


def gru(p, x, name="q_emb.gru"):
    """Keras GRU (reset_after=True): returns all states [B,T,u]."""
    W, U, b = p[name + "/kernel"], p[name + "/recurrent_kernel"], p[name + "/bias"]
    u = U.shape[0]
    B, T, _ = x.shape
    xi = torch.matmul(x, W) + b[0]                                              # input projections of every step at once
    h = torch.zeros(B, u, dtype=x.dtype)
    out = []
    for t in range(T):
        hi = torch.matmul(h, U) + b[1]
        z = torch.sigmoid(xi[:, t, :u] + hi[:, :u])
        r = torch.sigmoid(xi[:, t, u:2 * u] + hi[:, u:2 * u])
        c = torch.tanh(xi[:, t, 2 * u:] + r * hi[:, 2 * u:])
        h = z * h + (1 - z) * c
        out.append(h)
    return torch.stack(out, dim=1)


def question_self_attention(p, q_seq, name="q_att"):
    """language_model.py:151-173, batch-axis softmax and raw reshape included."""
    B, T, H = q_seq.shape
    if B == 1:
        raise ValueError("batch 1: tf.squeeze at language_model.py:159 would drop the batch axis")
    a1 = torch.tanh(_fc(q_seq, p, name + ".linear1"))
    logits = _fc(a1, p, name + ".linear2").squeeze(-1)                           # [B,T]
    w = torch.softmax(logits.t(), dim=1)                                         # [T,B], normalised over the batch
    w = w.reshape(B, 1, T)                                                       # raw reshape
    return torch.matmul(w, q_seq).reshape(B, H)


def forward(p, tokens, n_token, op="c"):
    """-> dict(w_emb [B,T,E'], q_seq [B,T,H], q_att [B,H], q_last [B,H])  (rel_graph_net.py:41-45,57)."""
    w = word_embedding(p, tokens, n_token, op)
    seq = gru(p, w)
    return dict(w_emb=w, q_seq=seq, q_att=question_self_attention(p, seq), q_last=gru(p, w)[:, -1])


# ---- layout of the front-end's trainable variables, Keras order (w_emb, q_emb, q_att; rel_graph_net.py:16-18)
def param_shapes(n_token, emb_dim, num_hid, op="c", emb2_trainable=False):
    """[(name, shape, trainable)].  emb_ is frozen unless tf-idf initialisation ran (language_model.py:58,79)."""
    e_in = emb_dim * (2 if "c" in op else 1)
    out = [("w_emb.emb/emb", (n_token + 1, emb_dim), True)]
    if "c" in op:
        out.append(("w_emb.emb_/emb_", (n_token + 1, emb_dim), emb2_trainable))
    out += [("q_emb.gru/kernel", (e_in, 3 * num_hid), True), ("q_emb.gru/recurrent_kernel", (num_hid, 3 * num_hid), True),
            ("q_emb.gru/bias", (2, 3 * num_hid), True),
            ("q_att.linear1/v", (num_hid, num_hid), True), ("q_att.linear1/g", (), True), ("q_att.linear1/bias", (num_hid,), True),
            ("q_att.linear2/v", (num_hid, 1), True), ("q_att.linear2/g", (), True), ("q_att.linear2/bias", (1,), True)]
    return out


def make_params(n_token, emb_dim, num_hid, op="c", seed=11):
    """Seeded synthetic front-end weights (fp32 values): N(0, 0.3) tables with a zero padding row, Glorot-like GRU kernels,
    small non-zero biases, g perturbed away from ||v|| so that W != v."""
    rng = np.random.default_rng(seed)
    p = {}
    for name, shape, _ in param_shapes(n_token, emb_dim, num_hid, op):
        if name.endswith("/g"):
            v = p[name[:-2] + "/v"]
            p[name] = np.float32(np.sqrt((v.astype(np.float64) ** 2).sum()) * rng.uniform(0.7, 1.4))
        elif "bias" in name:
            p[name] = (0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif "emb" in name:
            t = (0.3 * rng.standard_normal(shape)).astype(np.float32)
            t[-1] = 0.0
            p[name] = t
        else:
            lim = np.sqrt(6.0 / (shape[0] + shape[1]))
            p[name] = rng.uniform(-lim, lim, shape).astype(np.float32)
    return p


def make_tokens(batch, n_token, seq_len=14, seed=21):
    """dataset.py:250-263: token ids in [0, n_token), questions shorter than seq_len padded at the back with n_token."""
    rng = np.random.default_rng(seed)
    tok = rng.integers(0, n_token, (batch, seq_len))
    for b in range(batch):
        tok[b, int(rng.integers(3, seq_len + 1)):] = n_token
    tok[0, :] = rng.integers(0, n_token, seq_len)                               # one question of full length
    return tok.astype(np.int64)


def sparse_clip_adamax(w, tokens, values, m, u, step, lr, clip, beta1, beta2, eps, n_token):
    """train.py:112-113 for an embedding table: its tape gradient is tf.IndexedSlices (the variable is read through
    tf.nn.embedding_lookup, language_model.py:33) -- `values` [B,T,E] are the per-occurrence gradients of the MASKED lookup output
    (the gradient of the raw lookup output is that times the padding mask, applied here).
      tf.clip_by_norm(IndexedSlices): values * clip / max(||values||_2, clip), the norm over the un-deduplicated occurrences;
      Keras Adamax, sparse branch:    m <- beta1 m + (1-beta1) scatter_add(values); u <- beta2 u; every occurrence adds
                                      max(u[row], |value|) - u[row] of the DECAYED row it gathered; w -= lr_t m / (u + eps), all rows.
    NumPy float64; returns (w, m, u)."""
    tok = np.asarray(tokens).reshape(-1)
    vals = np.asarray(values, dtype=np.float64).reshape(tok.size, -1)
    vals = vals * (tok != n_token)[:, None]        # the mask multiplies the lookup's output (language_model.py:34-38): padding occurrences carry 0
    nrm = np.sqrt((vals * vals).sum())
    vals = vals * (clip / max(nrm, clip))
    m = beta1 * m
    np.add.at(m, tok, (1.0 - beta1) * vals)
    u = beta2 * u
    us = u[tok]
    np.add.at(u, tok, np.maximum(us, np.abs(vals)) - us)
    w = w - (lr / (1.0 - beta1 ** step)) * m / (u + eps)
    return w, m, u

