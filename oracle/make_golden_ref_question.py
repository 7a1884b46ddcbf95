"""Golden vectors for the question front-end and for the WHOLE model, tokens in -> logits out, produced by
executing the reference's own files over oracle/tf_shim (see oracle/make_golden_ref.py for what that means and assumes).
Build container only:

    python -m oracle.make_golden_ref_question        # rewrites tests/golden/refexec_question_*.npz

Runs unmodified from /root/reference: model/language_model.py (WordEmbedding, QuestionEmbedding, QuestionSelfAttention),
model/rel_graph_net.py (RelationGraphAttentionNetwork.call: the GRU is run twice, :44 and :57), every hot-path file, and
train.train() / train.evaluate().  Restated in the stand-in: keras.layers.GRU (checked against torch.nn.GRU in
tests/test_refexec_question.py), tf.nn.embedding_lookup, and the primitives make_golden_ref.py already lists.
"""
import os
import sys
import tempfile
import types

import numpy as np

from .make_golden_ref import GOLD, LR, SMALL, TINY, _import_reference, _norm_name

CASES = {
    # name: (hot-path cfg kwargs, B, N, adaptive, n_token, emb_dim, op, emb_ trainable (tf-idf init ran, language_model.py:79), steps)
    "question_tiny": (TINY, 3, 9, True, 40, 10, "c", False, 2),
    "question_small": (SMALL, 4, 36, True, 60, 12, "c", True, 2),
    "question_small_noconcat": (SMALL, 2, 20, False, 60, 12, "", False, 2),
    # a vocabulary of 6 words: every batch repeats every token many times, so the IndexedSlices semantics of the embedding
    # gradient (clip by the un-deduplicated norm, per-occurrence Adamax `u` increments) differ visibly from the dense ones
    "question_tiny_repeats": (TINY, 3, 9, True, 6, 10, "c", True, 3),
}


def run_case(name, mods, save=True):
    tf, ref_train, ref_enc, BUTD, SimpleClassifier, Net = mods
    from model.language_model import QuestionEmbedding, QuestionSelfAttention, WordEmbedding
    from oracle import language_model as olm
    from tf_vqa_regat_b200 import synthetic as syn
    from tf_vqa_regat_b200.config import HotPathConfig, param_layout
    kw, B, N, adaptive, n_token, emb_dim, op, emb2_tr, steps = CASES[name]
    cfg = HotPathConfig(**kw)
    entries, _ = param_layout(cfg)
    hot = syn.unflatten(cfg, syn.make_params(cfg, seed=7, trained_like=True))
    front = olm.make_params(n_token, emb_dim, cfg.q_dim, op, seed=11)
    fshapes = olm.param_shapes(n_token, emb_dim, cfg.q_dim, op, emb2_tr)
    batches = [syn.make_inputs(cfg, B, N, seed=1000 + s, adaptive=adaptive) for s in range(steps + 1)]
    tokens = [olm.make_tokens(B, n_token, 14, seed=21 + s) for s in range(steps + 1)]
    out = dict(cfg=np.array(repr(kw)), B=B, N=N, adaptive=adaptive, n_token=n_token, emb_dim=emb_dim, op=np.array(op),
               emb2_trainable=emb2_tr, steps=steps, lr=LR, tokens=np.stack(tokens),
               input_check=np.array([float(np.sum(b["features"], dtype=np.float64)) for b in batches]))

    def fresh_model():
        tf.keras.backend.set_floatx("float64")
        tf.random.set_seed(0)
        w_emb = WordEmbedding(n_token, emb_dim, 0.2, op)                                     # rel_graph_net.py:71
        q_emb = QuestionEmbedding(emb_dim if "c" not in op else 2 * emb_dim, cfg.q_dim, 1, False, 0.2)   # :72-73
        q_att = QuestionSelfAttention(cfg.q_dim, 0.2)                                        # :74
        v_relation = ref_enc.ImplicitRelationEncoder(cfg.v_dim, cfg.q_dim, cfg.rel_dim, cfg.dir_num, cfg.pos_emb_dim, cfg.nongt_dim,
                                                     num_heads=cfg.num_heads, num_steps=1, residual_connection=cfg.residual,
                                                     label_bias=cfg.label_bias)
        model = Net(w_emb, q_emb, q_att, v_relation, BUTD(cfg.rel_dim, cfg.q_dim, cfg.q_dim),
                    SimpleClassifier(cfg.q_dim, cfg.q_dim * 2, cfg.num_answers, 0.2), "butd", "implicit")
        b0 = batches[0]
        pos, _, _ = ref_train.prepare_graph_variables("implicit", b0["boxes"], None, None, N, cfg.nongt_dim, cfg.pos_emb_dim, 11, 15)
        model(b0["features"], None, tf.constant(tokens[0]), pos, None, None)                 # builds every variable
        if emb2_tr and "c" in op:
            model.w_emb.emb_.trainable = True                                                # language_model.py:79
        nw = model.named_weights()
        want = [n for n, _, _ in fshapes] + [e.name for e in entries]
        assert [_norm_name(n) for n, _ in nw] == want, [(a, b) for (a, _), b in zip(nw, want) if _norm_name(a) != b][:4]
        assert [tuple(w.shape) for _, w in nw] == [tuple(s) for _, s, _ in fshapes] + [tuple(e.shape) for e in entries]
        model.set_weights([front[n] for n, _, _ in fshapes] + [hot[e.name] for e in entries])
        tr = [_norm_name(n) for n, w in nw if any(w is t for t in model.trainable_variables)]
        assert tr == [n for n, _, t in fshapes if t] + [e.name for e in entries], tr
        return model, pos

    # ---- forward through the reference's model, intermediates of the front-end, one GradientTape
    model, pos = fresh_model()
    b0, target = batches[0], tf.convert_to_tensor(batches[0]["target"])
    tok = tf.constant(tokens[0])
    with tf.GradientTape() as tape:
        w_emb = model.w_emb(tok)                                                             # rel_graph_net.py:41
        q_seq = model.q_emb(w_emb)                                                           # :44
        q_att = model.q_att(q_seq)                                                           # :45
        q_last = model.q_emb.call_last(w_emb)                                                # :57
        logits = model(b0["features"], None, tok, pos, None, None)
        loss = tf.reduce_mean(ref_train.instance_bce_with_logits(logits, target)) * tf.cast(tf.shape(target)[1], tf.float32)
    tv = list(model.trainable_variables)
    grads = tape.gradient(loss, tv)
    out.update(w_emb=w_emb.numpy(), q_seq=q_seq.numpy(), q_att=q_att.numpy(), q_last=q_last.numpy(), logits=logits.numpy(),
               loss=loss.numpy())
    names = [n for n, _, t in fshapes if t] + [e.name for e in entries]
    for n, g in zip(names, grads):
        if n.startswith("w_emb"):
            assert type(g).__name__ == "IndexedSlices", "the embedding tables' gradient must come back sparse"
            out["grad/" + n] = g.numpy()                                                     # densified (duplicates summed)
            out["grad_occurrence_norm/" + n] = float(np.sqrt((g.values.numpy() ** 2).sum()))  # what tf.clip_by_norm divides by
        elif n.startswith(("q_emb", "q_att")):
            out["grad/" + n] = g.numpy()
        else:
            out["gradnorm/" + n] = float(np.sqrt((g.numpy() ** 2).sum()))

    # ---- the reference's train() / evaluate() on token batches
    model, _ = fresh_model()
    step_logits = []
    orig_call = type(model).call

    class Recording(type(model)):
        def call(self, *a, **k):
            r = orig_call(self, *a, **k)
            step_logits.append(r.numpy().copy())
            return r
    model.__class__ = Recording

    class Loader:
        relation_type = "implicit"

        def __init__(self, idx):
            self.idx, self.data_loader_len, self.num_total_data = idx, len(idx), B * len(idx)

        def generator(self):
            for i in self.idx:
                b = batches[i]
                yield (b["features"], None, tf.constant(tokens[i]), b["boxes"], np.zeros((B, 1)), np.zeros((B, 1)),
                       tf.convert_to_tensor(b["target"]))

    with tempfile.TemporaryDirectory() as tmp:
        args = types.SimpleNamespace(base_lr=LR, epochs=1, lr_decay_step=2, lr_decay_rate=0.25, grad_clip=cfg.grad_clip,
                                     output=tmp + "/", relation_type="implicit", nongt_dim=cfg.nongt_dim,
                                     imp_pos_emb_dim=cfg.pos_emb_dim, spa_label_num=11, sem_label_num=15, print_freq=500)
        stdout, sys.stdout = sys.stdout, open(os.devnull, "w")
        try:
            ref_train.train(model, Loader(list(range(steps))), Loader([steps]), args)
        finally:
            sys.stdout.close(); sys.stdout = stdout
    assert len(step_logits) == steps + 1 and np.array_equal(step_logits[0], out["logits"])
    out["train.logits"] = np.stack(step_logits[:steps]); out["eval.logits"] = step_logits[steps]
    for (n, _), w in zip(model.named_weights(), model.weights):
        n = _norm_name(n)
        if n.startswith(("w_emb", "q_emb", "q_att")):
            out["param/" + n] = w.numpy()
    if save:
        np.savez_compressed(os.path.join(GOLD, f"refexec_{name}.npz"), **out)
    print(f"refexec {name}: loss {float(out['loss']):.6f}, |q_att| {np.abs(out['q_att']).max():.3f}, |q_last| {np.abs(out['q_last']).max():.3f}")
    return out


if __name__ == "__main__":
    mods = _import_reference()
    for name in (sys.argv[1:] or CASES):
        run_case(name, mods)
