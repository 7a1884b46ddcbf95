"""Golden vectors for the explicit relation encoders, produced by EXECUTING the reference's own
model/relation_encoder.py (ExplicitRelationEncoder, :95-143), graph_att_net.py, graph_att_layer.py, fc.py and weight_norm.py
over oracle/tf_shim, plus model/position_emb.py:23-90 (build_graph, pure NumPy, runs as it is).  Build container only:

    python -m oracle.make_golden_ref_explicit        # rewrites tests/golden/refexec_explicit_*.npz, explicit_build_graph.npz

ONE substitution beyond the TensorFlow stand-in, because the reference's class cannot be constructed as written: its
constructor spells the argument `residiual_connection` (relation_encoder.py:98) and then reads the undefined name
`residual_connection` (:104) -- a NameError.  The generator defines that ONE module-level name in the imported module
(`relation_encoder.residual_connection = <flag>`) before constructing; nothing else is touched.  The call sites of the class
(rel_graph_net.py:79-92) are broken too (`residual_connection=` keyword, `arg.relation_dim`), so the class is driven directly.

Recorded per case: inputs' checksums (inputs come from the seeded generator below), the encoder output, and the gradients of
sum(output * probe) w.r.t. every variable, the visual features and the question (GradientTape over the stand-in)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"

CASES = {
    # name: (v_dim, q_dim, out_dim, dir_num, label_num, nongt, heads, B, N, residual, label_bias)
    "spatial_small": (64, 32, 256, 2, 11, 20, 4, 3, 36, True, True),
    "semantic_dir1_nores": (512, 64, 512, 1, 15, 20, 8, 2, 24, False, False),           # v_dim == out_dim: no v2out; one direction
    "spatial_clamped_n12": (64, 32, 256, 2, 11, 20, 4, 2, 12, True, True),               # N < nongt_dim
    # trailing element = num_steps (relation_encoder.py:134-141: the q-mask, the network and the residual add repeat on `visual`)
    "spatial_small_steps2": (64, 32, 256, 2, 11, 20, 4, 2, 36, True, True, 2),
    "semantic_nores_steps3": (64, 32, 256, 2, 15, 20, 4, 2, 24, False, True, 3),
}


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    import tensorflow as tf
    assert tf.__version__.endswith("standin")
    tf.keras.backend.set_floatx("float64")
    from model import relation_encoder as ref_enc
    from model import position_emb as ref_pe
    from oracle.make_golden_ref import _norm_name

    for name, case in CASES.items():
        v_dim, q_dim, out_dim, dirs, L, nongt, heads, B, N, residual, label_bias = case[:11]
        steps = case[11] if len(case) > 11 else 1
        ref_enc.residual_connection = residual            # the one-name shim described in the module docstring
        enc = ref_enc.ExplicitRelationEncoder(v_dim, q_dim, out_dim, dirs, L, nongt_dim=nongt, num_heads=heads, num_steps=steps,
                                              label_bias=label_bias)
        from tf_vqa_regat_b200.synthetic import explicit_param_values, make_explicit_inputs
        from tensorflow._core import Variable
        seed = 4000 + len(name)
        visual, question, adj, n_obj = make_explicit_inputs(v_dim, q_dim, L, B, N, seed)
        tv, tq = (Variable.make(np.asarray(x, dtype=np.float64)) for x in (visual, question))     # leaves the tape can differentiate
        ta = tf.constant(adj, dtype=tf.float64)
        enc(tv, ta, tq)                                   # creates the variables (Keras: on first call)
        nw = enc.named_weights()                          # (path, variable) in Keras-2 tracking order
        assert [id(w) for _, w in nw] == [id(w) for w in enc.trainable_variables]
        shapes = [tuple(w.shape) for _, w in nw]
        for (_, w), a in zip(nw, explicit_param_values(shapes, seed + 1)):
            w.assign(a)
        rng = np.random.default_rng(77)
        probe = rng.standard_normal((B, N, out_dim))
        with tf.GradientTape() as tape:
            out = enc(tv, ta, tq)
            loss = tf.reduce_sum(out * tf.constant(probe, dtype=tf.float64))
        grads = tape.gradient(loss, list(enc.trainable_variables) + [tv, tq])
        rec = dict(cfg=str(dict(v_dim=v_dim, q_dim=q_dim, out_dim=out_dim, dir_num=dirs, label_num=L, nongt_dim=nongt, num_heads=heads,
                                residual=residual, label_bias=label_bias, **({'num_steps': steps} if steps > 1 else {}))),
                   B=B, N=N, seed=seed, input_check=np.array([visual.sum(), question.sum(), adj.sum()]),
                   param_check=np.array([float(np.sum(w.numpy())) for _, w in nw]), output=out.numpy().astype(np.float32),
                   names=np.array([_norm_name(p) for p, _ in nw]), shapes=np.array([str(s_) for s_ in shapes]))
        for i, ((path, w), g) in enumerate(zip(nw, grads[:-2])):
            a = np.zeros(w.numpy().shape) if g is None else np.asarray(g.numpy(), dtype=np.float64)
            r = np.random.default_rng(100 + i).standard_normal(a.size)
            idx = np.random.default_rng(200 + i).choice(a.size, size=min(64, a.size), replace=False)
            n = _norm_name(path)
            rec["grad.norm/" + n] = np.sqrt((a * a).sum()); rec["grad.proj/" + n] = float(a.ravel() @ r)
            rec["grad.absmax/" + n] = np.abs(a).max() if a.size else 0.0
            rec["grad.idx/" + n] = idx.astype(np.int64); rec["grad.sample/" + n] = a.ravel()[idx]
        rec["grad_visual"] = grads[-2].numpy().astype(np.float32)
        rec["grad_question"] = grads[-1].numpy().astype(np.float32)
        np.savez_compressed(os.path.join(GOLD, f"refexec_explicit_{name}.npz"), **rec)
        print(name, "output", out.numpy().shape, "variables", len(nw), [_norm_name(p) for p, _ in nw][:6])

    # build_graph (position_emb.py:23-90): the reference's own NumPy function on seeded boxes, incl. padded (all-zero) boxes,
    # nested boxes (inside / cover), heavy overlaps (IoU >= 0.5) and far-apart pairs (no edge)
    rng = np.random.default_rng(9)
    rec = {}
    for k, n in enumerate((6, 17, 36)):
        x1 = rng.uniform(0, 500, n); y1 = rng.uniform(0, 380, n)
        w = rng.uniform(8, 300, n); h = rng.uniform(8, 220, n)
        bb = np.stack([x1, y1, np.minimum(x1 + w, 639), np.minimum(y1 + h, 479)], 1)
        bb[1] = [bb[0, 0] + 2, bb[0, 1] + 2, bb[0, 2] - 2, bb[0, 3] - 2]         # box 1 inside box 0
        bb[2] = bb[0] + [3, 3, 3, 3]                                              # box 2 overlaps box 0 heavily
        if n > 8:
            bb[-3:] = 0.0                                                         # padded objects
        spatial = np.concatenate([bb / [640, 480, 640, 480], ((bb[:, 2:3] - bb[:, 0:1] + 1) / 640), ((bb[:, 3:4] - bb[:, 1:2] + 1) / 480)], 1)
        rec[f"bbox{k}"], rec[f"spatial{k}"] = bb, spatial
        rec[f"adj{k}"] = ref_pe.build_graph(bb, spatial)
    np.savez_compressed(os.path.join(GOLD, "explicit_build_graph.npz"), **rec)
    print("build_graph", {k: v.shape for k, v in rec.items() if k.startswith("adj")}, "labels used", sorted(set(np.concatenate([rec[f"adj{k}"].ravel() for k in range(3)]).astype(int))))


if __name__ == "__main__":
    main()
