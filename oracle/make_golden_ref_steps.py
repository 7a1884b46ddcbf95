"""Golden vectors for ImplicitRelationEncoder with num_steps > 1 (the constructor's `num_steps` argument,
relation_encoder.py:82-91 -- the q-mask, the graph attention network and the residual add are repeated on the running `visual`
with the SAME variables), produced by EXECUTING the reference's own model/relation_encoder.py, graph_att_net.py,
graph_att_layer.py, fc.py, weight_norm.py and position_emb.py over oracle/tf_shim.  Build container only:

    python -m oracle.make_golden_ref_steps        # rewrites tests/golden/encsteps_*.npz

Nothing of the reference is substituted besides the `tensorflow` package (see oracle/make_golden_ref.py).  Recorded per case: the
inputs' checksums (inputs and variables come from the seeded generators in tf_vqa_regat_b200/synthetic.py), the encoder output, and
the gradients of sum(output * probe) w.r.t. every variable, the visual features and the question (GradientTape over the stand-in).
TEST INFRASTRUCTURE."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

SMALL = dict(v_dim=192, q_dim=96, rel_dim=256, num_heads=4, nongt_dim=20, num_answers=301)   # head dim 64: runs on the kernels
CASES = {
    # name: (cfg kwargs, B, N, adaptive, num_steps)
    "small_n36_m20_steps2": (SMALL, 3, 36, True, 2),
    "small_n36_m20_steps3": (SMALL, 2, 36, True, 3),
    "small_n12_clamped_nores_steps2": (dict(SMALL, residual=False), 2, 12, False, 2),      # no residual: visual = imp_rel (:90-91)
    # v_dim == out_dim: no v2out (:52-55); one direction (the kernels want dir_num * num_heads to be a multiple of 8)
    "small_nov2out_dir1_steps2": (dict(SMALL, v_dim=512, rel_dim=512, num_heads=8, dir_num=1), 2, 24, False, 2),
}


def main():
    from oracle.make_golden_ref import _import_reference, _norm_name
    mods = _import_reference()
    tf, ref_train, ref_enc = mods[0], mods[1], mods[2]
    from tensorflow._core import Variable
    from tf_vqa_regat_b200 import synthetic as syn
    from tf_vqa_regat_b200.config import HotPathConfig
    tf.keras.backend.set_floatx("float64")
    for name, (kw, B, N, adaptive, steps) in CASES.items():
        cfg = HotPathConfig(**kw)
        named = syn.unflatten(cfg, syn.make_params(cfg, seed=7, trained_like=True).astype(np.float64))
        seed = 5000 + len(name)
        inp = syn.make_inputs(cfg, B, N, seed=seed, adaptive=adaptive)
        enc = ref_enc.ImplicitRelationEncoder(cfg.v_dim, cfg.q_dim, cfg.rel_dim, cfg.dir_num, cfg.pos_emb_dim, cfg.nongt_dim,
                                              num_heads=cfg.num_heads, num_steps=steps, residual_connection=cfg.residual,
                                              label_bias=cfg.label_bias)
        pos, _, _ = ref_train.prepare_graph_variables("implicit", inp["boxes"], None, None, N, cfg.nongt_dim, cfg.pos_emb_dim, 11, 15)
        tv = Variable.make(inp["features"].astype(np.float64))
        tq = Variable.make(inp["q_att"].astype(np.float64))
        enc(tv, pos, tq)                                  # creates the variables (Keras: on first call)
        nw = enc.named_weights()
        assert [id(w) for _, w in nw] == [id(w) for w in enc.trainable_variables]
        names = ["v_relation." + _norm_name(p) for p, _ in nw]
        assert all(n in named for n in names), [n for n in names if n not in named]
        for (_, w), n in zip(nw, names):
            assert tuple(w.shape) == tuple(named[n].shape), (n, w.shape, named[n].shape)
            w.assign(named[n])
        probe = np.random.default_rng(78).standard_normal((B, N, cfg.rel_dim))
        with tf.GradientTape() as tape:
            out = enc(tv, pos, tq)
            loss = tf.reduce_sum(out * tf.constant(probe, dtype=tf.float64))
        grads = tape.gradient(loss, list(enc.trainable_variables) + [tv, tq])
        rec = dict(cfg=str(kw), B=B, N=N, adaptive=adaptive, seed=seed, num_steps=steps,
                   input_check=np.array([float(np.sum(inp[k], dtype=np.float64)) for k in ("features", "boxes", "q_att")]),
                   output=out.numpy(), names=np.array(names))
        for i, (n, g) in enumerate(zip(names, grads[:-2])):
            a = np.zeros(named[n].shape) if g is None else np.asarray(g.numpy(), dtype=np.float64)
            r = np.random.default_rng(100 + i).standard_normal(a.size)
            idx = np.random.default_rng(200 + i).choice(a.size, size=min(64, a.size), replace=False)
            rec["grad.norm/" + n] = np.sqrt((a * a).sum()); rec["grad.proj/" + n] = float(a.ravel() @ r)
            rec["grad.idx/" + n] = idx.astype(np.int64); rec["grad.sample/" + n] = a.ravel()[idx]
        rec["grad_visual"] = grads[-2].numpy().astype(np.float32)
        rec["grad_question"] = grads[-1].numpy()
        # the same encoder at num_steps = 1 on the same inputs: the fixture shows the extra steps change the result
        enc.num_steps = 1
        rec["output_steps1"] = enc(tv, pos, tq).numpy().astype(np.float32)
        np.savez_compressed(os.path.join(GOLD, f"encsteps_{name}.npz"), **rec)
        print(name, "output", rec["output"].shape, "variables", len(nw), "max |steps_k - steps_1|",
              float(np.abs(rec["output"] - rec["output_steps1"]).max()))


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
