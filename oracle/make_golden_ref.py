"""Golden vectors for stages 2-3, the loss and the optimizer, produced by EXECUTING THE
REFERENCE'S OWN FILES.  Build container only (needs /root/reference):

    python -m oracle.make_golden_ref            # rewrites tests/golden/refexec_*.npz

What runs unmodified from /root/reference: model/{weight_norm,fc,graph_att_layer,graph_att_net,
relation_encoder,fusion,classifier,rel_graph_net,position_emb}.py and train.py -- i.e. every line
of the reference's own logic on this path: layer wiring and constructor-argument slips, the
row-slice + raw-reshape scramble, transposes, the q-mask, the GradientTape / clip_by_norm /
Adamax training loop (train.train) and the evaluation loop (train.evaluate).

What is substituted: the `tensorflow` package, which is not installed and not installable here.
oracle/tf_shim/tensorflow restates the ~40 TensorFlow/Keras primitives those files call (Dense,
grouped 1x1 Conv2D, matmul, softmax, l2_normalize, sigmoid_cross_entropy_with_logits,
clip_by_norm, experimental.Adamax, Keras-2 weight tracking ...) on torch-CPU; each one names the
published semantics it follows.  So these vectors pin the oracle (and through it the CUDA path)
to the reference's code, with TensorFlow's primitives as the remaining, documented assumption.

The question front-end (language_model.py) is outside the path: the three objects the model
calls for it are replaced by feeders that hand over the synthetic q_att / q_last tensors.

Runs in float64 (the reference's own `set_floatx('float64')` line, main.py:101) for vectors that
are exact to ~1e-15, plus one float32 forward per case to record the fp32 noise floor.
"""
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"

TINY = dict(v_dim=96, q_dim=48, rel_dim=64, num_heads=4, nongt_dim=5, num_answers=37)
SMALL = dict(v_dim=192, q_dim=96, rel_dim=256, num_heads=4, nongt_dim=20, num_answers=301)   # head dim 64: runs on the kernels
CASES = {
    # name: (cfg kwargs, B, N, adaptive, trained_like, train steps)
    "tiny_n9_m5": (TINY, 3, 9, True, True, 3),
    "tiny_n4_m5_init": (TINY, 2, 4, False, False, 2),
    "tiny_n100_m100_fullkk": (dict(TINY, nongt_dim=100), 2, 100, True, True, 2),      # K x K bias at the adaptive maximum
    "small_n36_m20": (SMALL, 3, 36, True, True, 3),
    "small_n36_m20_init": (SMALL, 3, 36, True, False, 2),
    "small_n12_clamped": (SMALL, 2, 12, False, True, 2),
    "small_n36_m36_fullkk": (dict(SMALL, nongt_dim=36), 2, 36, False, True, 2),
    "small_n100_m20_adaptive": (SMALL, 2, 100, True, True, 2),
    "small_nov2out": (dict(SMALL, v_dim=256), 2, 20, False, True, 2),
    "small_dir1_labelbias_nores": (dict(SMALL, dir_num=1, num_heads=8, rel_dim=512, residual=False, label_bias=True),
                                   2, 24, False, True, 2),
    "full_b4_n36_m20": (dict(), 4, 36, False, True, 2),   # BASELINE.json configs[0]: batch 4, K=36, full widths
}
LR = 1e-3
SAMPLE = 64            # elements kept per tensor for the large cases


def _import_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    import tensorflow as tf
    assert tf.__version__.endswith("standin")
    import train as ref_train
    from model import relation_encoder as ref_enc
    from model.classifier import SimpleClassifier
    from model.fusion import BUTD
    from model.rel_graph_net import RelationGraphAttentionNetwork
    return tf, ref_train, ref_enc, BUTD, SimpleClassifier, RelationGraphAttentionNetwork


class _QuestionFeeder:
    """Stands where w_emb / q_emb / q_att stand in RelationGraphAttentionNetwork.call
    (rel_graph_net.py:40-45,57): `question` is a (q_att, q_last) pair and passes through."""

    def __init__(self, role):
        self.role = role

    def __call__(self, x):
        if self.role == "q_att":               # q_att(q_emb(w_emb(question))) -> q_emb_self_att
            return x[0]
        return x                               # w_emb, q_emb: pass the pair along

    def call_last(self, x):                    # q_emb.call_last(w_emb) -> q_emb
        return x[1]


def _norm_name(n):
    """shim path -> layout name: drop the FullyConnected list level and the wrapped layer level."""
    parts = [p for p in n.split("/")[0].split(".")]
    out, i = [], 0
    while i < len(parts):
        if parts[i] == "layers" and i + 1 < len(parts) and out and out[-1] != "classifier":
            i += 2                              # FullyConnected.layers[k]
            continue
        if parts[i] == "layer":                 # WeightNorm.layer (the wrapped Dense / Conv2D)
            i += 1
            continue
        out.append(parts[i]); i += 1
    return ".".join(out) + "/" + n.split("/")[1]


def build_model(tf, mods, cfg):
    _, _, ref_enc, BUTD, SimpleClassifier, Net = mods
    v_relation = ref_enc.ImplicitRelationEncoder(                      # rel_graph_net.py:96-102
        cfg.v_dim, cfg.q_dim, cfg.rel_dim, cfg.dir_num, cfg.pos_emb_dim, cfg.nongt_dim,
        num_heads=cfg.num_heads, num_steps=1, residual_connection=cfg.residual, label_bias=cfg.label_bias)
    classifier = SimpleClassifier(cfg.q_dim, cfg.q_dim * 2, cfg.num_answers, 0.2)   # :104-105
    joint = BUTD(cfg.rel_dim, cfg.q_dim, cfg.q_dim)                                 # :108
    return Net(_QuestionFeeder("w_emb"), _QuestionFeeder("q_emb"), _QuestionFeeder("q_att"), v_relation, joint,
               classifier, "butd", "implicit")


def _summ(a, rng_seed, full):
    """Every element enters through sum / norm / a seeded random projection; `full` keeps the tensor itself."""
    a = np.asarray(a, dtype=np.float64)
    r = np.random.default_rng(rng_seed).standard_normal(a.size)
    idx = np.random.default_rng(rng_seed + 1).choice(a.size, size=min(SAMPLE, a.size), replace=False)
    d = dict(sum=a.sum(), norm=np.sqrt((a * a).sum()), absmax=np.abs(a).max(), proj=float(a.ravel() @ r),
             idx=idx.astype(np.int64), sample=a.ravel()[idx])
    if full:
        d["full"] = a
    return d


def run_case(name, mods, save=True):
    tf, ref_train, ref_enc = mods[0], mods[1], mods[2]
    from tf_vqa_regat_b200 import synthetic as syn
    from tf_vqa_regat_b200.config import HotPathConfig, param_layout
    kw, B, N, adaptive, tl, steps = CASES[name]
    cfg = HotPathConfig(**kw)
    entries, _ = param_layout(cfg)
    flat = syn.make_params(cfg, seed=7, trained_like=tl)
    named = syn.unflatten(cfg, flat)
    batches = [syn.make_inputs(cfg, B, N, seed=1000 + s, adaptive=adaptive) for s in range(steps + 1)]  # last = eval batch
    full = cfg.rel_dim <= 64                                   # tiny cases keep whole tensors
    out = dict(cfg=np.array(repr(kw)), B=B, N=N, adaptive=adaptive, trained_like=tl, steps=steps, lr=LR,
               input_check=np.array([float(np.sum(b["features"], dtype=np.float64)) for b in batches]))

    # record what concat_visual_question produced (the q-mask), without touching its arithmetic
    seen = {}
    orig_cvq = ref_enc.concat_visual_question

    def recording_cvq(q, v, mask=True):
        r = orig_cvq(q, v, mask=mask)
        seen["v_cat_q"] = r
        return r
    ref_enc.concat_visual_question = recording_cvq

    def fresh_model(floatx):
        tf.keras.backend.set_floatx(floatx)
        tf.random.set_seed(0)
        model = build_model(tf, mods, cfg)
        b0 = batches[0]
        pos, _, _ = ref_train.prepare_graph_variables("implicit", b0["boxes"], None, None, N, cfg.nongt_dim,
                                                      cfg.pos_emb_dim, 11, 15)
        model(b0["features"], None, (tf.constant(b0["q_att"]), tf.constant(b0["q_last"])), pos, None, None)  # builds
        nw = model.named_weights()
        assert [_norm_name(n) for n, _ in nw] == [e.name for e in entries], \
            [(a, e.name) for (a, _), e in zip(nw, entries) if _norm_name(a) != e.name][:4]
        assert [tuple(w.shape) for _, w in nw] == [tuple(e.shape) for e in entries]
        assert [id(w) for _, w in nw] == [id(w) for w in model.trainable_variables]
        model.set_weights([named[e.name] for e in entries])
        return model

    # ---- forward + one GradientTape in float64, straight through the reference's model and loss
    model = fresh_model("float64")
    b0 = batches[0]
    pos, _, _ = ref_train.prepare_graph_variables("implicit", b0["boxes"], None, None, N, cfg.nongt_dim,
                                                  cfg.pos_emb_dim, 11, 15)
    q_att = tf.Variable.make(b0["q_att"].astype(np.float64)); q_last = tf.Variable.make(b0["q_last"].astype(np.float64))
    target = tf.convert_to_tensor(b0["target"])
    with tf.GradientTape() as tape:
        v_emb = model.v_relation(b0["features"], pos, q_att)                       # rel_graph_net.py:53
        joint, weights = model.joint_emb(v_emb, q_last)                            # :58
        logits = model.classifier(joint)                                           # :62
        loss = ref_train.instance_bce_with_logits(logits, target)                  # train.py:107
        loss_avg = tf.reduce_mean(loss) * tf.cast(tf.shape(target)[1], tf.float32)  # train.py:108
    grads = tape.gradient(loss_avg, list(model.trainable_variables) + [q_att, q_last])
    D = cfg.rel_dim
    out.update(logits=logits.numpy(), loss=loss_avg.numpy(), joint=joint.numpy(), att_weights=weights.numpy(),
               mask=(np.abs(seen["v_cat_q"].numpy()[:, :, D:]).sum(-1) != 0).astype(np.float64),
               dq_att=grads[-2].numpy(), dq_last=grads[-1].numpy())
    v1 = v_emb.numpy()
    out.update(v1=v1) if cfg.rel_dim <= 256 else out.update(v1_head=v1[:, :, :16], v1_sum=v1.sum(-1))
    for i, (e, g) in enumerate(zip(entries, grads[:-2])):
        for k, v in _summ(g.numpy(), 100 + i, full).items():
            out[f"grad.{k}/{e.name}"] = v
    # the same forward through the full model object (feeders in place) must agree exactly
    again = model(b0["features"], None, (tf.constant(b0["q_att"]), tf.constant(b0["q_last"])), pos, None, None)
    assert np.array_equal(again.numpy(), out["logits"])

    # ---- float32 forward: the reference's own arithmetic at its working precision
    m32 = fresh_model("float32")
    l32 = m32(b0["features"], None, (tf.constant(b0["q_att"]), tf.constant(b0["q_last"])), pos, None, None)
    assert l32.dtype == __import__("torch").float32
    out["logits_f32"] = l32.numpy()

    # ---- the reference's train() and evaluate() loops, float64
    model = fresh_model("float64")
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

        def __init__(self, bs):
            self.bs = bs
            self.data_loader_len = len(bs)
            self.num_total_data = B * len(bs)

        def generator(self):                                     # dataset.py:357-361 yields trim_collate's tuple
            for b in self.bs:
                yield (b["features"], None, (tf.constant(b["q_att"]), tf.constant(b["q_last"])), b["boxes"],
                       np.zeros((B, 1)), np.zeros((B, 1)), tf.convert_to_tensor(b["target"]))

    with tempfile.TemporaryDirectory() as tmp:
        args = types.SimpleNamespace(base_lr=LR, epochs=1, lr_decay_step=2, lr_decay_rate=0.25, grad_clip=cfg.grad_clip,
                                     output=tmp + "/", relation_type="implicit", nongt_dim=cfg.nongt_dim,
                                     imp_pos_emb_dim=cfg.pos_emb_dim, spa_label_num=11, sem_label_num=15, print_freq=500)
        stdout, sys.stdout = sys.stdout, open(os.devnull, "w")
        try:
            ref_train.train(model, Loader(batches[:steps]), Loader(batches[steps:]), args)
            log = open(os.path.join(tmp, "log.txt")).read()
        finally:
            sys.stdout.close(); sys.stdout = stdout
    ref_enc.concat_visual_question = orig_cvq
    assert len(step_logits) == steps + 1 and model.optimizer.iterations == steps
    assert np.array_equal(step_logits[0], out["logits"])
    out["train.logits"] = np.stack(step_logits[:steps]); out["eval.logits"] = step_logits[steps]
    out["eval.score_pct"] = float(log.split("eval_score: ")[1].split()[0])
    for i, (e, w) in enumerate(zip(entries, model.trainable_variables)):
        for k, v in _summ(w.numpy(), 500 + i, full).items():
            out[f"param.{k}/{e.name}"] = v
        m, u = model.optimizer.slots(w)
        out[f"adamax_m.norm/{e.name}"] = float(np.sqrt((m.numpy() ** 2).sum()))
        out[f"adamax_u.norm/{e.name}"] = float(np.sqrt((u.numpy() ** 2).sum()))
    if save:
        np.savez_compressed(os.path.join(GOLD, f"refexec_{name}.npz"), **out)
    print(f"refexec {name}: loss {float(out['loss']):.6f}  fp32-vs-fp64 logits "
          f"{np.abs(out['logits_f32'] - out['logits']).max() / np.abs(out['logits']).max():.2e}  "
          f"eval score {out['eval.score_pct']:.3f}%")
    return out


if __name__ == "__main__":
    assert os.path.isdir(REF), "needs /root/reference (build container only)"
    mods = _import_reference()
    os.makedirs(GOLD, exist_ok=True)
    for name in (sys.argv[1:] or CASES):
        run_case(name, mods)
