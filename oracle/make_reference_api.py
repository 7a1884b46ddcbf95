"""Records the public interface of the reference's hot-path classes -- constructor, call and helper signatures with
their defaults, read with `inspect` from the reference's own files (imported over oracle/tf_shim) -- into
tests/golden/reference_api.json.  Build container only (needs /root/reference):

    python -m oracle.make_reference_api

tests/test_api_conformance.py holds tf_vqa_regat_b200.model to it.  TEST INFRASTRUCTURE."""
import inspect
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SURFACE = {                       # module -> names on the path
    "fc": ["FullyConnected"], "weight_norm": ["WeightNorm"], "graph_att_layer": ["GraphSelfAttentionLayer"],
    "graph_att_net": ["GraphAttentionNetwork"], "relation_encoder": ["ImplicitRelationEncoder", "concat_visual_question"],
    "fusion": ["BUTD"], "classifier": ["SimpleClassifier"], "position_emb": ["prepare_graph_variables"],
}
METHODS = {"BUTD": ["attention_weights"]}


def params(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        out.append({"name": p.name, "kind": p.kind.name,
                    "default": None if p.default is inspect.Parameter.empty else repr(p.default),
                    "has_default": p.default is not inspect.Parameter.empty})
    return out


def describe(obj, name):
    if inspect.isclass(obj):
        d = {"kind": "class", "__init__": params(obj.__init__), "call": params(obj.call)}
        for m in METHODS.get(name, []):
            d[m] = params(getattr(obj, m))
        return d
    return {"kind": "function", "signature": params(obj)}


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
    sys.path.insert(0, "/root/reference")
    import importlib
    api = {}
    for mod, names in SURFACE.items():
        m = importlib.import_module("model." + mod)
        for n in names:
            api[f"{mod}.{n}"] = describe(getattr(m, n), n)
    path = os.path.join(ROOT, "tests", "golden", "reference_api.json")
    with open(path, "w") as f:
        json.dump(api, f, indent=1, sort_keys=True)
    print("wrote", path, len(api), "entries")
    variable_order()


def variable_order():
    """`model.weights` of the reference's RelationGraphAttentionNetwork (built and called once, as rel_graph_net.py:113-123
    does) under the stand-in's Keras-2 tracking rules: [keras-style path, normalised name, shape, trainable] per variable, in
    the order `save_weights` / `load_weights` use (main.py:145,155): HDF5 checkpoints store, per top-level layer
    (w_emb, q_emb, q_att, v_relation, joint_emb, classifier -- rel_graph_net.py:16-21), that layer's `weights`.  (Keras lists a
    layer's non-trainable variables after its trainable ones; the only frozen variable, w_emb.emb_, already comes last within
    its layer, so the per-layer order is the same either way.)  Two builds: label_bias off (the shipped config) and on."""
    sys.path.insert(0, ROOT)
    import numpy as np
    import tensorflow as tf
    from model import relation_encoder as ref_enc
    from model.classifier import SimpleClassifier
    from model.fusion import BUTD
    from model.language_model import QuestionEmbedding, QuestionSelfAttention, WordEmbedding
    from model.position_emb import prepare_graph_variables
    from model.rel_graph_net import RelationGraphAttentionNetwork
    from oracle.make_golden_ref import _norm_name
    dims = dict(n_token=60, emb_dim=12, v_dim=192, q_dim=96, rel_dim=256, num_heads=4, nongt_dim=20, num_answers=301, dir_num=2,
                pos_emb_dim=64)
    out = {"dims": dims}
    for label_bias in (False, True):
        d = dims
        model = RelationGraphAttentionNetwork(
            WordEmbedding(d["n_token"], d["emb_dim"], 0.2, "c"), QuestionEmbedding(2 * d["emb_dim"], d["q_dim"], 1, False, 0.2),
            QuestionSelfAttention(d["q_dim"], 0.2),
            ref_enc.ImplicitRelationEncoder(d["v_dim"], d["q_dim"], d["rel_dim"], d["dir_num"], d["pos_emb_dim"], d["nongt_dim"],
                                            num_heads=d["num_heads"], num_steps=1, residual_connection=True, label_bias=label_bias),
            BUTD(d["rel_dim"], d["q_dim"], d["q_dim"]), SimpleClassifier(d["q_dim"], 2 * d["q_dim"], d["num_answers"], 0.2),
            "butd", "implicit")
        rng = np.random.default_rng(0)
        bb = np.sort(rng.uniform(0, 400, (2, 24, 4)).astype(np.float32), -1)
        pos, _, _ = prepare_graph_variables("implicit", bb, None, None, 24, d["nongt_dim"], 64, 11, 15)
        model(np.abs(rng.standard_normal((2, 24, d["v_dim"]))).astype(np.float32), None,
              tf.constant(rng.integers(0, d["n_token"], (2, 14))), pos, None, None)
        tv = model.trainable_variables
        out["label_bias_%s" % str(label_bias).lower()] = [
            [n, _norm_name(n), list(w.shape), any(w is t for t in tv)] for n, w in model.named_weights()]
    path = os.path.join(ROOT, "tests", "golden", "reference_variable_order.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, len(out["label_bias_false"]), "variables")


if __name__ == "__main__":
    main()
