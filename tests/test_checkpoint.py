"""Checkpoint I/O by variable order (tf_vqa_regat_b200/checkpoint.py; reference main.py:145,155)."""
import os

import numpy as np
import pytest

from tf_vqa_regat_b200 import checkpoint as ck
from tf_vqa_regat_b200 import synthetic as syn
from tf_vqa_regat_b200.config import HotPathConfig, param_layout

SMALL = HotPathConfig(v_dim=192, q_dim=96, rel_dim=256, num_heads=4, nongt_dim=20, num_answers=301)


def test_round_trip_and_order(tmp_path):
    flat = syn.make_params(SMALL, seed=3, trained_like=True)
    p = tmp_path / "w.npz"
    ck.save_weights(str(p), SMALL, flat)
    back = ck.load_weights(str(p), SMALL)
    entries, total = param_layout(SMALL)
    assert back.shape == (total,)
    for e in entries:                                   # every variable bit-exact; alignment padding is not part of the contract
        np.testing.assert_array_equal(back[e.offset:e.offset + e.numel], flat[e.offset:e.offset + e.numel])
    arrays = ck.flat_to_arrays(SMALL, flat)
    # Keras-2 order: per WeightNorm wrapper [v, g, bias]; v2out first, classifier last; the label FC has no bias
    names = [e.name for e in entries]
    assert names[:3] == ["v_relation.v2out/v", "v_relation.v2out/g", "v_relation.v2out/bias"]
    assert names[-3:] == ["classifier.layers.3/v", "classifier.layers.3/g", "classifier.layers.3/bias"]
    i = names.index("v_relation.implicit_relation.bias/v")
    assert names[i + 1] == "v_relation.implicit_relation.bias/g" and names[i + 2].endswith("neighbor_net.0.pair_pos_fc/v")
    assert arrays[names.index("v_relation.implicit_relation.neighbor_net.1.linear_out_/v")].shape == (1, 1, 256, 256)   # Conv2D kernel
    assert arrays[1].shape == ()                        # scalar g (weight_norm.py:27-29)


def test_load_goes_by_order_not_by_name(tmp_path):
    flat = syn.make_params(SMALL, seed=4, trained_like=True)
    arrays = ck.flat_to_arrays(SMALL, flat)
    p = tmp_path / "anon.npz"
    np.savez(str(p), **{f"{i:03d}": a for i, a in enumerate(arrays)})     # no names at all
    back = ck.load_weights(str(p), SMALL)
    for e in param_layout(SMALL)[0]:
        np.testing.assert_array_equal(back[e.offset:e.offset + e.numel], flat[e.offset:e.offset + e.numel])


def test_mismatches_raise_like_keras(tmp_path):
    flat = syn.make_params(SMALL, seed=5, trained_like=False)
    arrays = ck.flat_to_arrays(SMALL, flat)
    with pytest.raises(ValueError, match="length"):
        ck.arrays_to_flat(SMALL, arrays[:-1])
    bad = list(arrays)
    bad[0] = bad[0].T.copy()
    with pytest.raises(ValueError, match="v_relation.v2out/v"):
        ck.arrays_to_flat(SMALL, bad)
    with pytest.raises(ValueError, match="elements"):
        ck.flat_to_arrays(SMALL, flat[:-1])
    # a different architecture (label_bias on) has one more variable: a checkpoint of the other one must be refused
    other = HotPathConfig(**{**{k: getattr(SMALL, k) for k in SMALL.__dataclass_fields__}, "label_bias": True})
    with pytest.raises(ValueError):
        ck.arrays_to_flat(other, arrays)


def test_variable_order_is_the_executed_reference_order():
    """tests/golden/reference_variable_order.json: `model.weights` of the reference's own RelationGraphAttentionNetwork, built and
    called over the TensorFlow stand-in (oracle/make_reference_api.py).  The flat layouts used on the device follow it."""
    import json
    import os
    from tf_vqa_regat_b200.question import question_layout
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_variable_order.json")))
    d = ref["dims"]
    for label_bias in (False, True):
        cfg = HotPathConfig(v_dim=d["v_dim"], q_dim=d["q_dim"], rel_dim=d["rel_dim"], num_heads=d["num_heads"], nongt_dim=d["nongt_dim"],
                            num_answers=d["num_answers"], dir_num=d["dir_num"], pos_emb_dim=d["pos_emb_dim"], label_bias=label_bias)
        mine = [(n, list(s)) for n, s, _ in question_layout(d["n_token"], d["emb_dim"], d["q_dim"], "c")[0]] + \
               [(e.name, list(e.shape)) for e in param_layout(cfg)[0]]
        theirs = [(norm, shape) for _, norm, shape, _ in ref["label_bias_%s" % str(label_bias).lower()]]
        assert mine == theirs
    frozen = [norm for _, norm, _, tr in ref["label_bias_false"] if not tr]
    assert frozen == ["w_emb.emb_/emb_"]                     # language_model.py:58: the second table starts frozen
    assert ("v_relation.implicit_relation.bias/bias" in [r[1] for r in ref["label_bias_true"]]
            and "v_relation.implicit_relation.bias/bias" not in [r[1] for r in ref["label_bias_false"]])


def test_whole_model_checkpoint_round_trip(tmp_path):
    from tf_vqa_regat_b200.question import question_layout
    from oracle import language_model as olm
    front = olm.make_params(60, 12, SMALL.q_dim, "c", seed=11)
    shapes = [(n, s) for n, s, _ in question_layout(60, 12, SMALL.q_dim, "c")[0]]
    flat = syn.make_params(SMALL, seed=5, trained_like=True)
    p = tmp_path / "model.npz"
    ck.save_model_weights(str(p), [front[n] for n, _ in shapes], SMALL, flat)
    fa, back = ck.load_model_weights(str(p), SMALL, shapes)
    for (n, _), a in zip(shapes, fa):
        np.testing.assert_array_equal(a, np.asarray(front[n], dtype=np.float32))
    for e in param_layout(SMALL)[0]:
        np.testing.assert_array_equal(back[e.offset:e.offset + e.numel], flat[e.offset:e.offset + e.numel])
    with pytest.raises(ValueError, match="q_emb.gru/kernel"):
        ck.load_model_weights(str(p), SMALL, [(n, (s if n != "q_emb.gru/kernel" else (5, 5))) for n, s in shapes])
    with pytest.raises(ValueError):
        ck.load_model_weights(str(p), SMALL, shapes[:-1])      # one front-end variable short: the hot-path count no longer fits


def test_foreign_npz_is_ordered_by_integer_index_and_meta_is_validated(tmp_path):
    """'arr_10' must not sort before 'arr_2'; a `__meta__` record written for another variable list is refused."""
    import json
    cfg = HotPathConfig(v_dim=96, q_dim=48, rel_dim=64, num_heads=4, nongt_dim=5, num_answers=37)
    flat = syn.make_params(cfg, seed=3, trained_like=True)
    arrays = ck.flat_to_arrays(cfg, flat)
    assert len(arrays) > 11
    p = tmp_path / "foreign.npz"
    np.savez(p, *arrays)                                   # numpy's own keys: arr_0 ... arr_N (lexical order would scramble them)
    np.testing.assert_array_equal(ck.load_weights(str(p), cfg), flat)
    np.savez(tmp_path / "gap.npz", **{"0": arrays[0], "2": arrays[1]})
    with pytest.raises(ValueError, match="without gaps"):
        ck.load_weights(str(tmp_path / "gap.npz"), cfg)
    np.savez(tmp_path / "noindex.npz", kernel=arrays[0])
    with pytest.raises(ValueError, match="no variable index"):
        ck.load_weights(str(tmp_path / "noindex.npz"), cfg)
    good = tmp_path / "good.npz"
    ck.save_weights(str(good), cfg, flat)
    with np.load(good) as z:
        payload = {k: z[k] for k in z.files}
    meta = json.loads(bytes(payload["__meta__"]).decode())
    meta["variables"][3] = "somebody.else/v"
    payload["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez(tmp_path / "othermodel.npz", **payload)
    with pytest.raises(ValueError, match="different variable list"):
        ck.load_weights(str(tmp_path / "othermodel.npz"), cfg)


def test_keras_h5_needs_h5py_and_says_so(tmp_path):
    try:
        import h5py  # noqa: F401
        pytest.skip("h5py is installed here")
    except ImportError:
        pass
    with pytest.raises(ImportError, match="convert_keras_h5"):
        ck.read_keras_h5(str(tmp_path / "x.h5"))
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "convert_keras_h5.py"),
                        str(tmp_path / "x.h5"), str(tmp_path / "out")], capture_output=True, text=True)
    assert r.returncode != 0 and "h5py" in (r.stderr + r.stdout)
