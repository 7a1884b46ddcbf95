"""Checkpoint I/O by variable order (tf_vqa_regat_b200/checkpoint.py; reference main.py:145,155)."""
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
