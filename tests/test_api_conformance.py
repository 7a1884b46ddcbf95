"""The host-side mirror keeps the reference's layer interface: module names, class names, constructor and call
signatures with their defaults, as recorded from the reference's own files in tests/golden/reference_api.json
(oracle/make_reference_api.py).  A mirror may ADD trailing keyword arguments with defaults (documented extensions); it may
not rename, reorder, drop or re-default anything the reference has.  Also the error behaviour that needs no GPU."""
import importlib
import json
import os

import pytest

from oracle.make_reference_api import describe

HERE = os.path.dirname(os.path.abspath(__file__))
API = json.load(open(os.path.join(HERE, "golden", "reference_api.json")))


def _conforms(ref, mine, where):
    assert len(mine) >= len(ref), (where, "parameters dropped")
    for r, m in zip(ref, mine):
        assert (m["name"], m["kind"], m["has_default"], m["default"]) == (r["name"], r["kind"], r["has_default"], r["default"]), \
            (where, r, m)
    for extra in mine[len(ref):]:
        assert extra["has_default"] or extra["kind"] in ("VAR_KEYWORD", "VAR_POSITIONAL"), (where, "extension without default", extra)


@pytest.mark.parametrize("key", sorted(API))
def test_signatures_match_the_reference(key):
    mod, name = key.split(".")
    m = importlib.import_module("tf_vqa_regat_b200.model." + mod)        # same module names as the reference's model/ package
    mine = describe(getattr(m, name), name)
    ref = API[key]
    assert mine["kind"] == ref["kind"]
    for part in ref:
        if part != "kind":
            _conforms(ref[part], mine[part], f"{key}.{part}")


def test_extensions_are_the_documented_ones():
    """Everything beyond the reference's signatures, listed: a lazy box-geometry handle instead of the materialised embedding
    (SURVEY 8b 'extension needed'), and output-placement / fusion hints used between mirror layers."""
    extras = {}
    for key, ref in API.items():
        mod, name = key.split(".")
        mine = describe(getattr(importlib.import_module("tf_vqa_regat_b200.model." + mod), name), name)
        for part in ref:
            if part != "kind" and len(mine[part]) > len(ref[part]):
                extras[f"{key}.{part}"] = [p["name"] for p in mine[part][len(ref[part]):]]
    assert extras == {
        "position_emb.prepare_graph_variables.signature": ["lazy"],
        "weight_norm.WeightNorm.call": ["out", "out_ld", "relu"],
        "graph_att_net.GraphAttentionNetwork.call": ["residual"],
    }, extras


def test_error_behaviour_without_a_gpu():
    import torch
    from tf_vqa_regat_b200 import model as M
    from tf_vqa_regat_b200._lib import RegatError
    net = M.GraphAttentionNetwork(2, 1, 96, 64, nongt_dim=5, num_heads=1, pos_emb_dim=64)
    with pytest.raises(ValueError):                                      # graph_att_net.py:42-46
        net(torch.zeros(1, 4, 96), None, None)
    net2 = M.GraphAttentionNetwork(2, 1, 96, 64, nongt_dim=5, num_heads=1, pos_emb_dim=-1)
    with pytest.raises(ValueError):                                      # graph_att_net.py:47-51
        net2(torch.zeros(1, 4, 96), None, torch.zeros(1, 4, 4, 64))
    with pytest.raises(AssertionError):                                  # graph_att_net.py:18
        M.GraphAttentionNetwork(3, 1, 96, 64)
    with pytest.raises(ValueError):                                      # weight_norm.py:12-13
        M.WeightNorm(M.FullyConnected([4, 4]))
    with pytest.raises(RegatError):                                      # host tensors are refused: there is no CPU path
        net(torch.zeros(1, 4, 96), None, torch.zeros(1, 4, 4, 64))
