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


if __name__ == "__main__":
    main()
