"""Regenerates tests/golden/*.npz.  Run in the BUILD container only (needs
/root/reference, which does not exist on the GPU box):

    python -m oracle.make_golden

stage1_*.npz  : inputs + outputs of the REFERENCE's own model/position_emb.py
                (imported from /root/reference, pure NumPy) -- these pin the stage-1 oracle.
hotpath_*.npz : outputs of the fp64 NumPy oracle on seeded synthetic inputs/weights
                (tf_vqa_regat_b200.synthetic) -- regression anchors only; the reference
                cannot run stages 2-3 here (no TensorFlow), so they pin nothing.
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def load_reference_stage1():
    spec = importlib.util.spec_from_file_location("ref_position_emb", "/root/reference/model/position_emb.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def stage1():
    ref = load_reference_stage1()
    rng = np.random.default_rng(20261018)
    cases = {"n36_m20": (2, 36, 20), "n36_m36": (1, 36, 36), "n12_m20_clamped": (2, 12, 20),
             "n100_m20_padded": (1, 100, 20)}
    for name, (B, N, nongt) in cases.items():
        x1 = rng.uniform(0, 560, (B, N)); y1 = rng.uniform(0, 400, (B, N))
        w = rng.uniform(8, 320, (B, N)); h = rng.uniform(8, 240, (B, N))
        bb = np.stack([x1, y1, np.minimum(x1 + w, 639), np.minimum(y1 + h, 479)], -1).astype(np.float32)
        if "padded" in name:
            bb[:, 61:] = 0.0                        # zero post-padding, dataset.py:334,346
        bb[0, 1] = bb[0, 0]                         # identical boxes: hits the 1e-3 clamp and log(1)
        pos_mat = ref.tf_extract_position_matrix(bb, nongt_dim=nongt)
        pos_emb, a, b = ref.prepare_graph_variables("implicit", bb, None, None, N, nongt, 64, 11, 15)
        assert a is None and b is None
        np.savez_compressed(os.path.join(GOLD, f"stage1_{name}.npz"), bb=bb, nongt=np.int32(nongt),
                            pos_mat=pos_mat, pos_emb=pos_emb.astype(np.float32))
        print("stage1", name, pos_emb.shape, pos_emb.dtype)


def hotpath():
    from oracle import regat_numpy as onp
    from tf_vqa_regat_b200 import synthetic as syn
    from tf_vqa_regat_b200.config import HotPathConfig
    cases = {
        # name: (cfg kwargs, B, N, adaptive, trained_like)
        "tiny_n9_m5": (dict(v_dim=96, q_dim=48, rel_dim=64, num_heads=4, nongt_dim=5, num_answers=37), 3, 9, True, True),
        "tiny_n4_m5": (dict(v_dim=96, q_dim=48, rel_dim=64, num_heads=4, nongt_dim=5, num_answers=37), 2, 4, False, False),
        "full_b2_n36_m20": (dict(), 2, 36, False, True),
    }
    for name, (kw, B, N, adaptive, tl) in cases.items():
        cfg = HotPathConfig(**kw)
        inp = syn.make_inputs(cfg, B, N, seed=1000, adaptive=adaptive)
        named = syn.unflatten(cfg, syn.make_params(cfg, seed=7, trained_like=tl).astype(np.float64))
        f64 = lambda a: a.astype(np.float64)
        out = onp.forward(named, cfg, f64(inp["features"]), inp["boxes"], f64(inp["q_att"]), f64(inp["q_last"]),
                          f64(inp["target"]))
        np.savez_compressed(os.path.join(GOLD, f"hotpath_{name}.npz"),
                            cfg=np.array(repr(kw)), B=B, N=N, adaptive=adaptive, trained_like=tl,
                            logits=out["logits"], loss=out["loss"], joint=out["joint"],
                            att_weights=out["att_weights"], mask=out["mask"],
                            v1_head=out["v1"][:, :, :16], v1_sum=out["v1"].sum(-1))
        print("hotpath", name, float(out["loss"]))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    stage1()
    hotpath()
