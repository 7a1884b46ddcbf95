"""ImplicitRelationEncoder with num_steps > 1 (relation_encoder.py:82-91).  CPU: the oracle restatement against
vectors produced by executing the reference's own encoder (oracle/make_golden_ref_steps.py).  GPU: the layer mirror's
ImplicitRelationEncoder(num_steps=k) -- the kernels called once per propagation step on the running `visual` -- against the same
vectors.  (The fused train-step engine is built for the shipped num_steps = 1; DESIGN.md section 8, limits.)"""
import ast
import glob
import os

import numpy as np
import pytest
import torch

from tf_vqa_regat_b200 import synthetic as syn
from tf_vqa_regat_b200.config import HotPathConfig

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "encsteps_*.npz")))
IDS = [os.path.basename(f)[len("encsteps_"):-4] for f in FILES]


def _case(path):
    g = np.load(path)
    cfg = HotPathConfig(**ast.literal_eval(str(g["cfg"])))
    B, N, steps = int(g["B"]), int(g["N"]), int(g["num_steps"])
    inp = syn.make_inputs(cfg, B, N, seed=int(g["seed"]), adaptive=bool(g["adaptive"]))
    np.testing.assert_allclose([float(np.sum(inp[k], dtype=np.float64)) for k in ("features", "boxes", "q_att")], g["input_check"],
                               rtol=1e-12)
    flat = syn.make_params(cfg, seed=7, trained_like=True)
    return g, cfg, B, N, steps, inp, flat


def test_fixtures_exist():
    assert len(FILES) >= 4
    for f in FILES:
        g = np.load(f)
        assert int(g["num_steps"]) >= 2
        # the extra propagation steps are visible in the result: a mirror that ignored num_steps would not pass below
        assert np.abs(g["output"] - g["output_steps1"]).max() > 0.1 * np.abs(g["output"]).max()


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_oracle_matches_reference_execution(path):
    from oracle import position_emb as pe
    from oracle import regat_torch as ot
    g, cfg, B, N, steps, inp, flat = _case(path)
    p = ot.to_torch_params(syn.unflatten(cfg, flat.astype(np.float64)))
    pos = pe.prepare_graph_variables("implicit", inp["boxes"], None, None, N, cfg.nongt_dim, cfg.pos_emb_dim, 11, 15)[0]
    tv = torch.tensor(inp["features"].astype(np.float64), requires_grad=True)
    tq = torch.tensor(inp["q_att"].astype(np.float64), requires_grad=True)
    out = ot.encoder(p, cfg, tv, torch.tensor(np.asarray(pos), dtype=torch.float64), tq, num_steps=steps)
    ref = g["output"]
    assert np.abs(out.detach().numpy() - ref).max() < 1e-10 * np.abs(ref).max()
    one = ot.encoder(p, cfg, tv, torch.tensor(np.asarray(pos), dtype=torch.float64), tq).detach().numpy()
    assert np.abs(one - g["output_steps1"]).max() < 1e-5 * np.abs(one).max()                       # stored as float32
    probe = torch.tensor(np.random.default_rng(78).standard_normal((B, N, cfg.rel_dim)))
    names = [str(n) for n in g["names"]]
    grads = torch.autograd.grad((out * probe).sum(), [p[n] for n in names] + [tv, tq], allow_unused=True)
    for n, gr in zip(names, grads[:-2]):
        a = np.zeros(p[n].shape) if gr is None else gr.numpy()
        want = float(g["grad.norm/" + n])
        assert abs(np.sqrt((a * a).sum()) - want) <= 1e-9 * max(want, 1e-12) + 1e-12, n
        np.testing.assert_allclose(a.ravel()[g["grad.idx/" + n]], g["grad.sample/" + n], rtol=1e-8, atol=1e-10, err_msg=n)
    assert np.abs(grads[-2].numpy() - g["grad_visual"]).max() < 1e-5 * np.abs(g["grad_visual"]).max() + 1e-9
    assert np.abs(grads[-1].numpy() - g["grad_question"]).max() < 1e-9 * np.abs(g["grad_question"]).max() + 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("lazy", [True, False], ids=["box_geometry", "pos_emb_tensor"])
@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_layer_mirror_matches_reference_execution(path, lazy):
    from tf_vqa_regat_b200.model import build_hot_path, prepare_graph_variables
    g, cfg, B, N, steps, inp, flat = _case(path)
    model = build_hot_path(cfg)
    model.load_flat(cfg, flat)
    enc = model.v_relation
    dev = {k: torch.tensor(v).cuda() for k, v in inp.items() if k != "n_obj"}
    pos_emb, _, _ = prepare_graph_variables("implicit", dev["boxes"], None, None, N, cfg.nongt_dim, cfg.pos_emb_dim, 11, 15, lazy=lazy)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    assert enc.num_steps == 1
    assert rel(enc(dev["features"], pos_emb, dev["q_att"]).cpu().numpy(), g["output_steps1"]) < 1e-4
    enc.num_steps = steps
    got = enc(dev["features"], pos_emb, dev["q_att"]).cpu().numpy()
    assert rel(got, g["output"]) < 2e-4, rel(got, g["output"])
