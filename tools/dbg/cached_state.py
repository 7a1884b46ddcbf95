import sys, numpy as np, torch
sys.path.insert(0, '.')
from tf_vqa_regat_b200 import synthetic as syn
from tf_vqa_regat_b200.config import HotPathConfig
from tf_vqa_regat_b200.engine import HotPathEngine
SMALL = dict(v_dim=192, q_dim=96, rel_dim=256, num_heads=4, nongt_dim=20, num_answers=301)
cfg = HotPathConfig(**SMALL)
B, N = 3, 36
inp = syn.make_inputs(cfg, B, N, seed=1000, adaptive=False)
flat = syn.make_params(cfg, seed=7, trained_like=True)
dev = {k: torch.tensor(v).cuda() for k, v in inp.items() if k != "n_obj"}
for trial in range(3):
    eng = HotPathEngine(cfg, B, N, dtype="bf16"); eng.load_params(flat)
    eng.forward(dev["features"], dev["boxes"], dev["q_att"], dev["q_last"])
    eng.fwd_bwd(dev["features"], dev["boxes"], dev["q_att"], dev["q_last"], dev["target"])
    eng.update(1e-2, 1)
    after = eng.forward(dev["features"], dev["boxes"], dev["q_att"], dev["q_last"])
    fresh = HotPathEngine(cfg, B, N, dtype="bf16"); fresh.load_params(eng.params.clone())
    want = fresh.forward(dev["features"], dev["boxes"], dev["q_att"], dev["q_last"])
    torch.cuda.synchronize()
    M, D, dirs = 20, cfg.rel_dim, cfg.dir_num
    for name, shape, dt in (("v0", (B, N, D), None), ("s", (B, N, D), None), ("Qb", (B, N, dirs * D), None), ("KVb", (B, M, 2 * dirs * D), None),
                            ("v1", (B, N, D), None), ("alpha", (16,), torch.float32), ("scal", (16,), torch.float32), ("pooled", (B, D), None), ("logits", (B, 304), torch.float32)):
        a, b = eng.buffer(name, shape, dt).float(), fresh.buffer(name, shape, dt).float()
        d = (a - b).abs()
        print(trial, name, "max diff", float(d.max()), "rel", float(d.max() / b.abs().max()), "n differing", int((d > 0).sum()), "of", d.numel(),
              "per-graph" if d.dim() == 3 else "", [float(d[i].max()) for i in range(B)] if d.dim() >= 2 and d.shape[0] == B else "")
    print(trial, "logits rel", float((after - want).abs().max() / want.abs().max()))
