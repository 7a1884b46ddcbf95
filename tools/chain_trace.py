"""Stage timeline of the chained launch (gemm_chain_kernel) inside one eager train step at the bench shape.  GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tf_vqa_regat_b200 import _lib, synthetic
from tf_vqa_regat_b200.config import HotPathConfig
from tf_vqa_regat_b200.engine import HotPathEngine

B, N = 256, 36
cfg = HotPathConfig()
eng = HotPathEngine(cfg, B, N, "bf16")
eng.load_params(synthetic.make_params(cfg, seed=7, trained_like=True))
inp = synthetic.make_inputs(cfg, B, N, seed=3)
dev = {k: torch.as_tensor(v).cuda() for k, v in inp.items()}
args = (dev["features"], dev["boxes"], dev["q_att"], dev["q_last"], dev["target"])
eng.set_lr(1e-3)
for _ in range(3):
    eng.train_step_dev(*args)
torch.cuda.synchronize()
buf = torch.zeros(256, device="cuda", dtype=torch.int64)
l = _lib.lib()
l.regat_gemm_trace(buf.data_ptr())
eng.train_step_dev(*args)
torch.cuda.synchronize()
l.regat_gemm_trace(None)
t = buf.cpu().tolist()
ghz = 1.965
us = lambda x: (x - t[128]) / ghz / 1e3
names = ["pv", "hid", "logits", "loss", "dhid", "djoint", "dpooled"]
print("stage      wait_passed   stage_done (us from kernel start, CTA 0)")
for i, n in enumerate(names):
    print(f"{n:8s} {us(t[152 + i]) if t[152 + i] else 0.0:12.2f} {us(t[136 + i]):12.2f}")
