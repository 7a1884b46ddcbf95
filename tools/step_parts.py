"""Where the step's time goes beyond its kernels: graph-replayed timing of (a) forward only, (b) forward + backward without the
optimizer, (c) the whole train step, at the bench shape.  GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tf_vqa_regat_b200 import synthetic
from tf_vqa_regat_b200.config import HotPathConfig
from tf_vqa_regat_b200.engine import HotPathEngine

B, N = 256, 36
cfg = HotPathConfig()
eng = HotPathEngine(cfg, B, N, "bf16")
eng.load_params(synthetic.make_params(cfg, seed=7, trained_like=True))
inp = synthetic.make_inputs(cfg, B, N, seed=3)
dev = {k: torch.as_tensor(v).cuda() for k, v in inp.items()}
args = (dev["features"], dev["boxes"], dev["q_att"], dev["q_last"], dev["target"])
# -1: the step's stream outranks the engine's side stream (the engine's optimizer / exchange streams take the caller's priority)
ST = torch.cuda.Stream(priority=int(os.environ.get("PRIO", "0")))
ST.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(ST):
    eng.set_lr(1e-3)
    for _ in range(2):
        eng.train_step_dev(*args)
torch.cuda.synchronize()


def timed(fn, reps=30):
    st = ST
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            fn()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print("forward (eval)      %8.1f us" % timed(lambda: eng.forward(*args[:4])))
print("forward + backward  %8.1f us" % timed(lambda: eng.fwd_bwd(*args)))
print("whole train step    %8.1f us" % timed(lambda: eng.train_step_dev(*args)))
