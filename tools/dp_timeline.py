"""Per-range timeline of one data-parallel step (events on the compute and communication streams).
   torchrun --nproc-per-node N tools/dp_timeline.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from tf_vqa_regat_b200 import synthetic as syn
from tf_vqa_regat_b200.config import HotPathConfig
from tf_vqa_regat_b200.engine import HotPathEngine
from tf_vqa_regat_b200.dp import DataParallelTrainer, GraphedDPStep

rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
cfg = HotPathConfig(); B, N = 256, 36
eng = HotPathEngine(cfg, B, N, dtype="bf16", device=f"cuda:{lr_}")
eng.load_params(syn.make_params(cfg, seed=7, trained_like=True))
inp = syn.make_inputs(cfg, B, N, seed=1000 + rank)
d = {k: torch.tensor(v).cuda() for k, v in inp.items() if k != "n_obj"}
tr = DataParallelTrainer(eng)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for i in range(3):
        tr.step(d["features"], d["boxes"], d["q_att"], d["q_last"], d["target"], 1e-3)
    torch.cuda.synchronize()
    g = GraphedDPStep(tr, d["features"], d["boxes"], d["q_att"], d["q_last"], d["target"], st)
    step = 3
    for i in range(5):
        g.replay(); step += 1; eng.update(1e-3, step)
    torch.cuda.synchronize(); dist.barrier()
    g.trace = []
    g.replay(); step += 1; eng.update(1e-3, step)
    end = torch.cuda.Event(enable_timing=True); end.record(st)
    torch.cuda.synchronize()
if rank == 0:
    t0 = g.trace[0][1]
    for label, ev in g.trace:
        print(f"{t0.elapsed_time(ev) * 1e3:9.1f} us  {label}")
    print(f"{t0.elapsed_time(end) * 1e3:9.1f} us  update end")
dist.barrier(); dist.destroy_process_group()
