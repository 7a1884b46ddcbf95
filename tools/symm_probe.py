"""Probe torch symmetric memory on this box: rendezvous, multicast support, library all-reduce timing vs NCCL."""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
n = 8747328
t = symm_mem.empty(n, dtype=torch.bfloat16, device=f"cuda:{lr_}")
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
if rank == 0:
    print("handle attrs:", [a for a in dir(hdl) if not a.startswith("_")])
    print("multicast_ptr:", hex(hdl.multicast_ptr), "buffer_ptrs:", [hex(p) for p in hdl.buffer_ptrs][:2], "signal_pad_ptrs:", [hex(p) for p in hdl.signal_pad_ptrs][:2])
    print("signal pad size", symm_mem.get_signal_pad_size())
ops = [o for o in dir(torch.ops.symm_mem)]
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
x = torch.ones(n, dtype=torch.bfloat16, device="cuda")
gn = dist.group.WORLD.group_name
res = {"nccl": timeit(lambda: dist.all_reduce(x))}
for name in ("two_shot_all_reduce_", "multimem_all_reduce_", "one_shot_all_reduce"):
    try:
        op = getattr(torch.ops.symm_mem, name)
        t.fill_(1)
        res[name] = timeit(lambda: op(t, "sum", gn))
    except Exception as ex:
        res[name] = f"failed: {str(ex)[:80]}"
small = symm_mem.empty(787648, dtype=torch.bfloat16, device=f"cuda:{lr_}"); symm_mem.rendezvous(small, gn)
xs = torch.ones(787648, dtype=torch.bfloat16, device="cuda")
res["nccl_small"] = timeit(lambda: dist.all_reduce(xs))
for name in ("two_shot_all_reduce_", "multimem_all_reduce_", "one_shot_all_reduce"):
    try:
        op = getattr(torch.ops.symm_mem, name)
        res[name + "_small"] = timeit(lambda: op(small, "sum", gn))
    except Exception as ex:
        res[name + "_small"] = f"failed: {str(ex)[:80]}"
if rank == 0:
    for k, v in res.items(): print(k, v)
dist.barrier(); dist.destroy_process_group()
