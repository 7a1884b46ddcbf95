"""Data parallel on real GPUs (needs >= 2 devices: run with `gpurun --gpus 2`): R ranks on shards of the batch with the overlapped
bucketed NCCL all-reduce reach the same parameters as one GPU on the whole batch."""
import os

import numpy as np
import pytest
import torch

from tf_vqa_regat_b200 import synthetic as syn
from tf_vqa_regat_b200.config import HotPathConfig

pytestmark = pytest.mark.gpu
SMALL = dict(v_dim=192, q_dim=96, rel_dim=256, num_heads=4, nongt_dim=20, num_answers=301)


def _worker(rank, world, port, cfg_kw, B, N, steps, lr, out):
    import torch.distributed as dist
    from tf_vqa_regat_b200.dp import DataParallelTrainer, shard_batch
    from tf_vqa_regat_b200.engine import HotPathEngine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = HotPathConfig(**cfg_kw)
    inp = syn.make_inputs(cfg, B, N, seed=1000, adaptive=True)
    eng = HotPathEngine(cfg, B // world, N, dtype="fp32", device=f"cuda:{rank}")
    eng.load_params(syn.make_params(cfg, seed=7 + rank, trained_like=True))     # ranks start DIFFERENT: broadcast must fix it
    tr = DataParallelTrainer(eng, overlap=True)
    tr.broadcast_params(0)
    shard = shard_batch({k: v for k, v in inp.items() if k != "n_obj"}, rank, world)
    dev = {k: torch.tensor(v).cuda() for k, v in shard.items()}
    losses = []
    for _ in range(steps):
        o = tr.step(dev["features"], dev["boxes"], dev["q_att"], dev["q_last"], dev["target"], lr)
        losses.append(float(o["loss"]))
    torch.cuda.synchronize()
    out[rank] = (eng.params.cpu().numpy(), losses, tr.overlap)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpus_equal_one_gpu():
    import torch.multiprocessing as mp
    from tf_vqa_regat_b200.engine import HotPathEngine
    B, N, steps, lr = 8, 36, 2, 1e-3
    cfg = HotPathConfig(**SMALL)
    inp = syn.make_inputs(cfg, B, N, seed=1000, adaptive=True)
    ref = HotPathEngine(cfg, B, N, dtype="fp32", device="cuda:0")
    ref.load_params(syn.make_params(cfg, seed=7, trained_like=True))
    dev = {k: torch.tensor(v).cuda() for k, v in inp.items() if k != "n_obj"}
    ref_losses = [float(ref.train_step(dev["features"], dev["boxes"], dev["q_att"], dev["q_last"], dev["target"], lr, s + 1)[0])
                  for s in range(steps)]
    p_ref = ref.params.cpu().numpy()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29533, SMALL, B, N, steps, lr, out), nprocs=2, join=True)
    assert out[0][2], "the overlapped bucketed path did not run"
    assert np.array_equal(out[0][0], out[1][0]), \
        f"replicas diverged: max |diff| {np.abs(out[0][0] - out[1][0]).max()}"      # post-all-reduce math is deterministic
    # rank losses are shard means: their average is the global-batch loss
    np.testing.assert_allclose(np.mean([out[0][1], out[1][1]], axis=0), ref_losses, rtol=1e-5)
    # Adamax moves every element by <= lr per step: agreement far inside that (summation order differs across shards).  The
    # softmax-shift "zero directions" only carry rounding noise, which Adamax normalises to +-lr: excluded.
    from tf_vqa_regat_b200.config import param_layout
    for e in param_layout(cfg)[0]:
        if "implicit_relation.bias/" in e.name or e.name.endswith(".key/bias") or e.name in ("joint_emb.linear/bias", "joint_emb.v2attention/bias"):
            continue
        d = np.abs(out[0][0][e.offset:e.offset + e.numel] - p_ref[e.offset:e.offset + e.numel]).max()
        assert d < 0.05 * steps * lr + 1e-6, (e.name, d)


def test_exchange_kernels_single_rank_identity():
    """csrc/dp_exchange.cu with world = 1 (no peer, no multicast address): the flag protocol must complete with the rank
    signalling itself, the fp32 in-place reduction must leave the range unchanged, the bf16 staging path must return the
    bf16 rounding of the range, and ranges outside [offset, offset+numel) must not be touched.  Several epochs in a row."""
    import ctypes as C
    from tf_vqa_regat_b200 import _lib
    l = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    n, off, cnt = 1 << 16, 1024, 40000
    g = torch.randn(n, device="cuda")
    ref = g.clone()
    flags = torch.zeros(64, dtype=torch.int32, device="cuda")
    gp = (C.c_uint64 * 1)(g.data_ptr())
    fp = (C.c_uint64 * 1)(flags.data_ptr())
    for epoch in (1, 2, 3):
        _lib.check(l.regat_dp_allreduce_f32(gp, 0, fp, 0, 1, off, cnt, epoch, 4, st))
    torch.cuda.synchronize()
    assert torch.equal(g, ref)
    assert int(flags[0]) == 3 and int(flags[16]) == 3          # [kind 0][rank 0], [kind 1][rank 0]
    stage = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
    sp = (C.c_uint64 * 1)(stage.data_ptr())
    for epoch in (4, 5):
        _lib.check(l.regat_cast(_lib.F32, _lib.BF16, g.data_ptr() + 4 * off, stage.data_ptr() + 2 * off, cnt, st))
        _lib.check(l.regat_dp_reduce_bcast(sp, 0, fp, 0, 1, off, cnt, epoch, 4, st))
        _lib.check(l.regat_dp_wait_unpack(stage.data_ptr(), g.data_ptr() + 4 * off, fp, 0, 1, off, cnt, epoch, st))
    torch.cuda.synchronize()
    want = ref.clone()
    want[off:off + cnt] = ref[off:off + cnt].bfloat16().float()
    assert torch.equal(g, want)
    # argument errors surface as status codes, not as launches
    assert l.regat_dp_allreduce_f32(gp, 0, fp, 0, 1, 2, cnt, 6, 4, st) == -5          # offset not a multiple of 4
    assert l.regat_dp_allreduce_f32(gp, 0, fp, 1, 1, off, cnt, 6, 4, st) == -1          # rank >= world
