"""Data-parallel training of the hot path: one process per GPU, graphs sharded by image, weights replicated,
ONE exchange per step -- an all-reduce (sum) of the flat gradient buffer, carried by our own kernels over
symmetric NVLink / NVSwitch-multicast memory (csrc/dp_exchange.cu; REGAT_DP_COMM=nccl selects NCCL instead).  Every graph is independent
end to end, so rank r simply takes graphs [r*B/R, (r+1)*B/R); with equal shards the mean over the global batch is
the average of rank means, hence grad_scale = 1/R and a SUM all-reduce reproduce the single-GPU gradient.

What is reduced is dL/dW_eff (+ bias gradients), which is linear in the batch; the weight-norm projection,
per-tensor clip and Adamax run after the reduce on every rank (train.py:112-113).

Two ways to drive the exchange:
  fused   (default with symmetric memory, fp32 wire): the ENGINE launches the in-place exchange of each gradient range on its
          own stream as soon as the backward pass has finished the range, followed by that range's clip + Adamax
          (regat_engine_set_dp + regat_engine_train_step[_dev]); the whole step is one launch sequence without host
          scalars, so it replays from ONE CUDA graph (engine.GraphedTrainStep).
  callback: regat_engine_fwd_bwd announces each finished range to Python, which starts the exchange (own kernels with a
          host-side epoch, the staged bf16 wire format, or NCCL) on a communication stream; the optimizer follows eagerly."""
import os

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int):
    """Contiguous, equal shards.  The global batch must divide evenly (the loss mean needs equal shards)."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_range(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


def allreduce_flat_(flat: torch.Tensor, group=None, bucket_elems: int = 0):
    """In-place SUM all-reduce of a flat buffer, optionally in buckets (elements) so that buckets can overlap
    with compute issued on other streams."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return flat
    if bucket_elems <= 0 or bucket_elems >= flat.numel():
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return flat
    for lo in range(0, flat.numel(), bucket_elems):
        dist.all_reduce(flat[lo:lo + bucket_elems], op=dist.ReduceOp.SUM, group=group)
    return flat


class DataParallelTrainer:
    """engine: anything with fwd_bwd(..., grad_scale=) -> dict(loss, score), .grads (flat tensor), .update(lr, step) and
    optionally set_grad_callback(fn) (HotPathEngine): with it, each range of the flat gradient buffer is all-reduced on a side
    stream as soon as the backward pass has finished writing it, overlapping the remaining backward kernels."""

    SMALL_FP32 = 1 << 20     # ranges below this many elements are latency-bound: reduce them in fp32, without the two cast launches

    def __init__(self, engine, group=None, bucket_elems: int = 0, overlap: bool = True, comm_dtype: str = "auto"):
        """comm_dtype: "fp32", "bf16" or "auto" (bf16 when the engine computes in bf16: the gradients were produced from bf16
        activations anyway, and halving the exchanged bytes is what keeps the all-reduce hidden behind the backward pass)."""
        self.engine, self.group, self.bucket_elems = engine, group, bucket_elems
        if comm_dtype == "auto":
            comm_dtype = "bf16" if getattr(engine, "dtype_name", "fp32") == "bf16" else "fp32"
        self.comm_dtype = comm_dtype
        self._g16 = None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.step_count = 0
        self.overlap = bool(overlap and self.world > 1 and hasattr(engine, "set_grad_callback") and engine.grads.is_cuda)
        self._reduced = 0
        self._cb_error = None
        self.fused = False
        # Exchange backend: "symm" = our kernels over symmetric (peer-mapped / NVSwitch-multicast) memory, "nccl" = NCCL all-reduce
        self.backend = "nccl"
        want = os.environ.get("REGAT_DP_COMM", "symm")
        if want == "symm" and self.world > 1 and engine.grads.is_cuda:
            try:
                self._setup_symm(group)
                self.backend = "symm"
            except Exception as ex:                      # no symmetric memory on this system: NCCL carries the exchange
                if self.rank == 0:
                    print(f"[regat dp] symmetric-memory exchange unavailable ({type(ex).__name__}: {ex}); using NCCL", flush=True)
        # fused: the engine itself exchanges + optimizes each range (one CUDA graph per step); needs the in-place fp32 wire format
        if (self.backend == "symm" and self.wire == "f32" and hasattr(engine, "set_dp")
                and os.environ.get("REGAT_DP_FUSED", "1") != "0"):
            engine.set_dp(self._stage_ptrs, self._mc, self._flag_ptrs_dev, self.rank, self.world, self._blocks)
            self.fused = True
        if self.overlap:
            self.comm_stream = torch.cuda.Stream(engine.grads.device)
            engine.set_grad_callback(self._on_ready)

    def _on_ready(self, offset, numel):
        # runs inside a ctypes callback: an exception here would be swallowed, so it is kept and re-raised after fwd_bwd
        try:
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(main)
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                self._allreduce_range(offset, numel)
            self._reduced += numel
        except BaseException as ex:             # noqa: BLE001
            if self._cb_error is None:
                self._cb_error = ex

    def _setup_symm(self, group):
        """Symmetric staging + flag buffers (torch.distributed._symmetric_memory does the allocation and the address exchange;
        the data path is ours: csrc/dp_exchange.cu)."""
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        g = self.engine.grads
        gname = (group or dist.group.WORLD).group_name
        self.wire = os.environ.get("REGAT_DP_WIRE", "f32" if hasattr(self.engine, "rebind_grads") else "bf16")
        if self.comm_dtype == "fp32":
            self.wire = "f32"
        if self.wire == "f32":
            # the gradient buffer itself becomes the symmetric allocation: reduced in place, no staging, no casts
            g32 = symm_mem.empty(g.numel(), dtype=torch.float32, device=g.device)
            g32.zero_()
            h = symm_mem.rendezvous(g32, gname)
            self.engine.rebind_grads(g32)
            g = g32
        else:
            self._g16 = symm_mem.empty(g.numel(), dtype=torch.bfloat16, device=g.device)
            h = symm_mem.rendezvous(self._g16, gname)
        self._flags = symm_mem.empty(128, dtype=torch.int32, device=g.device)    # [0,64): host-epoch calls, [64,128): the engine's
        self._flags.zero_()
        hf = symm_mem.rendezvous(self._flags, gname)
        torch.cuda.synchronize(g.device)
        dist.barrier(group)                            # every rank's flags are zero before anyone signals
        self._stage_ptrs = (C.c_uint64 * self.world)(*[int(p) for p in h.buffer_ptrs])
        self._flag_ptrs = (C.c_uint64 * self.world)(*[int(p) for p in hf.buffer_ptrs])
        self._flag_ptrs_dev = (C.c_uint64 * self.world)(*[int(p) + 256 for p in hf.buffer_ptrs])
        self._mc = int(getattr(h, "multicast_ptr", 0) or 0)         # 0 where the fabric has no multicast: peers one by one
        if os.environ.get("REGAT_DP_MULTICAST", "1") == "0":
            self._mc = 0
        self._epoch = 0
        self._blocks = int(os.environ.get("REGAT_DP_BLOCKS", "32"))
        self._symm_handles = (h, hf)

    def _exchange_symm(self, offset, numel):
        from . import _lib
        g = self.engine.grads
        st = torch.cuda.current_stream().cuda_stream
        l = _lib.lib()
        self._epoch += 1
        if self.wire == "f32":
            _lib.check(l.regat_dp_allreduce_f32(self._stage_ptrs, self._mc, self._flag_ptrs, self.rank, self.world, offset, numel,
                                                self._epoch, self._blocks, st))
            return
        _lib.check(l.regat_cast(_lib.F32, _lib.BF16, g.data_ptr() + 4 * offset, self._g16.data_ptr() + 2 * offset, numel, st))
        _lib.check(l.regat_dp_reduce_bcast(self._stage_ptrs, self._mc, self._flag_ptrs, self.rank, self.world, offset, numel,
                                           self._epoch, self._blocks, st))
        _lib.check(l.regat_dp_wait_unpack(self._g16.data_ptr(), g.data_ptr() + 4 * offset, self._flag_ptrs, self.rank, self.world,
                                          offset, numel, self._epoch, st))

    def _allreduce_range(self, offset, numel):
        g = self.engine.grads
        if self.backend == "symm" and os.environ.get("REGAT_DP_DEBUG", "") == "":
            return self._exchange_symm(offset, numel)
        dbg = os.environ.get("REGAT_DP_DEBUG", "")      # timing experiments only: "skip" = no exchange at all, "nocomm" = casts only
        if dbg == "skip":
            return
        if self.comm_dtype == "bf16" and g.is_cuda and numel >= self.SMALL_FP32:
            from . import _lib
            if self._g16 is None:
                self._g16 = torch.empty(g.numel(), dtype=torch.bfloat16, device=g.device)
            st = torch.cuda.current_stream().cuda_stream
            l = _lib.lib()
            _lib.check(l.regat_cast(_lib.F32, _lib.BF16, g.data_ptr() + 4 * offset, self._g16.data_ptr() + 2 * offset, numel, st))
            if dbg != "nocomm":
                dist.all_reduce(self._g16[offset:offset + numel], op=dist.ReduceOp.SUM, group=self.group)
            _lib.check(l.regat_cast(_lib.BF16, _lib.F32, self._g16.data_ptr() + 2 * offset, g.data_ptr() + 4 * offset, numel, st))
        else:
            dist.all_reduce(g[offset:offset + numel], op=dist.ReduceOp.SUM, group=self.group)

    def broadcast_params(self, src: int = 0):
        if self.world > 1:
            dist.broadcast(self.engine.params, src=src, group=self.group)
            if hasattr(self.engine, "params_changed"):
                self.engine.params_changed()

    def fwd_bwd_allreduce(self, features, boxes, q_att, q_last, target):
        """Forward + backward on this rank's shard with the gradient all-reduce (overlapped when possible); the optimizer has
        NOT run.  Callback-driven path (see the module docstring)."""
        self._reduced = 0
        self._cb_error = None
        out = self.engine.fwd_bwd(features, boxes, q_att, q_last, target, grad_scale=1.0 / self.world)
        if self.overlap:
            torch.cuda.current_stream().wait_stream(self.comm_stream)       # always joined, whatever happened in the callbacks
            if self._cb_error is not None:
                raise self._cb_error
            if self._reduced != self.engine.grads.numel():
                raise RuntimeError(f"gradient-ready ranges covered {self._reduced} of {self.engine.grads.numel()} elements: "
                                   "refusing to guess which ranges were exchanged")
        elif self.world > 1 and (self.backend == "symm" or (self.comm_dtype == "bf16" and self.engine.grads.is_cuda)):
            self._allreduce_range(0, self.engine.grads.numel())
        else:
            allreduce_flat_(self.engine.grads, self.group, self.bucket_elems)
        return out

    def step(self, features, boxes, q_att, q_last, target, lr):
        """One optimizer step on this rank's shard (already sliced)."""
        self.step_count += 1
        if self.fused:
            loss = self.engine.train_step(features, boxes, q_att, q_last, target, lr, self.step_count)
            return {"loss": loss[0], "score": loss[1]}
        out = self.fwd_bwd_allreduce(features, boxes, q_att, q_last, target)
        self.engine.update(lr, self.step_count)
        return out

    def capture_step(self, features, boxes, q_att, q_last, target, stream=None):
        """ONE CUDA graph of the whole data-parallel step on these input tensors (fused mode only): forward, backward, the
        in-place exchange of every gradient range, clip + Adamax.  lr / step live on the device (engine.set_lr / set_step)."""
        if not self.fused:
            raise RuntimeError("capture_step needs the fused exchange (symmetric memory, fp32 wire)")
        from .engine import GraphedTrainStep
        self.engine.set_grad_callback(None)           # the warm-up fwd_bwd inside must not start a Python-driven exchange
        try:
            return GraphedTrainStep(self.engine, features, boxes, q_att, q_last, target, stream)
        finally:
            if self.overlap:
                self.engine.set_grad_callback(self._on_ready)


def _timed_event(stream):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record(stream)
    return ev


class GraphedDPStep:
    """CUDA-graph replay of the data-parallel forward+backward for ONE fixed set of input tensors, without putting NCCL inside
    a graph: the engine's gradient-ready callbacks split the capture into compute-only segments; on replay the bucketed
    all-reduces are issued eagerly on the comm stream between the segments (so they still overlap the following segment)."""

    def __init__(self, trainer: DataParallelTrainer, features, boxes, q_att, q_last, target, stream):
        self.tr, self.stream = trainer, stream
        eng = trainer.engine
        self.graphs, self.ranges = [torch.cuda.CUDAGraph()], []
        pool = torch.cuda.graph_pool_handle()

        def on_ready(offset, numel):               # called from inside fwd_bwd while it is being captured
            self.graphs[-1].capture_end()
            self.ranges.append((offset, numel))
            self.graphs.append(torch.cuda.CUDAGraph())
            self.graphs[-1].capture_begin(pool=pool)

        eng.set_grad_callback(on_ready)
        try:
            with torch.cuda.stream(stream):
                self.graphs[0].capture_begin(pool=pool)
                eng.fwd_bwd(features, boxes, q_att, q_last, target, grad_scale=1.0 / trainer.world)
                self.graphs[-1].capture_end()
        finally:
            eng.set_grad_callback(trainer._on_ready if trainer.overlap else None)
        assert sum(n for _, n in self.ranges) == eng.grads.numel(), "gradient-ready ranges do not cover the flat buffer"
        self.trace = None

    def replay(self):
        """Enqueue on self.stream (must be the current stream): segments + overlapped all-reduces; joined at the end."""
        tr = self.tr
        trace = self.trace          # None, or a list that receives (label, event) pairs with timing enabled (tools/dp_timeline.py)
        mark = (lambda label, stream: trace.append((label, _timed_event(stream)))) if trace is not None else (lambda *a: None)
        mark("step begin", self.stream)
        for k, g in enumerate(self.graphs):
            g.replay()
            if k < len(self.ranges):
                off, n = self.ranges[k]
                ev = torch.cuda.Event()
                ev.record(self.stream)
                mark(f"range {k} ready ({n} el)", self.stream)
                tr.comm_stream.wait_event(ev)
                with torch.cuda.stream(tr.comm_stream):
                    mark(f"range {k} comm begin", tr.comm_stream)
                    tr._allreduce_range(off, n)
                    mark(f"range {k} comm end", tr.comm_stream)
        mark("compute end", self.stream)
        self.stream.wait_stream(tr.comm_stream)
        mark("step end", self.stream)
