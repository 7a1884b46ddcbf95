"""Python face of the C engine (include/regat.h: regat_engine_*): the hot path of
rel_graph_net.py:53-62 and train.py:103-113 as single calls.  torch is used for device memory and streams
only; every number is produced by libregat.so kernels."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .config import HotPathConfig, param_layout

_DT = {"fp32": _lib.F32, "float32": _lib.F32, "bf16": _lib.BF16, "bfloat16": _lib.BF16}


def _c_config(cfg: HotPathConfig) -> _lib.Config:
    return _lib.Config(cfg.v_dim, cfg.q_dim, cfg.rel_dim, cfg.num_heads, cfg.pos_emb_dim, cfg.nongt_dim, cfg.dir_num,
                       cfg.num_answers, int(cfg.label_bias), int(cfg.residual), cfg.grad_clip, cfg.beta1, cfg.beta2, cfg.eps)


def _stream():
    return torch.cuda.current_stream().cuda_stream


class HotPathEngine:
    """Owns the flat parameter / gradient / Adamax buffers and the workspace of one GPU.

    dtype "fp32": exact-fp32 kernels (parity mode, 1e-4);  "bf16": bf16 activations and tcgen05 GEMMs with fp32
    accumulation, fp32 master weights (1e-2 on logits, same argmax).
    """

    def __init__(self, cfg: HotPathConfig, max_batch: int, max_rois: int, dtype: str = "bf16", device="cuda:0",
                 training: bool = True):
        if not torch.cuda.is_available():
            raise _lib.RegatError(-6, "HotPathEngine needs a CUDA device; there is no CPU fallback")
        self.cfg, self.dtype_name, self.dtype = cfg, dtype, _DT[dtype]
        self.device = torch.device(device)
        self.max_batch, self.max_rois = max_batch, max_rois
        self.lib = _lib.lib()
        torch.cuda.set_device(self.device)
        self._h = C.c_void_p()
        cc = _c_config(cfg)
        _lib.check(self.lib.regat_engine_create(C.byref(cc), self.dtype, max_batch, max_rois, C.byref(self._h)))
        pe, wb = C.c_int64(), C.c_int64()
        _lib.check(self.lib.regat_engine_sizes(self._h, C.byref(pe), C.byref(wb)))
        self.entries, total = param_layout(cfg)
        assert total == pe.value, "Python and C parameter layouts disagree"
        self.param_elems, self.workspace_bytes = pe.value, wb.value
        z = lambda: torch.zeros(pe.value, dtype=torch.float32, device=self.device)
        self.params = z()
        self.grads = z() if training else None
        self.adamax_m = z() if training else None
        self.adamax_u = z() if training else None
        self.workspace = torch.zeros(wb.value, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.regat_engine_bind(self._h, self.params.data_ptr(), _lib.ptr(self.grads), _lib.ptr(self.adamax_m),
                                              _lib.ptr(self.adamax_u), self.workspace.data_ptr(), wb.value))
        self._wd = _lib.wave_divisors(cfg.pos_emb_dim)
        _lib.check(self.lib.regat_engine_set_wave_div(self._h, self._wd.ctypes.data))
        self._loss = torch.zeros(2, dtype=torch.float32, device=self.device)
        self.step_count = 0

    def rebind_grads(self, grads: torch.Tensor):
        """Use `grads` (fp32, param_elems long, 256-byte aligned) as the flat gradient buffer from now on -- the data-parallel
        layer passes a symmetric (peer-mapped) allocation so that gradients are reduced in place over NVLink."""
        assert grads.dtype == torch.float32 and grads.numel() == self.param_elems and grads.is_cuda
        self.grads = grads
        _lib.check(self.lib.regat_engine_bind(self._h, self.params.data_ptr(), _lib.ptr(self.grads), _lib.ptr(self.adamax_m),
                                              _lib.ptr(self.adamax_u), self.workspace.data_ptr(), self.workspace_bytes))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self.lib.regat_engine_destroy(h)
            self._h = None

    # ---- parameters
    def load_params(self, flat):
        """flat fp32 buffer in config.param_layout order (np.ndarray or tensor)."""
        t = torch.as_tensor(np.asarray(flat, dtype=np.float32)) if not isinstance(flat, torch.Tensor) else flat
        assert t.numel() == self.param_elems
        self.params.copy_(t.to(self.device, torch.float32))
        self.params_changed()

    def params_changed(self):
        """Call after writing `self.params` in place: alpha = g/||v||, the bf16 kernels and the gathered biases are re-derived
        right away on the current stream, so that eager calls AND replays of already captured graphs see the new weights."""
        _lib.check(self.lib.regat_engine_params_changed(self._h))
        _lib.check(self.lib.regat_engine_refresh_weights(self._h, _stream()))

    def save_weights(self, path: str):
        """model.save_weights (main.py:145): the hot path's variables in Keras variable order (checkpoint.py)."""
        from . import checkpoint
        checkpoint.save_weights(path, self.cfg, self.params.detach().cpu().numpy())

    def load_weights(self, path: str):
        """model.load_weights (main.py:155), by order."""
        from . import checkpoint
        self.load_params(checkpoint.load_weights(path, self.cfg))

    def named(self, buf=None):
        """name -> view into a flat buffer (default: params)."""
        buf = self.params if buf is None else buf
        return {e.name: buf[e.offset:e.offset + e.numel].view(e.shape) for e in self.entries}

    # ---- calls
    @staticmethod
    def _chk(t, shape):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == tuple(shape), \
            f"expected contiguous fp32 CUDA tensor of shape {tuple(shape)}, got {tuple(t.shape)} {t.dtype} {t.device}"

    def _inputs(self, features, boxes, q_att, q_last):
        B, N, V = features.shape
        self._chk(features, (B, N, self.cfg.v_dim)); self._chk(boxes, (B, N, 4))
        self._chk(q_att, (B, self.cfg.q_dim)); self._chk(q_last, (B, self.cfg.q_dim))
        return B, N

    def forward(self, features, boxes, q_att, q_last, return_att=False):
        """Eval forward (train.py:136-177): logits [B, A] fp32 (+ BUTD attention weights [B,N,1])."""
        B, N = self._inputs(features, boxes, q_att, q_last)
        logits = torch.empty(B, self.cfg.num_answers, dtype=torch.float32, device=self.device)
        att = torch.empty(B, N, dtype=torch.float32, device=self.device) if return_att else None
        _lib.check(self.lib.regat_engine_forward(self._h, B, N, features.data_ptr(), boxes.data_ptr(), q_att.data_ptr(),
                                                 q_last.data_ptr(), logits.data_ptr(), _lib.ptr(att), _stream()))
        return (logits, att.unsqueeze(-1)) if return_att else logits

    def fwd_bwd(self, features, boxes, q_att, q_last, target, grad_scale=1.0, want_logits=False, want_dq=False):
        """GradientTape step (train.py:103-111).  Leaves dL/dW_eff + dL/db in self.grads; returns a dict with the
        device scalars 'loss' and 'score' (views of one 2-float tensor) and optional logits / dq_att / dq_last."""
        B, N = self._inputs(features, boxes, q_att, q_last)
        self._chk(target, (B, self.cfg.num_answers))
        out = {}
        logits = torch.empty(B, self.cfg.num_answers, dtype=torch.float32, device=self.device) if want_logits else None
        dqa = torch.empty(B, self.cfg.q_dim, dtype=torch.float32, device=self.device) if want_dq else None
        dql = torch.empty(B, self.cfg.q_dim, dtype=torch.float32, device=self.device) if want_dq else None
        _lib.check(self.lib.regat_engine_fwd_bwd(self._h, B, N, features.data_ptr(), boxes.data_ptr(), q_att.data_ptr(),
                                                 q_last.data_ptr(), target.data_ptr(), float(grad_scale), self._loss.data_ptr(),
                                                 _lib.ptr(logits), _lib.ptr(dqa), _lib.ptr(dql), _stream()))
        out.update(loss=self._loss[0], score=self._loss[1], logits=logits, dq_att=dqa, dq_last=dql)
        return out

    def finalize_grads(self):
        """In place: dL/dW_eff -> the reference's tape.gradient values (dv, dg)."""
        _lib.check(self.lib.regat_engine_finalize_grads(self._h, _stream()))

    def update(self, lr, step=None):
        """Per-tensor clip_by_norm + Adamax (train.py:112-113)."""
        self.step_count = step if step is not None else self.step_count + 1
        _lib.check(self.lib.regat_engine_update(self._h, float(lr), int(self.step_count), _stream()))

    def train_step(self, features, boxes, q_att, q_last, target, lr, step=None):
        B, N = self._inputs(features, boxes, q_att, q_last)
        self._chk(target, (B, self.cfg.num_answers))
        self.step_count = step if step is not None else self.step_count + 1
        _lib.check(self.lib.regat_engine_train_step(self._h, B, N, features.data_ptr(), boxes.data_ptr(), q_att.data_ptr(),
                                                    q_last.data_ptr(), target.data_ptr(), float(lr), int(self.step_count),
                                                    self._loss.data_ptr(), _stream()))
        return self._loss

    # ---- device-resident optimizer state: whole train steps without host scalars (one CUDA graph per step)
    def set_lr(self, lr):
        _lib.check(self.lib.regat_engine_set_lr(self._h, float(lr), _stream()))

    def set_step(self, steps_done):
        """Number of optimizer steps already taken (Adamax bias correction uses steps_done + 1 in the next step)."""
        self.step_count = int(steps_done)
        _lib.check(self.lib.regat_engine_set_step(self._h, int(steps_done), _stream()))

    def get_step(self):
        """(steps taken, learning rate) as stored on the device; synchronises the current stream."""
        n, lr = C.c_int(), C.c_float()
        _lib.check(self.lib.regat_engine_get_step(self._h, C.byref(n), C.byref(lr), _stream()))
        return n.value, lr.value

    def train_step_dev(self, features, boxes, q_att, q_last, target):
        """One whole step (train.py:103-113) with lr / step taken from the device (set_lr, set_step): capturable in one CUDA
        graph.  Returns the 2-float device tensor (loss, score)."""
        B, N = self._inputs(features, boxes, q_att, q_last)
        self._chk(target, (B, self.cfg.num_answers))
        _lib.check(self.lib.regat_engine_train_step_dev(self._h, B, N, features.data_ptr(), boxes.data_ptr(), q_att.data_ptr(),
                                                        q_last.data_ptr(), target.data_ptr(), self._loss.data_ptr(), _stream()))
        return self._loss

    def set_dp(self, grad_ptrs, multicast_ptr, flag_ptrs, rank, world, blocks=32):
        """Data parallel inside train_step / train_step_dev: in-place exchange of each gradient range over NVLink
        (regat_engine_set_dp).  grad_ptrs / flag_ptrs: sequences of `world` device addresses."""
        if world <= 1:
            _lib.check(self.lib.regat_engine_set_dp(self._h, None, 0, None, 0, 1, 0))
            return
        gp = (C.c_uint64 * world)(*[int(p) for p in grad_ptrs])
        fp = (C.c_uint64 * world)(*[int(p) for p in flag_ptrs])
        _lib.check(self.lib.regat_engine_set_dp(self._h, gp, int(multicast_ptr), fp, int(rank), int(world), int(blocks)))

    def set_grad_callback(self, fn):
        """fn(offset, numel) is invoked inside fwd_bwd whenever grads[offset:offset+numel] is final on the current stream
        (tail of the buffer first).  Pass None to remove it."""
        if fn is None:
            self._cb = None
            _lib.check(self.lib.regat_engine_set_grad_callback(self._h, None, None))
            return
        self._cb = _lib.GRAD_READY_FN(lambda user, off, n: fn(int(off), int(n)))
        _lib.check(self.lib.regat_engine_set_grad_callback(self._h, C.cast(self._cb, C.c_void_p), None))

    def last_launches(self):
        return self.lib.regat_engine_last_launches(self._h)

    def profile_gemms(self, fn):
        """Runs fn() (eager engine calls) with timing events around every dense product; returns [(M, N, K, ms), ...] in
        launch order.  Measurement aid of bench.py (GEMM-class throughput inside the real step)."""
        _lib.check(self.lib.regat_engine_profile(self._h, 1))
        try:
            fn()
        finally:
            cap = 256
            mnk, ms, n = (C.c_int32 * (3 * cap))(), (C.c_float * cap)(), C.c_int()
            _lib.check(self.lib.regat_engine_profile_read(self._h, cap, mnk, ms, C.byref(n)))
            _lib.check(self.lib.regat_engine_profile(self._h, 0))
        return [(mnk[3 * i], mnk[3 * i + 1], mnk[3 * i + 2], ms[i]) for i in range(n.value)]

    def capture_train_step(self, features, boxes, q_att, q_last, target, stream=None):
        """CUDA graph of one whole train step on these (static) input tensors; see GraphedTrainStep."""
        return GraphedTrainStep(self, features, boxes, q_att, q_last, target, stream)

    def buffer(self, name, shape, dtype=None):
        """Copy of a named internal activation as a torch tensor (tests)."""
        p = C.c_void_p()
        _lib.check(self.lib.regat_engine_buffer(self._h, name.encode(), C.byref(p)))
        off = p.value - self.workspace.data_ptr()
        dt = dtype or (torch.bfloat16 if self.dtype == _lib.BF16 else torch.float32)
        n = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
        return self.workspace[off:off + n].view(dt).view(*shape).clone()


class GraphedTrainStep:
    """ONE CUDA graph = one whole train step (forward, backward, gradient exchange when data parallel, per-tensor clip, Adamax,
    re-derived bf16 kernels) on fixed input tensors.  Nothing in it depends on a host scalar: the learning rate and the step
    counter live on the device (HotPathEngine.set_lr / set_step), so `replay()` is the entire step.  The engine must have run
    at least one eager call before (streams, tensor maps and kernel attributes are created on first use)."""

    def __init__(self, engine: HotPathEngine, features, boxes, q_att, q_last, target, stream=None):
        self.engine = engine
        self.stream = stream or torch.cuda.Stream(engine.device)      # graphs cannot be captured on the default stream
        self.graph = torch.cuda.CUDAGraph()
        self.inputs = (features, boxes, q_att, q_last, target)      # keep the captured addresses alive
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            # eager warm-up that leaves the parameters alone: creates the side streams, tensor maps, shared-memory attributes
            engine.fwd_bwd(features, boxes, q_att, q_last, target)
            torch.cuda.synchronize()
            with torch.cuda.graph(self.graph, stream=self.stream):
                engine.train_step_dev(features, boxes, q_att, q_last, target)
            self.launches = engine.last_launches()

    def replay(self):
        """Enqueues the step on the CURRENT stream (stream-ordered after whatever the caller enqueued there, e.g. input copies)."""
        self.graph.replay()
        self.engine.step_count += 1
        return self.engine._loss
