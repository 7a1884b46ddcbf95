"""Question front-end on the GPU: WordEmbedding -> GRU -> QuestionSelfAttention of the reference's
model/language_model.py:10-174, forward and backward, fp32.  Produces the two tensors the hot path consumes --
q_emb_self_att (rel_graph_net.py:45) and q_emb = the last GRU state (:57) -- and consumes the dq_att / dq_last gradients
the hot path returns (HotPathEngine.fwd_bwd(want_dq=True)), so that the whole model trains on the device.

Every number comes from libregat.so: the dense products are regat_gemm calls (exact-fp32 kernel), everything else the
regat_q_* kernels of csrc/question.cu; torch supplies device memory and the stream.  This module only sequences launches.

Reference behaviour kept: padding tokens (== n_token) embed to zero but still run through the GRU (no mask); the second
table `emb_` (op 'c') is frozen unless emb2_trainable (language_model.py:58,79); dropout is inert; the attention softmax
runs over the BATCH axis and its [T,B] result is raw-reshaped to [B,1,T] (:163-167) -- so q_att of one question depends
on the other questions of the batch, and batch 1 is refused.  The reference runs the GRU twice per step on the same input
(rel_graph_net.py:44,57); the second run reproduces the first bit for bit, so it is computed once here.

Data parallel (group given): the batch-axis softmax is the front-end's one real exchange step.  Every rank gathers the
[B_local, T] attention logits of all ranks (14 floats per question), normalises over the GLOBAL batch and reads its own rows
of the raw-reshaped weight matrix; backward gathers the dW rows the same way.  With equal shards this reproduces the
reference's single-process numbers; weight gradients are then summed over ranks (allreduce_grads) like the hot path's.

Status: first correct path, one launch sequence per call (about 3T + 12 launches forward); not tuned.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

ALIGN = 64


def question_layout(n_token, emb_dim, num_hid, op="c"):
    """[(name, shape, offset)] + total elements: Keras variable order of w_emb, q_emb, q_att (rel_graph_net.py:16-18),
    every tensor 256-byte aligned in one flat fp32 buffer."""
    e_in = emb_dim * (2 if "c" in op else 1)
    shapes = [("w_emb.emb/emb", (n_token + 1, emb_dim))]
    if "c" in op:
        shapes.append(("w_emb.emb_/emb_", (n_token + 1, emb_dim)))
    shapes += [("q_emb.gru/kernel", (e_in, 3 * num_hid)), ("q_emb.gru/recurrent_kernel", (num_hid, 3 * num_hid)),
               ("q_emb.gru/bias", (2, 3 * num_hid)),
               ("q_att.linear1/v", (num_hid, num_hid)), ("q_att.linear1/g", ()), ("q_att.linear1/bias", (num_hid,)),
               ("q_att.linear2/v", (num_hid, 1)), ("q_att.linear2/g", ()), ("q_att.linear2/bias", (1,))]
    out, off = [], 0
    for name, shape in shapes:
        n = int(np.prod(shape)) if shape else 1
        out.append((name, shape, off))
        off = (off + n + ALIGN - 1) // ALIGN * ALIGN
    return out, off


class QuestionFrontEnd:
    def __init__(self, n_token, emb_dim, num_hid, op="c", seq_len=14, max_batch=256, emb2_trainable=False, device="cuda:0",
                 grad_clip=0.25, beta1=0.9, beta2=0.999, eps=1e-8, group=None, _ops=None):
        self.n_token, self.E, self.H, self.op, self.T, self.max_batch = n_token, emb_dim, num_hid, op, seq_len, max_batch
        self.Ein = emb_dim * (2 if "c" in op else 1)
        self.emb2_trainable = bool(emb2_trainable and "c" in op)
        self.grad_clip, self.beta1, self.beta2, self.eps = grad_clip, beta1, beta2, eps
        if _ops is None:
            if not torch.cuda.is_available():
                raise _lib.RegatError(-6, "QuestionFrontEnd needs a CUDA device; there is no CPU fallback")
            self.L = _lib.lib()
        else:
            self.L = _ops                                   # tests only: a host emulation of the entry points (dry run of the sequencing)
        self.device = torch.device(device)
        self.group, self.world, self.rank = group, 1, 0
        if group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.entries, total = question_layout(n_token, emb_dim, num_hid, op)
        f = lambda n: torch.zeros(int(n), dtype=torch.float32, device=self.device)
        self.params, self.grads, self.adamax_m, self.adamax_u = f(total), f(total), f(total), f(total)
        B, T, H, Ein = max_batch, seq_len, num_hid, self.Ein
        self._X, self._XI, self._HI = f(B * T * Ein), f(B * T * 3 * H), f(T * B * 3 * H)
        self._Z, self._R, self._C, self._HP = f(T * B * H), f(T * B * H), f(T * B * H), f(T * B * H)
        self._SEQ, self._A1, self._LOG, self._P = f(B * T * H), f(B * T * H), f(B * T), f(self.world * B * T)
        if self.world > 1:                                  # global-batch copies of the attention logits and of their gradients
            self._LOG_G, self._DW_G, self._DLOG_G = f(self.world * B * T), f(self.world * B * T), f(self.world * B * T)
        self._ZERO = f(B * H)
        self._DSEQ, self._DW, self._DLOG, self._DA1 = f(B * T * H), f(B * T), f(B * T), f(B * T * H)
        self._DXI, self._DHI, self._DX = f(B * T * 3 * H), f(T * B * 3 * H), f(B * T * Ein)
        self._DH = [f(B * H), f(B * H)]
        self._G1, self._G2 = f(H * H), f(H)
        self._ONES = torch.ones(B * T, dtype=torch.float32, device=self.device)
        self._scal = f(64)                                  # [0]=vv1 [1]=vv2 [2]=alpha1 [3]=alpha2 [4]=Gv1 [5]=Gv2 [16+k]=||grad_k||^2
        self._saved_B = None
        self.step_count = 0

    # ---- parameters
    def named(self, buf=None):
        buf = self.params if buf is None else buf
        return {n: buf[o:o + (int(np.prod(s)) if s else 1)].view(s) for n, s, o in self.entries}

    def load_named(self, named):
        for n, t in self.named().items():
            t.copy_(torch.as_tensor(np.asarray(named[n], dtype=np.float32)).reshape(t.shape))

    def trainable(self):
        return [n for n, _, _ in self.entries if n != "w_emb.emb_/emb_" or self.emb2_trainable]

    # ---- plumbing
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None

    def _p(self, name, buf=None):
        for n, s, o in self.entries:
            if n == name:
                return (self.params if buf is None else buf).data_ptr() + 4 * o
        raise KeyError(name)

    def _gemm(self, tA, tB, M, N, K, A, lda, B, ldb, Cp, ldc, alpha=None, bias=None, acc=False):
        epi = _lib.Epilogue()
        epi.alpha, epi.alpha_cols, epi.bias, epi.accumulate = alpha, 0, bias, int(acc)
        _lib.check(self.L.regat_gemm(_lib.F32, tA, tB, M, N, K, A, lda, B, ldb, Cp, ldc, _lib.F32, C.byref(epi), self._stream()))

    # ---- forward: rel_graph_net.py:41-45,57
    def forward(self, tokens):
        """tokens int32 [B, T] on the device -> (q_emb_self_att [B,H], q_emb [B,H])."""
        L, st, T, H, E, Ein = self.L, self._stream(), self.T, self.H, self.E, self.Ein
        if tokens.dtype != torch.int32 or tokens.dim() != 2 or tokens.shape[1] != T or not tokens.is_contiguous():
            raise ValueError(f"tokens must be a contiguous int32 [B, {T}] tensor")
        if tokens.device != self.device:
            raise _lib.RegatError(-4, f"tokens must live on {self.device}")
        B = tokens.shape[0]
        if B < 2 or B > self.max_batch:
            raise ValueError(f"batch must be in 2..{self.max_batch} (tf.squeeze drops a batch of 1, language_model.py:159)")
        ck, ptr = _lib.check, lambda t: t.data_ptr()
        self._tok, self._saved_B = tokens, B
        emb2 = self._p("w_emb.emb_/emb_") if "c" in self.op else None
        ck(L.regat_q_embed_fwd(ptr(tokens), B * T, self.n_token, E, self._p("w_emb.emb/emb"), emb2, ptr(self._X), st))
        Wk, U, b0 = self._p("q_emb.gru/kernel"), self._p("q_emb.gru/recurrent_kernel"), self._p("q_emb.gru/bias")
        b1 = b0 + 4 * 3 * H
        self._gemm(0, 0, B * T, 3 * H, Ein, ptr(self._X), Ein, Wk, 3 * H, ptr(self._XI), 3 * H, bias=b0)
        seq = ptr(self._SEQ)
        for t in range(T):
            hp = None if t == 0 else seq + 4 * (t - 1) * H
            hi = ptr(self._HI) + 4 * t * B * 3 * H
            self._gemm(0, 0, B, 3 * H, H, ptr(self._ZERO) if t == 0 else hp, H if t == 0 else T * H, U, 3 * H, hi, 3 * H, bias=b1)
            o = 4 * t * B * H
            ck(L.regat_q_gru_gates_fwd(B, H, ptr(self._XI) + 4 * t * 3 * H, T * 3 * H, hi, hp, T * H, seq + 4 * t * H, T * H,
                                       ptr(self._Z) + o, ptr(self._R) + o, ptr(self._C) + o, ptr(self._HP) + o, st))
        sc = ptr(self._scal)
        self._scal[:6].zero_()
        v1, v2 = self._p("q_att.linear1/v"), self._p("q_att.linear2/v")
        ck(L.regat_q_dot(v1, v1, H * H, sc, st)); ck(L.regat_q_dot(v2, v2, H, sc + 4, st))
        ck(L.regat_q_wn_alpha(self._p("q_att.linear1/g"), sc, sc + 8, st)); ck(L.regat_q_wn_alpha(self._p("q_att.linear2/g"), sc + 4, sc + 12, st))
        self._gemm(0, 0, B * T, H, H, seq, H, v1, H, ptr(self._A1), H, alpha=sc + 8, bias=self._p("q_att.linear1/bias"))
        ck(L.regat_q_tanh_fwd(ptr(self._A1), B * T * H, st))
        self._gemm(0, 0, B * T, 1, H, ptr(self._A1), H, v2, 1, ptr(self._LOG), 1, alpha=sc + 12, bias=self._p("q_att.linear2/bias"))
        log_ptr, Bg, roff = ptr(self._LOG), B, 0
        if self.world > 1:                                  # language_model.py:163-165 normalises over the whole batch
            import torch.distributed as dist
            Bg, roff = B * self.world, 4 * self.rank * B * T
            dist.all_gather_into_tensor(self._LOG_G[:Bg * T], self._LOG[:B * T], group=self.group)
            log_ptr = ptr(self._LOG_G)
        self._Bg, self._roff = Bg, roff
        ck(L.regat_q_batch_softmax_fwd(log_ptr, Bg, T, ptr(self._P), st))
        q_att = torch.empty(B, H, dtype=torch.float32, device=self.device)
        ck(L.regat_q_pool_fwd(ptr(self._P) + roff, seq, B, T, H, ptr(q_att), st))
        q_last = self._SEQ[:B * T * H].view(B, T, H)[:, T - 1].contiguous()         # language_model.py:120 output[:, -1]
        return q_att, q_last

    # ---- backward: what tape.gradient (train.py:111) computes for the front-end's variables
    def backward(self, dq_att, dq_last):
        """dq_att, dq_last fp32 [B,H] (from the hot path) -> self.grads holds the tape gradients of every front-end variable."""
        L, st, T, H, E, Ein, B = self.L, self._stream(), self.T, self.H, self.E, self.Ein, self._saved_B
        if B is None:
            raise RuntimeError("backward() before forward()")
        for t in (dq_att, dq_last):
            if tuple(t.shape) != (B, H) or t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
                raise ValueError(f"dq_att / dq_last must be contiguous fp32 [{B}, {H}] on {self.device}")
        ck, ptr = _lib.check, lambda t: t.data_ptr()
        g = lambda name: self._p(name, self.grads)
        self.grads.zero_()
        self._scal[4:6].zero_()
        sc, seq, ones = ptr(self._scal), ptr(self._SEQ), ptr(self._ONES)
        v1, v2 = self._p("q_att.linear1/v"), self._p("q_att.linear2/v")
        Bg, roff = self._Bg, self._roff
        ck(L.regat_q_pool_bwd(ptr(self._P) + roff, seq, ptr(dq_att), ptr(dq_last), B, T, H, ptr(self._DSEQ), ptr(self._DW), st))
        dw_ptr, dlog = ptr(self._DW), ptr(self._DLOG)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self._DW_G[:Bg * T], self._DW[:B * T], group=self.group)
            dw_ptr, dlog = ptr(self._DW_G), ptr(self._DLOG_G)
        ck(L.regat_q_batch_softmax_bwd(ptr(self._P), dw_ptr, Bg, T, dlog, st))
        dlog += roff                                        # this rank's rows of the global dlogits
        # linear2 (H -> 1): G2 = a1^T dlogits, bias gradient, da1 = dlogits (alpha2 v2)^T
        self._gemm(1, 0, H, 1, B * T, ptr(self._A1), H, dlog, 1, ptr(self._G2), 1)
        self._gemm(0, 0, 1, 1, B * T, ones, B * T, dlog, 1, g("q_att.linear2/bias"), 1)
        self._gemm(0, 1, B * T, H, 1, dlog, 1, v2, 1, ptr(self._DA1), H, alpha=sc + 12)
        ck(L.regat_q_tanh_bwd(ptr(self._DA1), ptr(self._A1), B * T * H, st))
        # linear1 (H -> H): G1 = seq^T da1, bias gradient, dseq += da1 (alpha1 v1)^T
        self._gemm(1, 0, H, H, B * T, seq, H, ptr(self._DA1), H, ptr(self._G1), H)
        self._gemm(0, 0, 1, H, B * T, ones, B * T, ptr(self._DA1), H, g("q_att.linear1/bias"), H)
        self._gemm(0, 1, B * T, H, H, ptr(self._DA1), H, v1, H, ptr(self._DSEQ), H, alpha=sc + 8, acc=True)
        ck(L.regat_q_dot(ptr(self._G1), v1, H * H, sc + 16, st)); ck(L.regat_q_dot(ptr(self._G2), v2, H, sc + 20, st))
        ck(L.regat_q_wn_bwd(ptr(self._G1), v1, self._p("q_att.linear1/g"), sc, sc + 16, H * H, g("q_att.linear1/v"), g("q_att.linear1/g"), st))
        ck(L.regat_q_wn_bwd(ptr(self._G2), v2, self._p("q_att.linear2/g"), sc + 4, sc + 20, H, g("q_att.linear2/v"), g("q_att.linear2/g"), st))
        # back-propagation through time
        U, Wk = self._p("q_emb.gru/recurrent_kernel"), self._p("q_emb.gru/kernel")
        dh_rec = None
        for t in range(T - 1, -1, -1):
            out = ptr(self._DH[t % 2])
            o = 4 * t * B * H
            dhi = ptr(self._DHI) + 4 * t * B * 3 * H
            ck(L.regat_q_gru_gates_bwd(B, H, ptr(self._DSEQ) + 4 * t * H, T * H, dh_rec, ptr(self._Z) + o, ptr(self._R) + o, ptr(self._C) + o,
                                       ptr(self._HP) + o, ptr(self._HI) + 4 * t * B * 3 * H, ptr(self._DXI) + 4 * t * 3 * H, T * 3 * H, dhi, out, st))
            if t > 0:
                self._gemm(0, 1, B, H, 3 * H, dhi, 3 * H, U, 3 * H, out, H, acc=True)
            dh_rec = out
        gb = g("q_emb.gru/bias")
        self._gemm(1, 0, H, 3 * H, T * B, ptr(self._HP), H, ptr(self._DHI), 3 * H, g("q_emb.gru/recurrent_kernel"), 3 * H)
        self._gemm(0, 0, 1, 3 * H, T * B, ones, T * B, ptr(self._DHI), 3 * H, gb + 4 * 3 * H, 3 * H)
        self._gemm(1, 0, Ein, 3 * H, B * T, ptr(self._X), Ein, ptr(self._DXI), 3 * H, g("q_emb.gru/kernel"), 3 * H)
        self._gemm(0, 0, 1, 3 * H, B * T, ones, B * T, ptr(self._DXI), 3 * H, gb, 3 * H)
        self._gemm(0, 1, B * T, Ein, 3 * H, ptr(self._DXI), 3 * H, Wk, 3 * H, ptr(self._DX), Ein)
        demb2 = g("w_emb.emb_/emb_") if self.emb2_trainable else None
        ck(L.regat_q_embed_bwd(ptr(self._tok), B * T, self.n_token, E, Ein, ptr(self._DX), g("w_emb.emb/emb"), demb2, st))

    def allreduce_grads(self):
        """Data parallel: sum the front-end's weight gradients over the ranks (the hot path hands over dq_att / dq_last already
        scaled by 1/R when its loss is the mean over the global batch)."""
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.grads, group=self.group)
            # the embedding tables' clip norm and Adamax `u` are functions of the per-occurrence values of the WHOLE batch
            # (IndexedSlices semantics, see update()): every rank needs every rank's tokens and lookup-output gradients
            B, T = self._saved_B, self.T
            if getattr(self, "_tok_g", None) is None:
                self._tok_g = torch.zeros(self.world * self.max_batch * T, dtype=torch.int32, device=self.device)
                self._DX_g = torch.zeros(self.world * self.max_batch * T * self.Ein, dtype=torch.float32, device=self.device)
            dist.all_gather_into_tensor(self._tok_g[:self.world * B * T], self._tok.reshape(-1)[:B * T].contiguous(), group=self.group)
            dist.all_gather_into_tensor(self._DX_g[:self.world * B * T * self.Ein], self._DX[:B * T * self.Ein], group=self.group)
            self._gathered = True

    # ---- train.py:112-113 for the front-end's variables
    def update(self, lr, step=None):
        L, st = self.L, self._stream()
        self.step_count = step if step is not None else self.step_count + 1
        sc = self._scal.data_ptr()
        names = self.trainable()
        self._scal[32:32 + len(names)].zero_()
        B = self._saved_B
        for k, name in enumerate(names):
            n = [int(np.prod(s)) if s else 1 for nm, s, _ in self.entries if nm == name][0]
            gp, ss = self._p(name, self.grads), sc + 4 * (32 + k)
            if name.startswith("w_emb."):
                # the table was read through tf.nn.embedding_lookup: its tape gradient is IndexedSlices, which train.py:112 clips by
                # the norm of the per-occurrence values and Adamax applies through its sparse branch (duplicate tokens are visible)
                if B is None:
                    raise RuntimeError("update() before forward() / backward()")
                col0 = 0 if name == "w_emb.emb/emb" else self.E
                if getattr(self, "_UINC", None) is None:
                    self._UINC = torch.zeros((self.n_token + 1) * self.E, dtype=torch.float32, device=self.device)
                if self.world > 1 and not getattr(self, "_gathered", False):
                    raise RuntimeError("data parallel: call allreduce_grads() before update() (the embedding update needs every rank's tokens)")
                tokp, dxp, BT = ((self._tok_g.data_ptr(), self._DX_g.data_ptr(), self.world * B * self.T) if self.world > 1
                                 else (self._tok.data_ptr(), self._DX.data_ptr(), B * self.T))
                _lib.check(L.regat_q_embed_sumsq(tokp, BT, self.n_token, self.E, self.Ein, col0, dxp, ss, st))
                _lib.check(L.regat_q_embed_clip_adamax(tokp, BT, self.n_token, self.E, self.Ein, col0, dxp,
                                                       self._p(name), gp, self._p(name, self.adamax_m), self._p(name, self.adamax_u),
                                                       self._UINC.data_ptr(), ss, self.grad_clip, float(lr), int(self.step_count), self.beta1,
                                                       self.beta2, self.eps, st))
                continue
            _lib.check(L.regat_q_dot(gp, gp, n, ss, st))
            _lib.check(L.regat_q_clip_adamax(self._p(name), gp, self._p(name, self.adamax_m), self._p(name, self.adamax_u), n, ss,
                                             self.grad_clip, float(lr), int(self.step_count), self.beta1, self.beta2, self.eps, st))
        self._gathered = False
