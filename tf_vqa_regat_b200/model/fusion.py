"""BUTD -- mirrors model/fusion.py:12-54: question-guided attention pooling over objects + joint embedding.

All five FullyConnected layers are plain linear (the reference passes `dropout` in the `activation` slot, fusion.py:15-20);
that linearity is what lets the [B*N, v_dim] x [v_dim, hidden] v2attention product be re-associated away:
    logit[b,n] = sum_c (v[b,n] Wva + bva)_c * u[b,c] * wl_c + bl = < v[b,n], Wva (u[b] * wl) > + < bva, u[b] * wl > + bl."""
import torch

from .. import _lib
from . import _rt
from .fc import FullyConnected
from .weight_norm import Dropout, Layer


class BUTD(Layer):
    def __init__(self, v_dim, q_dim, hidden_dim, dropout=0.2):
        self.v2attention = FullyConnected([v_dim, hidden_dim], dropout)
        self.q2attention = FullyConnected([q_dim, hidden_dim], dropout)
        self.dropout = Dropout(dropout)
        self.linear = FullyConnected([hidden_dim, 1], dropout)
        self.visual_embed = FullyConnected([v_dim, hidden_dim], dropout)
        self.question_embed = FullyConnected([q_dim, hidden_dim], dropout)
        self._dims = (v_dim, q_dim, hidden_dim)

    def _build(self, device):
        v_dim, q_dim, h = self._dims
        for fc, d in ((self.v2attention, v_dim), (self.q2attention, q_dim), (self.linear, h), (self.visual_embed, v_dim),
                      (self.question_embed, q_dim)):
            if not fc.dense.built:
                fc.dense.build(d, device)

    def _pool(self, visual, question):
        visual, question = _rt.need_cuda(visual, "visual"), _rt.need_cuda(question, "question")
        B, N, D = visual.shape
        Hd = self._dims[2]
        self._build(visual.device)
        l = _lib.lib()
        u = self.q2attention(question)                                                   # fusion.py:48
        va, lin = self.v2attention.dense, self.linear.dense
        uw = _rt.empty(B, Hd, device=visual.device)
        cb = _rt.empty(B, device=visual.device)
        _lib.check(l.regat_butd_prep(_rt.DT, B, Hd, u.data_ptr(), Hd, lin.v.data_ptr(), lin.alpha_ptr(),
                                     va.bias.data_ptr() if va.bias is not None else None,
                                     lin.bias.data_ptr() if lin.bias is not None else None, uw.data_ptr(), cb.data_ptr(), _rt.stream()))
        weff = _rt.empty(B, D, device=visual.device)
        epi = _lib.Epilogue()
        epi.alpha = va.alpha_ptr()
        _rt.gemm(0, 1, B, D, Hd, uw.data_ptr(), Hd, va.v.data_ptr(), Hd, weff.data_ptr(), D, epi)
        att = _rt.empty(B, N, device=visual.device)
        pooled = _rt.empty(B, D, device=visual.device)
        _lib.check(l.regat_butd_pool_fwd(_rt.DT, B, N, D, visual.data_ptr(), weff.data_ptr(), cb.data_ptr(), att.data_ptr(),
                                         pooled.data_ptr(), _rt.stream()))
        return att.view(B, N, 1), pooled

    def attention_weights(self, visual, question):
        """softmax over the N objects, padded rows included (fusion.py:43-54) -> [B, N, 1]."""
        return self._pool(visual, question)[0]

    def call(self, visual, question):
        """-> (joint_emb [B, hidden], weights [B, N, 1])   (fusion.py:22-41)."""
        weights, pooled = self._pool(visual, question)
        pv = self.visual_embed(pooled)
        qe = self.question_embed(_rt.need_cuda(question, "question"))
        joint = torch.empty_like(pv)
        B, Hd = pv.shape
        _lib.check(_lib.lib().regat_mul(_rt.DT, B, Hd, pv.data_ptr(), Hd, qe.data_ptr(), Hd, joint.data_ptr(), Hd, _rt.stream()))
        return joint, weights
