"""concat_visual_question, ImplicitRelationEncoder and ExplicitRelationEncoder -- mirror model/relation_encoder.py:13-143."""
import torch

from .. import _lib
from . import _rt
from .fc import FullyConnected
from .graph_att_net import GraphAttentionNetwork
from .weight_norm import Layer


def concat_visual_question(q, v, mask=True):
    """[v || mask * q] with mask[b,n] = (sum_d v[b,n,d] != 0): the implicit path's only padded-object masking
    (relation_encoder.py:13-37).  mask=False is a NameError in the reference (masked_q undefined); refused here."""
    if not mask:
        raise NotImplementedError("concat_visual_question(mask=False) is broken in the reference (relation_encoder.py:35)")
    q, v = _rt.need_cuda(q, "q"), _rt.need_cuda(v, "v")
    B, N, D = v.shape
    Q = q.shape[1]
    out = _rt.empty(B, N, D + Q, device=v.device)
    _lib.check(_lib.lib().regat_concat_visual_question(_rt.DT, B, N, D, Q, v.data_ptr(), q.data_ptr(), out.data_ptr(), None,
                                                       _rt.stream()))
    return out


class ImplicitRelationEncoder(Layer):
    def __init__(self, v_dim, q_dim, out_dim, dir_num, pos_emb_dim, nongt_dim, num_heads=16, num_steps=1,
                 residual_connection=True, label_bias=True):
        self.v_dim = v_dim
        self.q_dim = q_dim
        self.out_dim = out_dim
        self.residual_connection = residual_connection
        self.num_steps = num_steps
        self.v2out = FullyConnected([v_dim, out_dim], dropout=0.2) if self.v_dim != self.out_dim else None   # :52-55
        in_dim = out_dim + q_dim
        self.implicit_relation = GraphAttentionNetwork(dir_num, 1, in_dim, out_dim, nongt_dim=nongt_dim, label_bias=label_bias,
                                                       num_heads=num_heads, pos_emb_dim=pos_emb_dim)

    def call(self, visual, pos_emb, question):
        """visual [B,N,v_dim], pos_emb [B,M,N,E] tensor or BoxGeometry, question [B,q_dim] -> [B,N,out_dim]."""
        visual = _rt.need_cuda(visual, "visual")
        question = _rt.need_cuda(question, "question")
        B, N = visual.shape[0], visual.shape[1]
        adj_mat = None            # the reference builds ones[B,N,N,1] (:76); the kernels treat it as all-ones implicitly
        if self.v2out:
            visual = self.v2out(visual)                                             # :78-79
        for _ in range(self.num_steps):                                             # :82
            v_cat_q = concat_visual_question(question, visual, mask=True)
            # the residual add of :88-89 is fused into the attention kernel's epilogue
            visual = self.implicit_relation(v_cat_q, adj_mat, pos_emb, residual=visual if self.residual_connection else None)
        return visual


class ExplicitRelationEncoder(Layer):
    """relation_encoder.py:95-143 (spatial: label_num 11, semantic: 15; built by rel_graph_net.py:78-92).

    The reference spells the constructor argument `residiual_connection` (:98) while its body (:104) and its call sites
    (rel_graph_net.py:84,91) say `residual_connection`, so the class cannot be constructed there as written; both spellings are
    accepted here and mean the same thing.  v2out has no dropout argument in the explicit encoder (:109)."""

    def __init__(self, v_dim, q_dim, out_dim, dir_num, label_num, nongt_dim=20, num_heads=16, num_steps=1,
                 residiual_connection=True, label_bias=True, residual_connection=None):
        self.v_dim = v_dim
        self.q_dim = q_dim
        self.out_dim = out_dim
        self.residual_connection = residiual_connection if residual_connection is None else residual_connection
        self.num_steps = num_steps
        self.v2out = FullyConnected([v_dim, out_dim]) if self.v_dim != self.out_dim else None                # :108-111
        in_dim = out_dim + q_dim
        self.explicit_relation = GraphAttentionNetwork(dir_num, label_num, in_dim, out_dim, nongt_dim=nongt_dim, label_bias=label_bias,
                                                       num_heads=num_heads, pos_emb_dim=-1)

    def call(self, visual, adj_mat, question):
        """visual [B,N,v_dim], adj_mat [B,N,N,label_num], question [B,q_dim] -> [B,N,out_dim]."""
        visual = _rt.need_cuda(visual, "visual")
        question = _rt.need_cuda(question, "question")
        if self.v2out:
            visual = self.v2out(visual)                                             # :131-132
        for _ in range(self.num_steps):                                             # :134
            v_cat_q = concat_visual_question(question, visual, mask=True)
            # the residual add of :138-139 is fused into the attention kernel's epilogue
            visual = self.explicit_relation(v_cat_q, adj_mat, residual=visual if self.residual_connection else None)
        return visual
