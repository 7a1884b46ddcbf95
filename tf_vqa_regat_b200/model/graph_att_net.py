"""GraphAttentionNetwork -- mirrors model/graph_att_net.py:12-83.

call(v_feat, adj_mat, pos_emb) -> relu(self_weights(v_feat) + sum_d neighbor_net[d](...)) with ONE launch of the fused
attention kernel covering every direction and head (the sum over directions and the ReLU are its epilogue).
  implicit relation (pos_emb_dim > 0, label_num = 1): geometry bias rebuilt on chip, all-ones adjacency;
  explicit relation (pos_emb_dim = -1, label_num = 11 / 15): adj_mat [B,N,N,label_num] masks the affinities with -9e15 and the
  label FC adds a per-pair bias; the second direction uses the adjacency transposed in its object axes (:56)."""
import torch

from .. import _lib
from . import _rt
from .fc import FullyConnected
from .graph_att_layer import GraphSelfAttentionLayer, _geometry_args
from .weight_norm import Dropout, Layer


class GraphAttentionNetwork(Layer):
    def __init__(self, dir_num, label_num, in_feat_dim, out_feat_dim, nongt_dim=20, dropout=0.2, label_bias=True,
                 num_heads=16, pos_emb_dim=-1):
        assert dir_num <= 2, "Got more than two directions in a graph."          # graph_att_net.py:18
        self.dir_num = dir_num
        self.label_num = label_num
        self.in_feat_dim = in_feat_dim
        self.out_feat_dim = out_feat_dim
        self.dropout = Dropout(dropout)
        self.self_weights = FullyConnected([in_feat_dim, out_feat_dim], None, dropout)
        self.bias = FullyConnected([label_num, 1], None, 0.2, label_bias)
        self.nongt_dim = nongt_dim
        self.pos_emb_dim = pos_emb_dim
        self.neighbor_net = [GraphSelfAttentionLayer(pos_emb_dim=pos_emb_dim, num_heads=num_heads, hidden_dim=out_feat_dim,
                                                     nongt_dim=nongt_dim) for _ in range(dir_num)]
        self.num_heads = num_heads

    def label_constant(self, device):
        """graph_att_net.py:69-71 on the all-ones adjacency: every entry of v_biases_neighbors is WN(1x1)(1) (+ bias)."""
        ones = torch.ones(1, self.label_num, device=device)
        return self.bias(ones).view(1)

    def call(self, v_feat, adj_mat, pos_emb=None, residual=None):
        if self.pos_emb_dim > 0 and pos_emb is None:
            raise ValueError(f"position embedding is set to None with pos_emb_dim {self.pos_emb_dim}")     # :42-46
        elif self.pos_emb_dim < 0 and pos_emb is not None:
            raise ValueError("position embedding is NOT None with pos_emb_dim < 0")                        # :47-51
        if self.pos_emb_dim < 0:
            return self._call_explicit(v_feat, adj_mat, residual)
        v_feat = _rt.need_cuda(v_feat, "v_feat")
        B, N, _ = v_feat.shape
        D, dirs, H = self.out_feat_dim, self.dir_num, self.num_heads
        M = self.nongt_dim if self.nongt_dim < N else N
        s = self.self_weights(v_feat)                                              # :58
        q = _rt.empty(B * N, dirs * D, device=v_feat.device)
        kv = _rt.empty(B * M, 2 * dirs * D, device=v_feat.device)
        for d, net in enumerate(self.neighbor_net):
            net.project(s, q.data_ptr() + 4 * d * D, dirs * D, kv.data_ptr(), 2 * dirs * D, d * D, (dirs + d) * D)
        c = self.label_constant(v_feat.device)
        pps = [n.pair_pos_fc.dense for n in self.neighbor_net]
        alpha_g = torch.empty(dirs, device=v_feat.device)
        for d, pp in enumerate(pps):
            pp.alpha_ptr()
            alpha_g[d:d + 1].copy_(pp._stats[32:33])
        wstride = (pps[1].v.data_ptr() - pps[0].v.data_ptr()) // 4 if dirs > 1 else 0
        bstride = (pps[1].bias.data_ptr() - pps[0].bias.data_ptr()) // 4 if dirs > 1 else 0
        out = _rt.empty(B, N, D, device=v_feat.device)
        boxes, pe = _geometry_args(pos_emb, B, N, M, self.pos_emb_dim)
        wd = _lib.wave_divisors(self.pos_emb_dim)
        res = _rt.need_cuda(residual, "residual") if residual is not None else None
        _lib.check(_lib.lib().regat_geoattn_fwd(_rt.DT, B, N, self.nongt_dim, D, H, dirs, self.pos_emb_dim, q.data_ptr(), kv.data_ptr(),
                                                boxes, pe, wd.ctypes.data, pps[0].v.data_ptr(), wstride, alpha_g.data_ptr(),
                                                pps[0].bias.data_ptr(), bstride, c.data_ptr(), s.data_ptr(),
                                                res.data_ptr() if res is not None else None, int(res is not None), out.data_ptr(),
                                                None, None, None, _rt.stream()))
        return out                                                                 # relu(dropout(s + sum_d o_d)), :78-81

    def _call_explicit(self, v_feat, adj_mat, residual):
        """graph_att_net.py:56-81 with a labelled adjacency: per direction, input_adj = adj_d[:, :, :nongt] (:65), its sum over
        labels is the where-mask (:69, graph_att_layer.py:90-98), the label FC on it the additive bias (:71)."""
        v_feat = _rt.need_cuda(v_feat, "v_feat")
        adj = _rt.need_cuda(adj_mat, "adj_mat")
        B, N, _ = v_feat.shape
        D, dirs, H, L = self.out_feat_dim, self.dir_num, self.num_heads, self.label_num
        if tuple(adj.shape) != (B, N, N, L):
            raise ValueError(f"adj_mat must be [batch, num_rois, num_rois, label_num] = {(B, N, N, L)}; got {tuple(adj.shape)}")
        M = self.nongt_dim if self.nongt_dim < N else N
        s = self.self_weights(v_feat)                                              # :58
        q = _rt.empty(B * N, dirs * D, device=v_feat.device)
        kv = _rt.empty(B * M, 2 * dirs * D, device=v_feat.device)
        for d, net in enumerate(self.neighbor_net):
            net.project(s, q.data_ptr() + 4 * d * D, dirs * D, kv.data_ptr(), 2 * dirs * D, d * D, (dirs + d) * D)
        lab = self.bias.dense                                                      # WN Dense(label_num -> 1), shared by both directions
        if not lab.built:
            lab.build(L, v_feat.device)
        lab.alpha_ptr()
        w_eff = (lab.v.view(-1) * lab._stats[32]).contiguous()                     # effective kernel alpha * v  [L]
        pair_bias = _rt.empty(B, dirs, N, M, device=v_feat.device)
        l = _lib.lib()
        _lib.check(l.regat_explicit_pair_bias(B, N, self.nongt_dim, L, dirs, adj.data_ptr(), w_eff.data_ptr(),
                                              lab.bias.data_ptr() if lab.bias is not None else None, pair_bias.data_ptr(), _rt.stream()))
        out = _rt.empty(B, N, D, device=v_feat.device)
        res = _rt.need_cuda(residual, "residual") if residual is not None else None
        _lib.check(l.regat_graphattn_explicit_fwd(_rt.DT, B, N, self.nongt_dim, D, H, dirs, q.data_ptr(), kv.data_ptr(), pair_bias.data_ptr(),
                                                  s.data_ptr(), res.data_ptr() if res is not None else None, int(res is not None),
                                                  out.data_ptr(), None, None, _rt.stream()))
        return out
