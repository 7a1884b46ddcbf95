"""GraphSelfAttentionLayer -- mirrors model/graph_att_layer.py:14-121.

Same constructor and call signature; variables in the reference's order (pair_pos_fc, query, key, linear_out_).  call()
returns the layer's raw output [B, N, hidden] (graph_att_layer.py:121).  The arithmetic is the fused kernel's:
  * value projection re-association: the grouped 1x1 conv of (softmax . roi[:, :M]) equals softmax . (roi[:, :M] Kc + bc)
    because Conv2D(groups=H) reads, for head h, all `hidden` channels of block h (graph_att_layer.py:31-37,110-117);
  * the geometry bias is rebuilt on chip from boxes (BoxGeometry) or read from a materialised pos_emb, with the reference's
    raw-reshape index scramble (graph_att_layer.py:74,81);
  * implicit relation: adj_mat must be the all-ones adjacency (relation_encoder.py:76): tf.where is then a no-op; label_att
    must be the constant produced by the label FC on ones (graph_att_net.py:71) -- its first element is used;
  * explicit relation (pos_emb None, pos_emb_dim = -1): adj_mat [B,N,M] masks the affinities with -9e15 (:90-98) and
    label_att [B,N,M] is added after the mask (:100)."""

from .. import _lib
from . import _rt
from .fc import FullyConnected
from .position_emb import BoxGeometry
from .weight_norm import Conv2D, Layer, WeightNorm


class GraphSelfAttentionLayer(Layer):
    def __init__(self, hidden_dim, nongt_dim, pos_emb_dim=-1, num_heads=16, dropout=[0.2, 0.5]):
        self.num_heads = num_heads
        self.hidden_dim = hidden_dim
        self.pos_emb_dim = pos_emb_dim
        self.fc_dim = num_heads
        self.head_dim = int(hidden_dim / num_heads)
        if self.pos_emb_dim > 0:
            self.pair_pos_fc = FullyConnected([pos_emb_dim, self.fc_dim], activation=None, dropout=dropout[0])
        self.query = FullyConnected([hidden_dim, hidden_dim], None, dropout[0])
        self.key = FullyConnected([hidden_dim, hidden_dim], None, dropout[0])
        self.nongt_dim = nongt_dim
        self.linear_out_ = WeightNorm(Conv2D(filters=hidden_dim, input_shape=(None, 1, 1, self.fc_dim * hidden_dim),
                                             kernel_size=(1, 1), groups=self.fc_dim))

    def _build_all(self, device):
        D = self.hidden_dim
        if self.pos_emb_dim > 0 and not self.pair_pos_fc.dense.built:
            self.pair_pos_fc.dense.build(self.pos_emb_dim, device)
        for fc in (self.query, self.key):
            if not fc.dense.built:
                fc.dense.build(D, device)
        if not self.linear_out_.built:
            self.linear_out_.build(self.fc_dim * D, device)

    def project(self, roi, q_out, q_ld, kv_base, kv_ld, k_col, v_col):
        """Q = query(roi) -> q_out;  K = key(roi[:, :M]) and V' = roi[:, :M] Kc + bc -> column slices of the KV buffer."""
        B, N, D = roi.shape
        M = self.nongt_dim if self.nongt_dim < N else N                        # graph_att_layer.py:42
        self._build_all(roi.device)
        trunc = roi[:, :M, :].contiguous()                                      # graph_att_layer.py:43
        self.query.dense.call(roi, out=q_out, out_ld=q_ld)                      # :47
        self.key.dense.call(trunc, out=kv_base + 4 * k_col, out_ld=kv_ld)       # :55
        lo = self.linear_out_
        epi = _lib.Epilogue()
        epi.alpha = lo.alpha_ptr()
        epi.bias = lo.bias.data_ptr()
        _rt.gemm(0, 0, B * M, D, D, trunc.data_ptr(), D, lo.v.data_ptr(), D, kv_base + 4 * v_col, kv_ld, epi)
        return M

    def call(self, roi, adj_mat, pos_emb, label_att):
        roi = _rt.need_cuda(roi, "roi")
        B, N, D = roi.shape
        M = self.nongt_dim if self.nongt_dim < N else N
        if self.pos_emb_dim <= 0 or pos_emb is None:
            # explicit relation, one direction (graph_att_layer.py:90-102): adj_mat [B,N,M] (condensed: > 0 where an edge exists),
            # label_att [B,N,M]; masked pairs get -9e15 (which absorbs the label term in fp32 exactly as where(...) + label_att does)
            import torch
            adj = _rt.need_cuda(adj_mat, "adj_mat")
            lab = _rt.need_cuda(label_att, "label_att")
            if tuple(adj.shape) != (B, N, M) or tuple(lab.shape) != (B, N, M):
                raise ValueError(f"adj_mat and label_att must be [batch, num_rois, nongt] = {(B, N, M)}")
            pair_bias = torch.where(adj > 0, lab, torch.full_like(lab, -9e15)).contiguous()
            q = _rt.empty(B * N, D, device=roi.device)
            kv = _rt.empty(B * M, 2 * D, device=roi.device)
            self.project(roi, q.data_ptr(), D, kv.data_ptr(), 2 * D, 0, D)
            out = _rt.empty(B, N, D, device=roi.device)
            _lib.check(_lib.lib().regat_graphattn_explicit_fwd(_rt.DT, B, N, self.nongt_dim, D, self.num_heads, 1, q.data_ptr(), kv.data_ptr(),
                                                               pair_bias.data_ptr(), None, None, 0, out.data_ptr(), None, None,
                                                               _rt.stream()))
            return out
        q = _rt.empty(B * N, D, device=roi.device)
        kv = _rt.empty(B * M, 2 * D, device=roi.device)
        self.project(roi, q.data_ptr(), D, kv.data_ptr(), 2 * D, 0, D)
        out = _rt.empty(B, N, D, device=roi.device)
        pp = self.pair_pos_fc.dense
        a_ptr = pp.alpha_ptr()
        label_ptr = _rt.need_cuda(label_att, "label_att").data_ptr() if label_att is not None else None
        boxes, pe = _geometry_args(pos_emb, B, N, M, self.pos_emb_dim)
        wd = _lib.wave_divisors(self.pos_emb_dim)
        _lib.check(_lib.lib().regat_geoattn_fwd(_rt.DT, B, N, self.nongt_dim, D, self.num_heads, 1, self.pos_emb_dim, q.data_ptr(),
                                                kv.data_ptr(), boxes, pe, wd.ctypes.data, pp.v.data_ptr(), 0, a_ptr,
                                                pp.bias.data_ptr() if pp.bias is not None else None, 0, label_ptr, None, None, 0,
                                                out.data_ptr(), None, None, None, _rt.stream()))
        return out


def _geometry_args(pos_emb, B, N, M, E):
    """(boxes_ptr, pos_emb_ptr): exactly one is non-null."""
    if isinstance(pos_emb, BoxGeometry):
        if tuple(pos_emb.boxes.shape) != (B, N, 4):
            raise ValueError(f"BoxGeometry boxes {tuple(pos_emb.boxes.shape)} do not match roi batch {(B, N)}")
        return pos_emb.boxes.data_ptr(), None
    pe = _rt.need_cuda(pos_emb, "pos_emb")
    if pe.numel() != B * M * N * E:
        raise ValueError(f"pos_emb has {pe.numel()} elements, expected [B={B}, M={M}, N={N}, {E}]")
    return None, pe.data_ptr()
