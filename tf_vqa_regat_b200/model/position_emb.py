"""Stage 1 -- mirrors model/position_emb.py:96-160 (prepare_graph_variables and its two helpers) on the GPU.

prepare_graph_variables keeps the reference's signature and return value (pos_emb, None, None).  Two forms of pos_emb:
  * default: the materialised [B, M, N, 64] tensor the reference builds on the host (compat path);
  * lazy=True: a BoxGeometry handle (boxes + nongt_dim).  The attention layers accept it in place of the tensor and
    rebuild every embedding on chip, so the 47-655 MB array never exists (the fast path, SURVEY 8b "extension needed")."""
import numpy as np
import torch

from .. import _lib
from . import _rt


class BoxGeometry:
    def __init__(self, boxes, nongt_dim, feat_dim=64):
        self.boxes, self.nongt_dim, self.feat_dim = boxes, nongt_dim, feat_dim
        B, N, _ = boxes.shape
        self.shape = (B, min(nongt_dim, N), N, feat_dim)          # what the materialised tensor would be


def _to_device_boxes(bb):
    if isinstance(bb, np.ndarray):
        if not torch.cuda.is_available():
            raise _lib.RegatError(-6, "prepare_graph_variables: no CUDA device (there is no CPU path)")
        bb = torch.from_numpy(np.ascontiguousarray(bb, dtype=np.float32)).cuda()
    bb = _rt.need_cuda(bb, "bb")
    if bb.dim() != 3 or bb.shape[-1] != 4:
        raise ValueError(f"bb must be [batch, num_boxes, 4] absolute (x1,y1,x2,y2); got {tuple(bb.shape)}")
    return bb


def tf_extract_position_embedding_from_boxes(bb, nongt_dim, feat_dim=64):
    """position_emb.py:117-151 followed by :96-115, fused: boxes -> [B, M, N, feat_dim]."""
    bb = _to_device_boxes(bb)
    B, N, _ = bb.shape
    M = min(nongt_dim, N)
    out = _rt.empty(B, M, N, feat_dim, device=bb.device)
    wd = _lib.wave_divisors(feat_dim)
    _lib.check(_lib.lib().regat_position_embedding(bb.data_ptr(), B, N, nongt_dim, feat_dim, wd.ctypes.data, out.data_ptr(),
                                                   _rt.stream()))
    return out


def prepare_graph_variables(relation_Type, bb, sem_adj_matrix, spa_adj_matrix, num_objects, nongt_dim, pos_emb_dim,
                            spa_label_num, sem_label_num, lazy=False):
    """Same positional signature as position_emb.py:153-155; the relation type and adjacency arguments are accepted and
    ignored exactly as there.  Returns (pos_emb, None, None)."""
    if lazy:
        return BoxGeometry(_to_device_boxes(bb), nongt_dim, pos_emb_dim), None, None
    return tf_extract_position_embedding_from_boxes(bb, nongt_dim, pos_emb_dim), None, None
