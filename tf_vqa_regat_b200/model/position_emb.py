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


def build_graph(bbox, spatial, label_num=11):
    """Spatial relation labels of one image -- position_emb.py:23-90, vectorised (NumPy, host: it is dataset preparation).

    bbox [n,4] absolute (x1,y1,x2,y2); spatial [n,6] normalised, columns 4/5 = box width/height over image width/height.
    Returns the [n,n] label matrix of the reference: 12 on the diagonal of real boxes, 1 = j inside i, 2 = j covers i,
    3 = IoU >= 0.5, 4..11 = direction octant of the centre offset when the centres are closer than half the image diagonal,
    0 = no edge; boxes whose coordinates sum to 0 (padding) have no edges.  (`label_num` is accepted and unused, as there.)"""
    bbox = np.asarray(bbox, dtype=np.float64)
    n = bbox.shape[0]
    x1, y1, x2, y2 = (bbox[:, k] for k in range(4))
    w, h = x2 - x1 + 1.0, y2 - y1 + 1.0
    image_h, image_w = h[0] / spatial[0, -1], w[0] / spatial[0, -2]
    cx, cy = 0.5 * (x1 + x2), 0.5 * (y1 + y2)
    half_diag = 0.5 * np.sqrt(image_h ** 2 + image_w ** 2)
    real = bbox.sum(1) != 0
    I, J = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    pair = real[I] & real[J] & (I < J)                               # the reference fills (i,j) and (j,i) from the i < j visit
    inside = (x1[I] < x1[J]) & (x2[I] > x2[J]) & (y1[I] < y1[J]) & (y2[I] > y2[J])       # j inside i
    cover = (x1[J] < x1[I]) & (x2[J] > x2[I]) & (y1[J] < y1[I]) & (y2[J] > y2[I])        # j covers i
    iw = np.maximum(0.0, np.minimum(x2[I], x2[J]) - np.maximum(x1[I], x1[J]) + 1)
    ih = np.maximum(0.0, np.minimum(y2[I], y2[J]) - np.maximum(y1[I], y1[J]) + 1)
    inter = iw * ih
    iou = inter / (w[I] * h[I] + w[J] * h[J] - inter)
    yd, xd = cy[I] - cy[J], cx[I] - cx[J]
    diag = np.sqrt(yd ** 2 + xd ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        sin_ij, cos_ij = yd / diag, xd / diag
        q1, q4 = (sin_ij >= 0) & (cos_ij >= 0), (sin_ij < 0) & (cos_ij >= 0)
        q2 = (sin_ij >= 0) & (cos_ij < 0)
        asin, acos_c, acos_s = np.arcsin(np.clip(sin_ij, -1, 1)), np.arccos(np.clip(cos_ij, -1, 1)), np.arccos(np.clip(sin_ij, -1, 1))
        two_pi = 2 * np.pi
        label_i = np.where(q1, asin, np.where(q4, asin + two_pi, np.where(q2, acos_c, -acos_s + two_pi)))
        label_j = np.where(q1 | q2, two_pi - label_i, label_i - np.pi)
        oct_i = np.ceil(label_i / (np.pi / 4)).astype(np.int64) + 3
        oct_j = np.ceil(label_j / (np.pi / 4)).astype(np.int64) + 3
    rest = pair & ~inside & ~cover
    near = rest & (iou < 0.5) & (diag < half_diag)
    upper = np.where(pair & inside, 1, np.where(pair & cover & ~inside, 2, np.where(rest & (iou >= 0.5), 3, np.where(near, oct_i, 0))))
    lower = np.where(pair & inside, 2, np.where(pair & cover & ~inside, 1, np.where(rest & (iou >= 0.5), 3, np.where(near, oct_j, 0))))
    adj = (upper + lower.T).astype(np.float64)
    adj[np.arange(n)[real], np.arange(n)[real]] = 12
    return adj


def one_hot_adjacency(adj_labels, label_num):
    """[.., n, n] integer labels (0 = no edge) -> [.., n, n, label_num] one-hot on label - 1, the [B,N,N,label_num] adjacency that
    ExplicitRelationEncoder.call consumes (relation_encoder.py:124; the reference's own tf_broadcast_adj_matrix is an empty stub,
    position_emb.py:92-93).  Labels above label_num (the diagonal's 12 when label_num = 11) are dropped, as in the original ReGAT."""
    a = np.asarray(adj_labels).astype(np.int64)
    out = np.zeros(a.shape + (label_num,), dtype=np.float32)
    ok = (a > 0) & (a <= label_num)
    idx = np.nonzero(ok)
    out[idx + (a[ok] - 1,)] = 1.0
    return out
