"""Stage-1 oracle: pairwise box geometry -> sinusoidal position embedding.

Restates /root/reference/model/position_emb.py:96-160 (NumPy there too).  The
floating-point operation ORDER of the reference is kept on purpose: the sin/cos
arguments reach |x| ~ 690, where one fp32 ulp of the argument is 6e-5, so a
re-associated formula is visibly different from the reference.

TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import numpy as np

CLAMP = 1e-3  # position_emb.py:131


def box_terms(bbox):
    """w, h, cx, cy per box, each [B,N,1] (position_emb.py:119-125)."""
    x1, y1, x2, y2 = (bbox[..., k:k + 1] for k in range(4))
    w = x2 - x1 + 1.0
    h = y2 - y1 + 1.0
    cx = 0.5 * (x1 + x2)
    cy = 0.5 * (y1 + y2)
    return w, h, cx, cy


def extract_position_matrix(bbox, nongt_dim=36):
    """[B,N,4] boxes -> [B,M,N,4] log-geometry, M=min(nongt_dim,N).

    Entry [b,i,j] = (log max(|cx_i-cx_j|/w_i,1e-3), log max(|cy_i-cy_j|/h_i,1e-3),
                     log(w_i/w_j), log(h_i/h_j));  rows i are then SLICED to the
    first nongt_dim (position_emb.py:127-149).
    """
    w, h, cx, cy = box_terms(bbox)
    t = lambda a: np.swapaxes(a, 1, 2)

    def clamped_log_ratio(c, size):
        d = np.abs((c - t(c)) / size)            # [b,i,j] = |c_i - c_j| / size_i
        d = np.where(d < CLAMP, np.asarray(CLAMP, dtype=d.dtype), d)
        return np.log(d)

    feats = [clamped_log_ratio(cx, w), clamped_log_ratio(cy, h),
             np.log(w / t(w)), np.log(h / t(h))]
    return np.stack([f[:, :nongt_dim] for f in feats], axis=3)


def wave_divisors(feat_dim, wave_length=1000, dtype=np.float32):
    """1000^(8k/feat_dim), k=0..feat_dim/8-1, computed in fp32 like position_emb.py:98-100."""
    k = np.arange(0, feat_dim / 8, dtype=np.float32)
    return np.power(np.full((1,), wave_length, dtype=np.float32), (8.0 / feat_dim) * k).astype(dtype)


def extract_position_embedding(position_mat, feat_dim, wave_length=1000):
    """[B,M,N,4] -> [B,M,N,feat_dim]; feature c*(feat_dim/4)+k = sin, +feat_dim/8 = cos
    of 100*P_c / 1000^(8k/feat_dim)   (position_emb.py:96-115)."""
    div = wave_divisors(feat_dim, wave_length)                       # fp32, as the reference
    arg = (100.0 * position_mat)[..., None] / div                    # [B,M,N,4,F/8]
    emb = np.concatenate([np.sin(arg), np.cos(arg)], axis=-1)        # [B,M,N,4,F/4]
    return emb.reshape(emb.shape[0], emb.shape[1], emb.shape[2], feat_dim)


def prepare_graph_variables(relation_type, bb, sem_adj_matrix, spa_adj_matrix,
                            num_objects, nongt_dim, pos_emb_dim, spa_label_num,
                            sem_label_num):
    """Same signature and return as position_emb.py:153-160: (pos_emb, None, None)."""
    pos_mat = extract_position_matrix(bb, nongt_dim=nongt_dim)
    return extract_position_embedding(pos_mat, feat_dim=pos_emb_dim), None, None


def scrambled_pair_index(n_rois, nongt_dim):
    """Which box pair (i', j') the attention layer's bias[i, j] really uses.

    pos_emb is [B,M,N,E] (rows sliced) but graph_att_layer.py:74,81 raw-reshapes the
    flattened M*N axis as [N,M]; so bias[i,j] reads flat index f=i*M+j -> (f//N, f%N).
    Returns two int arrays of shape [N,M].  Identity when M == N.
    """
    m = min(nongt_dim, n_rois)
    f = np.arange(n_rois * m).reshape(n_rois, m)
    return f // n_rois, f % n_rois
