"""The hot-path slice of RelationGraphAttentionNetwork -- mirrors model/rel_graph_net.py:53-62 and the builder :94-106.

The language front-end (w_emb, q_emb, q_att; rel_graph_net.py:41-45,57) is outside this path: its two outputs
q_emb_self_att and q_emb (= call_last) are inputs here."""
import numpy as np
import torch

from ..config import HotPathConfig, param_layout
from .classifier import SimpleClassifier
from .fusion import BUTD
from .relation_encoder import ImplicitRelationEncoder
from .weight_norm import Layer


class ReGATHotPath(Layer):
    def __init__(self, v_relation, joint_emb, classifier):
        self.v_relation = v_relation
        self.joint_emb = joint_emb
        self.classifier = classifier

    def call(self, visual, q_emb_self_att, q_emb, implicit_pos_emb):
        v_emb = self.v_relation(visual, implicit_pos_emb, q_emb_self_att)       # rel_graph_net.py:53
        joint_emb, weights = self.joint_emb(v_emb, q_emb)                       # :58
        return self.classifier(joint_emb) if self.classifier else joint_emb     # :61-64

    # ---- the flat parameter buffer of engine.py uses exactly this variable order
    def load_flat(self, cfg: HotPathConfig, flat):
        arrays = [np.asarray(flat[e.offset:e.offset + e.numel], dtype=np.float32).reshape(e.shape) for e in param_layout(cfg)[0]]
        self.set_weights(arrays)

    def to_flat(self, cfg: HotPathConfig):
        entries, total = param_layout(cfg)
        flat = np.zeros(total, dtype=np.float32)
        for e, a in zip(entries, self.get_weights()):
            flat[e.offset:e.offset + e.numel] = a.ravel()
        return flat


def build_hot_path(cfg: HotPathConfig, dropout=0.2, device="cuda:0"):
    """rel_graph_net.py:94-110 for relation_type == 'implicit', fusion == 'butd'; variables are created immediately
    (the reference creates them on the first call, :113-123)."""
    v_relation = ImplicitRelationEncoder(cfg.v_dim, cfg.q_dim, cfg.rel_dim, cfg.dir_num, cfg.pos_emb_dim, cfg.nongt_dim,
                                         num_heads=cfg.num_heads, num_steps=1, residual_connection=cfg.residual,
                                         label_bias=cfg.label_bias)
    classifier = SimpleClassifier(cfg.q_dim, cfg.q_dim * 2, cfg.num_answers, dropout)
    joint = BUTD(cfg.rel_dim, cfg.q_dim, cfg.q_dim)
    model = ReGATHotPath(v_relation, joint, classifier)
    dev = torch.device(device)
    # one dummy forward to create every variable, like the reference's eval-mode builder
    B, N = 1, max(2, min(cfg.nongt_dim, 4))
    z = lambda *s: torch.zeros(*s, device=dev)
    from .position_emb import BoxGeometry
    boxes = torch.tensor([[[0., 0., 10., 10.]] * N], device=dev)
    model(z(B, N, cfg.v_dim) + 1.0, z(B, cfg.q_dim), z(B, cfg.q_dim), BoxGeometry(boxes, cfg.nongt_dim, cfg.pos_emb_dim))
    return model
