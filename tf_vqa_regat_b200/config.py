"""Hot-path hyper-parameters and the flat parameter layout.

Values default to the reference's config/butd_vqa.json:1-29.  The parameter ORDER is
the Keras-2 variable order of the reference layers: within a layer its
own weights first, then children in attribute-assignment order, and every WeightNorm
wrapper contributes [v, g, bias] (weight_norm.py:17-33).  The C engine
(csrc/engine.cu: build_layout) computes the same offsets; tests compare the two.
"""
from dataclasses import dataclass, field
from typing import List, Tuple

ALIGN = 64  # elements; every tensor in a flat buffer starts on a 256-byte boundary


@dataclass(frozen=True)
class HotPathConfig:
    v_dim: int = 2048          # region feature dim          (train.ipynb:89-90)
    q_dim: int = 768           # num_hid                     (butd_vqa.json:20)
    rel_dim: int = 1024        # relation_dim                (butd_vqa.json:23)
    num_heads: int = 16        # butd_vqa.json:12
    pos_emb_dim: int = 64      # imp_pos_emb_dim             (butd_vqa.json:11)
    nongt_dim: int = 20        # butd_vqa.json:19
    dir_num: int = 2           # butd_vqa.json:8
    num_answers: int = 3129    # train.ipynb:90
    label_bias: bool = False   # butd_vqa.json:26
    residual: bool = True      # butd_vqa.json:25
    grad_clip: float = 0.25    # main.py:24
    beta1: float = 0.9         # train.py:48
    beta2: float = 0.999
    eps: float = 1e-8

    @property
    def head_dim(self) -> int:
        return self.rel_dim // self.num_heads

    @property
    def hid(self) -> int:      # BUTD hidden = num_hid (rel_graph_net.py:106)
        return self.q_dim

    def m_keys(self, n_rois: int) -> int:
        """graph_att_layer.py:42 -- keys are the first min(nongt_dim, N) objects."""
        return min(self.nongt_dim, n_rois)


@dataclass(frozen=True)
class ParamEntry:
    name: str
    shape: Tuple[int, ...]
    offset: int                # element offset into the flat fp32 buffer
    kind: str                  # 'v' | 'g' | 'b'
    layer: int                 # index of the weight-normed layer this belongs to

    @property
    def numel(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


# (layer name, kernel shape builder, has-bias builder) in reference variable order
def layer_table(cfg: HotPathConfig) -> List[Tuple[str, Tuple[int, ...], bool]]:
    V, Q, D, H, E, A, Hd = (cfg.v_dim, cfg.q_dim, cfg.rel_dim, cfg.num_heads,
                            cfg.pos_emb_dim, cfg.num_answers, cfg.hid)
    t: List[Tuple[str, Tuple[int, ...], bool]] = []
    if V != D:                                               # relation_encoder.py:52
        t.append(("v_relation.v2out", (V, D), True))
    t.append(("v_relation.implicit_relation.self_weights", (D + Q, D), True))
    t.append(("v_relation.implicit_relation.bias", (1, 1), cfg.label_bias))
    for d in range(cfg.dir_num):
        p = f"v_relation.implicit_relation.neighbor_net.{d}"
        t.append((p + ".pair_pos_fc", (E, H), True))
        t.append((p + ".query", (D, D), True))
        t.append((p + ".key", (D, D), True))
        t.append((p + ".linear_out_", (1, 1, D, D), True))
    t.append(("joint_emb.v2attention", (D, Hd), True))
    t.append(("joint_emb.q2attention", (Q, Hd), True))
    t.append(("joint_emb.linear", (Hd, 1), True))
    t.append(("joint_emb.visual_embed", (D, Hd), True))
    t.append(("joint_emb.question_embed", (Q, Hd), True))
    t.append(("classifier.layers.0", (Hd, 2 * Hd), True))
    t.append(("classifier.layers.3", (2 * Hd, A), True))
    return t


def param_layout(cfg: HotPathConfig) -> Tuple[List[ParamEntry], int]:
    """Flat layout: for each layer v, g, [bias]; each tensor 256-byte aligned.
    Returns (entries, total elements incl. padding)."""
    entries: List[ParamEntry] = []
    off = 0
    for li, (name, kshape, has_bias) in enumerate(layer_table(cfg)):
        n = 1
        for s in kshape:
            n *= s
        entries.append(ParamEntry(name + "/v", kshape, off, "v", li)); off = _round_up(off + n, ALIGN)
        entries.append(ParamEntry(name + "/g", (), off, "g", li)); off = _round_up(off + 1, ALIGN)
        if has_bias:
            entries.append(ParamEntry(name + "/bias", (kshape[-1],), off, "b", li))
            off = _round_up(off + kshape[-1], ALIGN)
    return entries, off


def num_trainable(cfg: HotPathConfig) -> int:
    return sum(e.numel for e in param_layout(cfg)[0])
