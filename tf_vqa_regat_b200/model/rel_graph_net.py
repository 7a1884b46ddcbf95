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

    # ---- training / compiled execution (train.py:103-113 tapes model(...); here the composed model runs through the engine)
    def compile(self, cfg: HotPathConfig, max_batch: int, max_rois: int, dtype: str = "bf16"):
        """Routes this model through regat_engine_* with ITS OWN variables as the engine's parameter buffer: after the call the
        layers' v / g / bias tensors are views of the engine's flat buffer, so layer-by-layer calls, `train_step`, `get_weights`
        and `set_weights` all see one set of weights.  dtype "bf16" (tcgen05 GEMMs) or "fp32" (parity kernels)."""
        from ..engine import HotPathEngine
        dev = self.weights[0][1].device
        eng = HotPathEngine(cfg, max_batch, max_rois, dtype=dtype, device=dev)
        eng.load_params(self.to_flat(cfg))
        entries = eng.entries
        i = 0
        for wn in _weight_norm_layers(self):
            n = 3 if wn.bias is not None else 2
            e = entries[i:i + n]
            assert [x.kind for x in e] == ["v", "g", "b"][:n] and tuple(e[0].shape) == tuple(wn.v.shape), (e[0].name, wn.v.shape)
            wn.rebind(eng.params, e[0].offset, e[1].offset, e[2].offset if n == 3 else -1)
            i += n
        assert i == len(entries), "the model's WeightNorm layers do not cover the engine's parameter layout"
        eng.params_changed()
        self._engine, self._cfg = eng, cfg
        return self

    def _compiled(self):
        eng = getattr(self, "_engine", None)
        if eng is None:
            raise RuntimeError("call model.compile(cfg, max_batch, max_rois, dtype) first")
        return eng

    @staticmethod
    def _boxes(implicit_pos_emb):
        from .position_emb import BoxGeometry
        if not isinstance(implicit_pos_emb, BoxGeometry):
            raise TypeError("the compiled path takes the lazy BoxGeometry handle (prepare_graph_variables(..., lazy=True)): "
                            "it rebuilds the geometry on chip and never reads a materialised pos_emb")
        return implicit_pos_emb.boxes

    def predict(self, visual, q_emb_self_att, q_emb, implicit_pos_emb):
        """Same arguments and result as call(), executed by the engine in the compiled dtype (train.py:136-177 evaluate)."""
        return self._compiled().forward(visual, self._boxes(implicit_pos_emb), q_emb_self_att, q_emb)

    def train_step(self, visual, q_emb_self_att, q_emb, implicit_pos_emb, target, lr, step=None, want_dq=False):
        """train.py:103-113 for this model's variables: loss = mean BCE * num_answers, gradients, per-tensor clip_by_norm, Adamax.
        Returns the device tensor (loss, batch score); with want_dq also (dq_emb_self_att, dq_emb) for the language model's tape."""
        eng = self._compiled()
        boxes = self._boxes(implicit_pos_emb)
        if not want_dq:
            return eng.train_step(visual, boxes, q_emb_self_att, q_emb, target, lr, step)
        out = eng.fwd_bwd(visual, boxes, q_emb_self_att, q_emb, target, want_dq=True)
        eng.update(lr, step)
        return eng._loss, (out["dq_att"], out["dq_last"])

    def set_weights(self, arrays):
        super().set_weights(arrays)
        if getattr(self, "_engine", None) is not None:
            self._engine.params_changed()                 # the engine's derived state (alpha, bf16 kernels) follows the write

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


def _weight_norm_layers(layer):
    """WeightNorm layers in variable order (own weights first, then children in attribute order: Layer.weights)."""
    from .weight_norm import WeightNorm
    out = [layer] if isinstance(layer, WeightNorm) else []
    if not isinstance(layer, WeightNorm):
        for s in layer.sublayers():
            out.extend(_weight_norm_layers(s))
    return out


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
