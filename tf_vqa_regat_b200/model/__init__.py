"""Host-side mirror of the reference's model/ package for the hot path: same module names, class names, constructor
arguments, call signatures and variable order (so weights saved in Keras order load unchanged); every number comes from
libregat.so."""
from .classifier import SimpleClassifier  # noqa: F401
from .fc import FullyConnected  # noqa: F401
from .fusion import BUTD  # noqa: F401
from .graph_att_layer import GraphSelfAttentionLayer  # noqa: F401
from .graph_att_net import GraphAttentionNetwork  # noqa: F401
from .position_emb import BoxGeometry, prepare_graph_variables  # noqa: F401
from .rel_graph_net import ReGATHotPath, build_hot_path  # noqa: F401
from .relation_encoder import ExplicitRelationEncoder, ImplicitRelationEncoder, concat_visual_question  # noqa: F401
from .weight_norm import WeightNorm  # noqa: F401
