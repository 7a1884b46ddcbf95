"""Checkpoint I/O for the hot path's variables, by ORDER like Keras-2 `save_weights` / `load_weights`.

The reference saves with `model.save_weights(path)` / restores with `model.load_weights(path)` (main.py:145,155): an HDF5 file
whose arrays are matched to variables by their order inside each top-level layer, not by name.  h5py (and any HDF5 library) is not available in
this image, so the container here is `.npz` -- the HDF5 container itself is WAIVED (read_keras_h5 / tools/convert_keras_h5.py use
h5py where it exists); the ORDER is the contract and is the same one -- per WeightNorm wrapper `[v, g, bias]`
(weight_norm.py:21-31), layers in attribute-assignment order (rel_graph_net.py:16-21, relation_encoder.py:53-58,
graph_att_net.py:24-36, graph_att_layer.py:25-37, fusion.py:15-20, classifier.py:14-19) -- i.e. `config.param_layout`.
A maintainer with h5py converts a reference checkpoint by listing its datasets in file order and passing them to
`arrays_to_flat`; the question front-end's variables precede the hot path's in the file (w_emb, q_emb, q_att come first,
rel_graph_net.py:16-18) -- `save_model_weights` / `load_model_weights` keep that whole-model order
(tests/golden/reference_variable_order.json is the order recorded from the executed reference).

Arrays are stored under keys "000:<name>", "001:<name>", ... so that the file is self-describing, but loading goes by index
and validates count and shapes the way Keras does (ValueError naming the offending variable)."""
import json
from typing import Iterable, List

import numpy as np

from .config import HotPathConfig, param_layout

FORMAT = "regat-b200-weights/1"


def flat_to_arrays(cfg: HotPathConfig, flat) -> List[np.ndarray]:
    """Flat fp32 buffer (engine.params layout, with alignment padding) -> list of arrays in Keras variable order."""
    flat = np.asarray(flat, dtype=np.float32).reshape(-1)
    entries, total = param_layout(cfg)
    if flat.size != total:
        raise ValueError(f"flat buffer has {flat.size} elements, the layout needs {total}")
    return [flat[e.offset:e.offset + e.numel].reshape(e.shape).copy() for e in entries]


def arrays_to_flat(cfg: HotPathConfig, arrays: Iterable) -> np.ndarray:
    """List of arrays in Keras variable order -> flat fp32 buffer.  Count and shapes must match (Keras: 'You called
    `set_weights(weights)` ... with a weight list of length N, but the layer was expecting M weights')."""
    arrays = list(arrays)
    entries, total = param_layout(cfg)
    if len(arrays) != len(entries):
        raise ValueError(f"weight list of length {len(arrays)}, but the hot path expects {len(entries)} variables")
    flat = np.zeros(total, dtype=np.float32)
    for e, a in zip(entries, arrays):
        a = np.asarray(a, dtype=np.float32)
        if tuple(a.shape) != tuple(e.shape):
            raise ValueError(f"variable {e.name} has shape {tuple(e.shape)}, but the checkpoint holds {tuple(a.shape)}")
        flat[e.offset:e.offset + e.numel] = a.reshape(-1)
    return flat


def save_weights(path: str, cfg: HotPathConfig, flat) -> None:
    """Write the variables of `flat` to `path` (.npz) in Keras variable order."""
    entries, _ = param_layout(cfg)
    arrays = flat_to_arrays(cfg, flat)
    meta = {"format": FORMAT, "config": {k: getattr(cfg, k) for k in cfg.__dataclass_fields__}, "variables": [e.name for e in entries]}
    payload = {f"{i:03d}:{e.name}": a for i, (e, a) in enumerate(zip(entries, arrays))}
    payload["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    with open(path, "wb") as f:
        np.savez(f, **payload)


def _ordered_keys(files):
    """Keys of an .npz in VARIABLE order: by the integer index they start with ("000:<name>", "12", "arr_7" -> 0, 12, 7), never
    lexically -- 'arr_10' must not sort before 'arr_2'.  Keys without an index are refused: order is the whole contract."""
    import re
    out = []
    for k in files:
        if k == "__meta__":
            continue
        m = re.match(r"^(?:arr_)?(\d+)(?::.*)?$", k)
        if not m:
            raise ValueError(f"checkpoint key '{k}' carries no variable index ('<index>:<name>', '<index>' or 'arr_<index>' expected)")
        out.append((int(m.group(1)), k))
    out.sort()
    idx = [i for i, _ in out]
    if idx != list(range(len(idx))):
        raise ValueError(f"checkpoint variable indices are not 0..{len(idx) - 1} without gaps or repeats: {idx[:8]}...")
    return [k for _, k in out]


def load_weights(path: str, cfg: HotPathConfig) -> np.ndarray:
    """Read a file written by save_weights (or any .npz whose keys carry the Keras variable index) -> flat fp32 buffer.  When the
    file has a `__meta__` record its format tag and variable names are checked against this configuration's layout."""
    if str(path).endswith((".h5", ".hdf5")):
        return arrays_to_flat(cfg, read_keras_h5(path)[-len(param_layout(cfg)[0]):])
    with np.load(path) as z:
        arrays = [z[k] for k in _ordered_keys(z.files)]
        if "__meta__" in z.files:
            meta = json.loads(bytes(z["__meta__"]).decode())
            if meta.get("format") != FORMAT:
                raise ValueError(f"checkpoint format '{meta.get('format')}', expected '{FORMAT}'")
            names = [e.name for e in param_layout(cfg)[0]]
            if meta.get("variables") != names:
                bad = next((a, b) for a, b in zip(list(meta.get("variables", [])) + [None], names + [None]) if a != b)
                raise ValueError(f"checkpoint was written for a different variable list: first difference {bad[0]!r} vs {bad[1]!r}")
    return arrays_to_flat(cfg, arrays)


def read_keras_h5(path: str):
    """Arrays of a Keras-2 `model.save_weights('x.h5')` file (main.py:145) in the order `load_weights` consumes them: top-level
    layers in the file's `layer_names` attribute order, within each the `weight_names` attribute order (keras/saving/hdf5_format.py).
    Needs h5py, which is NOT part of this image (no HDF5 library here: the container format is waived, the by-order contract is
    what the rest of this module implements and tests); a maintainer runs tools/convert_keras_h5.py where h5py exists."""
    try:
        import h5py
    except ImportError as ex:
        raise ImportError("reading a Keras .h5 checkpoint needs h5py (not installed in this image); convert it to the by-order .npz "
                          "with tools/convert_keras_h5.py on a machine that has h5py") from ex
    out = []
    with h5py.File(path, "r") as f:
        g = f["model_weights"] if "model_weights" in f else f
        dec = lambda x: x.decode() if isinstance(x, bytes) else x
        for lname in [dec(n) for n in g.attrs["layer_names"]]:
            for wname in [dec(n) for n in g[lname].attrs["weight_names"]]:
                out.append(np.asarray(g[lname][wname]))
    return out


# ---- whole model: question front-end (question.py) + hot path, in the reference's top-level layer order
def save_model_weights(path: str, front_arrays: Iterable, cfg: HotPathConfig, flat) -> None:
    """front_arrays: the front-end's variables in Keras order (QuestionFrontEnd.named().values()); then the hot path's."""
    arrays = [np.asarray(a, dtype=np.float32) for a in front_arrays] + flat_to_arrays(cfg, flat)
    with open(path, "wb") as f:
        np.savez(f, **{f"{i:03d}": a for i, a in enumerate(arrays)})


def load_model_weights(path: str, cfg: HotPathConfig, front_shapes):
    """-> (front-end arrays in order, hot-path flat buffer).  front_shapes: [(name, shape)] the front-end expects
    (question.question_layout); count and shapes are validated like load_weights."""
    with np.load(path) as z:
        arrays = [z[k] for k in _ordered_keys(z.files)]
    front_shapes = list(front_shapes)
    n = len(front_shapes)
    if len(arrays) < n:
        raise ValueError(f"weight list of length {len(arrays)}, but the front-end alone expects {n} variables")
    for (name, shape), a in zip(front_shapes, arrays[:n]):
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"variable {name} has shape {tuple(shape)}, but the checkpoint holds {tuple(a.shape)}")
    return arrays[:n], arrays_to_flat(cfg, arrays[n:])
