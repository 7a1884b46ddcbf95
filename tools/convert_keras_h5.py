#!/usr/bin/env python
"""Convert a checkpoint written by the reference (`model.save_weights('...pretrained_model.h5')`, main.py:145) into the by-order
.npz containers of tf_vqa_regat_b200/checkpoint.py.  Needs h5py -- run it where the reference runs; this repository's build image
has no HDF5 library, so the script is exercised there only up to its argument handling (tests/test_checkpoint.py).

    python tools/convert_keras_h5.py pretrained_model.h5 out_prefix [--n-token N --emb-dim 300 --op c]

Writes <out_prefix>.model.npz (whole model, reference top-level order: w_emb, q_emb, q_att, v_relation, joint_emb, classifier --
rel_graph_net.py:16-21) and <out_prefix>.hotpath.npz (the hot path's variables only, with names and config: load with
HotPathEngine.load_weights)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("h5")
    ap.add_argument("out_prefix")
    for k in ("v_dim", "q_dim", "rel_dim", "num_heads", "nongt_dim", "dir_num", "num_answers"):
        ap.add_argument("--" + k.replace("_", "-"), type=int, default=None)
    ap.add_argument("--label-bias", action="store_true")
    args = ap.parse_args(argv)
    import numpy as np
    from tf_vqa_regat_b200 import checkpoint
    from tf_vqa_regat_b200.config import HotPathConfig, param_layout
    kw = {k: getattr(args, k) for k in ("v_dim", "q_dim", "rel_dim", "num_heads", "nongt_dim", "dir_num", "num_answers") if getattr(args, k) is not None}
    cfg = HotPathConfig(label_bias=args.label_bias, **kw)
    arrays = checkpoint.read_keras_h5(args.h5)          # raises ImportError with instructions when h5py is missing
    n_hot = len(param_layout(cfg)[0])
    if len(arrays) < n_hot:
        raise SystemExit(f"{args.h5} holds {len(arrays)} arrays, the hot path alone has {n_hot} variables: wrong configuration?")
    flat = checkpoint.arrays_to_flat(cfg, arrays[-n_hot:])       # validates every shape, names the offending variable
    checkpoint.save_weights(args.out_prefix + ".hotpath.npz", cfg, flat)
    with open(args.out_prefix + ".model.npz", "wb") as f:
        np.savez(f, **{f"{i:03d}": np.asarray(a, dtype=np.float32) for i, a in enumerate(arrays)})
    print(f"{len(arrays)} variables: {len(arrays) - n_hot} front-end + {n_hot} hot-path -> {args.out_prefix}.model.npz, {args.out_prefix}.hotpath.npz")


if __name__ == "__main__":
    main()
