"""The learning rate the reference's own train() (train.py:41-134) sets in every epoch, recorded by running it for several
epochs over oracle/tf_shim on a tiny model with one batch per epoch.  Build container only:

    python -m oracle.make_golden_ref_schedule        # rewrites tests/golden/refexec_lr_schedule.json
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

from .make_golden_ref import GOLD, TINY, _import_reference, build_model


def main():
    mods = _import_reference()
    tf, ref_train = mods[0], mods[1]
    from tf_vqa_regat_b200 import synthetic as syn
    from tf_vqa_regat_b200.config import HotPathConfig
    cfg = HotPathConfig(**TINY)
    B, N = 2, 6
    batch = syn.make_inputs(cfg, B, N, seed=1000, adaptive=False)
    out = []
    for base_lr, epochs, step, rate in ((9e-4, 12, 2, 0.75), (1e-3, 9, 3, 0.25), (1e-3, 4, 2, 0.25)):
        tf.keras.backend.set_floatx("float64")
        model = build_model(tf, mods, cfg)
        seen = []

        class Loader:
            relation_type = "implicit"
            data_loader_len, num_total_data = 1, B

            def generator(self):
                seen.append((float(model.optimizer.lr.numpy()), int(model.optimizer.iterations)))    # lr in force for this epoch
                yield (batch["features"], None, (tf.constant(batch["q_att"]), tf.constant(batch["q_last"])), batch["boxes"],
                       np.zeros((B, 1)), np.zeros((B, 1)), tf.convert_to_tensor(batch["target"]))

        class EvalLoader(Loader):
            def generator(self):
                yield (batch["features"], None, (tf.constant(batch["q_att"]), tf.constant(batch["q_last"])), batch["boxes"],
                       np.zeros((B, 1)), np.zeros((B, 1)), tf.convert_to_tensor(batch["target"]))

        with tempfile.TemporaryDirectory() as tmp:
            args = types.SimpleNamespace(base_lr=base_lr, epochs=epochs, lr_decay_step=step, lr_decay_rate=rate, grad_clip=0.25,
                                         output=tmp + "/", relation_type="implicit", nongt_dim=cfg.nongt_dim, imp_pos_emb_dim=64,
                                         spa_label_num=11, sem_label_num=15, print_freq=500)
            stdout, sys.stdout = sys.stdout, open(os.devnull, "w")
            try:
                ref_train.train(model, Loader(), EvalLoader(), args)
            finally:
                sys.stdout.close(); sys.stdout = stdout
        out.append({"base_lr": base_lr, "epochs": epochs, "lr_decay_step": step, "lr_decay_rate": rate,
                    "lr_per_epoch": [s[0] for s in seen], "optimizer_iterations_at_epoch_start": [s[1] for s in seen]})
        print(base_lr, epochs, step, rate, [round(s[0] / base_lr, 4) for s in seen])
    with open(os.path.join(GOLD, "refexec_lr_schedule.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
