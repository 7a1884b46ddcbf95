"""Host-side learning-rate schedule (tf_vqa_regat_b200/schedule.py) against the values the reference's own train() set per
epoch when executed over the TensorFlow stand-in (tests/golden/refexec_lr_schedule.json, oracle/make_golden_ref_schedule.py)."""
import json
import os

import numpy as np

from tf_vqa_regat_b200.schedule import lr_schedule


def test_lr_schedule_matches_the_executed_reference(golden_dir):
    runs = json.load(open(os.path.join(golden_dir, "refexec_lr_schedule.json")))
    assert len(runs) == 3
    for r in runs:
        got = lr_schedule(r["base_lr"], r["epochs"], r["lr_decay_step"], r["lr_decay_rate"])
        np.testing.assert_allclose(got, r["lr_per_epoch"], rtol=1e-12)
        # one optimizer step per batch, never reset between epochs (Adamax bias correction keeps counting)
        assert r["optimizer_iterations_at_epoch_start"] == list(range(r["epochs"]))


def test_schedule_quirks():
    s = lr_schedule(9e-4, 20, 2, 0.75)                        # the shipped config (butd_vqa.json)
    assert s[:5] == [9e-4, 9e-4, 1.2 * 9e-4, 1.3 * 9e-4, 1.4 * 9e-4]
    assert s[5] == s[6] and abs(s[5] - 1.4 * 9e-4 * 0.75) < 1e-18        # decay starts from the warm-up peak, holds between decay epochs
    assert abs(s[19] - 1.4 * 9e-4 * 0.75 ** 8) < 1e-15
