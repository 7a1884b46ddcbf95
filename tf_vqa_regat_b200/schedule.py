"""Learning-rate schedule of the reference's training loop (train.py:53-83), host side: the engine takes the step's learning
rate as an argument (regat_engine_update / regat_engine_train_step), this module says what it is for a given epoch.

Reference behaviour, quirks included:
  * epochs 0..4: lr = base * [1, 1, 1.2, 1.3, 1.4][epoch]                      (train.py:54,70-75)
  * epochs 5, 5+step, 5+2*step, ... (< epochs): lr = previous lr * decay_rate   (train.py:55,77-81) -- the decay starts from
    1.4 * base, `lr_decay_start` (main.py:21) is never read, and epochs between two decay epochs keep the last value;
  * the optimizer's step counter (Adamax bias correction) keeps running across epochs (train.py:48 creates it once).
Pinned by tests/golden/refexec_lr_schedule.json (the reference's own train() run for several epochs over the TensorFlow
stand-in, oracle/make_golden_ref_schedule.py)."""
from typing import List

WARMUP = (1.0, 1.0, 1.2, 1.3, 1.4)


def lr_schedule(base_lr: float, epochs: int, lr_decay_step: int = 2, lr_decay_rate: float = 0.25) -> List[float]:
    """Learning rate in force during each epoch 0..epochs-1."""
    out, lr = [], base_lr
    decay_epochs = set(range(5, epochs, lr_decay_step))
    for epoch in range(epochs):
        if epoch < len(WARMUP):
            lr = WARMUP[epoch] * base_lr
        elif epoch in decay_epochs:
            lr = lr * lr_decay_rate
        out.append(lr)
    return out
