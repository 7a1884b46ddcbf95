"""tf.keras.optimizers.experimental.Adamax (keras/optimizers/optimizer_experimental/adamax.py,
TF 2.9-2.10; the import at train.py:15 fixes the class).  Per variable, with t = iterations + 1:

    m <- m + (g - m) * (1 - beta_1)
    u <- max(beta_2 * u, |g|)
    w <- w - (lr * m) / ((1 - beta_1**t) * (u + epsilon))

iterations is incremented once per apply_gradients().  `lr` is a variable with read_value()/assign()
(train.py:70-80).  TEST INFRASTRUCTURE."""
import torch

from ..._core import Variable, _t, floatx


from ..._core import IndexedSlices  # noqa: E402


class Adamax:
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, **_):
        self.lr = Variable.make(torch.tensor(float(learning_rate), dtype=floatx()), name="learning_rate", trainable=False)
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.iterations = 0
        self._slots = {}

    learning_rate = property(lambda self: self.lr)

    def apply_gradients(self, grads_and_vars):
        t = self.iterations + 1
        b1p = self.beta_1 ** t
        lr = self.lr.detach()
        with torch.no_grad():
            for g, w in grads_and_vars:
                if g is None:
                    continue
                if id(w) not in self._slots:
                    self._slots[id(w)] = (torch.zeros_like(w), torch.zeros_like(w))
                m, u = self._slots[id(w)]
                if isinstance(g, IndexedSlices):
                    # keras/optimizers/adamax.py update_step, sparse branch: m decays everywhere and receives the (summed)
                    # occurrences; u decays everywhere, then EVERY occurrence adds max(u_row, |value|) - u_row of the decayed
                    # row it gathered (duplicates add up); the assign_sub covers the whole variable
                    idx, vals = _t(g.indices).long(), _t(g.values).detach()
                    m.add_(-m * (1 - self.beta_1))
                    m.index_add_(0, idx, vals * (1 - self.beta_1))
                    u.mul_(self.beta_2)
                    u_slice = u[idx]
                    u.index_add_(0, idx, torch.maximum(u_slice, vals.abs()) - u_slice)
                    w.sub_((lr * m) / ((1 - b1p) * (u + self.epsilon)))
                    continue
                g = _t(g).detach()
                m.add_((g - m) * (1 - self.beta_1))
                torch.maximum(self.beta_2 * u, g.abs(), out=u)
                w.sub_((lr * m) / ((1 - b1p) * (u + self.epsilon)))
        self.iterations = t

    def slots(self, w):
        return self._slots[id(w)]
