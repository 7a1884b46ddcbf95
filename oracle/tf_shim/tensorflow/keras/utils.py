"""tf.keras.utils stand-in: Sequence is only a base class for the reference's dataset (dataset.py:18,139)."""


class Sequence:
    pass
