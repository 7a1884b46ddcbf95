"""tf.keras.preprocessing.sequence.pad_sequences (keras/utils/data_utils.py): every sequence is cut to `maxlen` entries
(truncating='pre' drops from the front, 'post' from the back) and written into a [num, maxlen, *sample_shape] array filled
with `value`, at the front of its row for padding='post' and at the back for padding='pre'; empty sequences leave their row
untouched; dtype conversion by NumPy.  TEST INFRASTRUCTURE."""
import numpy as np


def pad_sequences(sequences, maxlen=None, dtype="int32", padding="pre", truncating="pre", value=0.0):
    lengths = [len(s) for s in sequences]
    if maxlen is None:
        maxlen = max(lengths) if lengths else 0
    sample_shape = ()
    for s in sequences:
        if len(s) > 0:
            sample_shape = np.asarray(s).shape[1:]
            break
    x = np.full((len(sequences), maxlen) + tuple(sample_shape), value, dtype=dtype)
    for i, s in enumerate(sequences):
        if not len(s):
            continue
        trunc = s[-maxlen:] if truncating == "pre" else s[:maxlen]
        trunc = np.asarray(trunc, dtype=dtype)
        if padding == "post":
            x[i, :len(trunc)] = trunc
        else:
            x[i, -len(trunc):] = trunc
    return x
