import numpy as np

import tensorflow as tf


def smart_cond(pred, true_fn=None, false_fn=None, name=None):
    """tf.cond on a tensor predicate: the branch is chosen when the graph runs.  Clair's dropout_selu takes the identity
    branch whenever phase_placeholder is False (clair/selu.py:72-74); its training branch is outside the forward path."""
    def run(f, c):
        if np.asarray(tf.convert_to_tensor(pred).eval(f, c)).item():
            raise NotImplementedError("training-mode dropout is outside the forward path")
        return false_fn().eval(f, c)
    return tf.Tensor(run)
