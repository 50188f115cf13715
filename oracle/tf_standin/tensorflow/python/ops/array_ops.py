import numpy as np

import tensorflow as tf

where = tf.where
shape = tf.shape


def zeros_like(tensor, dtype=None, name=None):
    return tf.Tensor(lambda f, c: np.zeros_like(tensor.eval(f, c)))


def identity(x, name=None):
    return tf.convert_to_tensor(x)
