import tensorflow as tf

floor, sqrt = tf.floor, tf.sqrt


def pow(x, y, name=None):
    return tf.convert_to_tensor(x) ** y
