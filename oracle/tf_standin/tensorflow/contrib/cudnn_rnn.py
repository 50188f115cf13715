import tensorflow as tf

CudnnCompatibleLSTMCell = tf.CudnnCompatibleLSTMCell


def CudnnLSTM(*a, **kw):
    raise NotImplementedError("the cuDNN branch needs a GPU; device_lib reports none here")
