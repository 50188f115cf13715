import tensorflow as tf

name_scope = tf.name_scope
convert_to_tensor = tf.convert_to_tensor
