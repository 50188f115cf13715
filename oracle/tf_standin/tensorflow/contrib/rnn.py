import tensorflow as tf

stack_bidirectional_dynamic_rnn = tf.stack_bidirectional_dynamic_rnn
