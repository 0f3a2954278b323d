"""TensorFlow twin, same API as ``modulation_functions/diffusion_modulator_tf.py:3-9`` (TF optional)."""

import tensorflow as tf


def diffusion_modulator_tf(length: tf.Tensor, beta: tf.Tensor) -> tf.Tensor:
    length = tf.cast(length, tf.float64)
    beta = tf.cast(beta, tf.float64)
    return tf.pow(-beta, length) / (tf.pow(tf.constant(2.0, tf.float64), length) * tf.exp(tf.math.lgamma(length + 1.0)))
