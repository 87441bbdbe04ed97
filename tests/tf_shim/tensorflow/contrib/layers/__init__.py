"""tf.contrib.layers: the initialisers named by the reference (values matter only for variables that are not fed)."""
import math

import torch

import tensorflow as tf


def variance_scaling_initializer(factor=2.0, mode="FAN_IN", uniform=False, seed=None, dtype=None):
    """Truncated normal, stddev sqrt(1.3 * factor / fan_in) (TF 1.13 contrib/layers/python/layers/initializers.py)."""
    def init(shape, dt=tf.float32):
        fan_in, fan_out = tf._fans(list(shape))
        n = {"FAN_IN": fan_in, "FAN_OUT": fan_out, "FAN_AVG": (fan_in + fan_out) / 2.0}[mode]
        std = math.sqrt(1.3 * factor / n)
        v = torch.randn(list(shape), generator=tf._S.rng, dtype=torch.float64).clamp(-2, 2) * std
        return v.to(dt.torch)
    return init


def xavier_initializer(uniform=True, seed=None, dtype=None):
    return tf.glorot_uniform_initializer()


from .python.layers import utils  # noqa: E402,F401
