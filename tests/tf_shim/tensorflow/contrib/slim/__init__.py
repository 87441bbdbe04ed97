"""tf.contrib.slim: arg_scope, conv2d, batch_norm, dropout, l2_regularizer (nets/posenn.py:203-240).

Defaults are TF-Slim 1.13's: conv2d(stride=1, padding='SAME', rate=1, activation_fn=relu, normalizer_fn=None,
weights_initializer=xavier, biases_initializer=zeros); batch_norm(decay=.999, center=True, scale=False,
epsilon=.001, is_training=True).
"""
import contextlib

import torch

import tensorflow as tf

_END_POINTS = {}


def _arg_defaults(fn):
    out = {}
    for level in tf._S.arg_scopes:
        out.update(level.get(fn.__name__, {}))
    return out


def add_arg_scope(fn):
    def wrapped(*args, **kwargs):
        merged = _arg_defaults(fn)
        merged.update(kwargs)
        return fn(*args, **merged)
    wrapped.__name__ = fn.__name__
    wrapped.__wrapped__ = fn
    return wrapped


@contextlib.contextmanager
def arg_scope(list_ops_or_scope, **kwargs):
    level = {op.__name__: dict(kwargs) for op in list_ops_or_scope}
    tf._S.arg_scopes.append(level)
    try:
        yield level
    finally:
        tf._S.arg_scopes.pop()


def l2_regularizer(scale, scope=None):
    return lambda w: None


@add_arg_scope
def batch_norm(inputs, decay=0.999, center=True, scale=False, epsilon=0.001, activation_fn=None,
               is_training=True, trainable=True, scope=None, outputs_collections=None, **kw):
    x = tf._raw(inputs)
    c = x.shape[-1]
    with tf.variable_scope(scope or "BatchNorm"):
        beta = tf.get_variable("beta", [c], initializer=tf.zeros_initializer(), trainable=trainable).t if center else 0.0
        gamma = tf.get_variable("gamma", [c], initializer=tf.ones_initializer(), trainable=trainable).t if scale else 1.0
        mm = tf.get_variable("moving_mean", [c], initializer=tf.zeros_initializer(), trainable=False).t
        mv = tf.get_variable("moving_variance", [c], initializer=tf.ones_initializer(), trainable=False).t
        if is_training:                                   # batch statistics, biased variance (fused_batch_norm)
            red = tuple(range(x.dim() - 1))
            mean = x.mean(dim=red)
            var = ((x - mean) ** 2).mean(dim=red)
        else:
            mean, var = mm, mv
        out = tf.Tensor((x - mean) * torch.rsqrt(var + epsilon) * gamma + beta)
        if activation_fn is not None:
            out = activation_fn(out)
    return out


@add_arg_scope
def conv2d(inputs, num_outputs, kernel_size, stride=1, padding="SAME", data_format=None, rate=1,
           activation_fn=tf.nn.relu, normalizer_fn=None, normalizer_params=None,
           weights_initializer=None, weights_regularizer=None, biases_initializer=tf.zeros_initializer(),
           biases_regularizer=None, reuse=None, variables_collections=None, outputs_collections=None,
           trainable=True, scope=None):
    x = tf._raw(inputs)
    kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    st = (stride, stride) if isinstance(stride, int) else tuple(stride)
    rt = (rate, rate) if isinstance(rate, int) else tuple(rate)
    with tf.variable_scope(scope or "Conv", reuse=reuse):
        w = tf.get_variable("weights", [kh, kw, x.shape[-1], int(num_outputs)],
                            initializer=weights_initializer or tf.glorot_uniform_initializer(), trainable=trainable)
        tf._record("conv2d_in", tf.Tensor(x))
        out = tf.Tensor(tf._conv2d(x, w, st, rt, padding))
        if normalizer_fn is not None:
            out = normalizer_fn(out, **(normalizer_params or {}))
        elif biases_initializer is not None:
            b = tf.get_variable("biases", [int(num_outputs)], initializer=biases_initializer, trainable=trainable)
            out = tf.Tensor(out.t + b.t)
        if activation_fn is not None:
            out = activation_fn(out)
        tf._record("conv2d", out)
        if outputs_collections:
            _END_POINTS.setdefault(outputs_collections, {})[tf._scope_prefix()] = out
    return out


@add_arg_scope
def conv2d_transpose(*a, **k):
    raise NotImplementedError("not on the pose path")


@add_arg_scope
def dropout(inputs, keep_prob=0.5, is_training=True, scope=None, **kw):
    assert not is_training, "the inference graph calls dropout with is_training=False (identity)"
    return inputs


@add_arg_scope
def max_pool2d(*a, **k):
    raise NotImplementedError("not on the pose path")


@add_arg_scope
def fully_connected(*a, **k):
    raise NotImplementedError("not on the pose path")
