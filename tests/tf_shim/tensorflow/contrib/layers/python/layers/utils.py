"""tensorflow.contrib.layers.python.layers.utils (nets/posenn.py:6)."""
import tensorflow as tf


def convert_collection_to_dict(collection, clear_collection=False):
    return dict(tf.contrib.slim._END_POINTS.get(collection, {}))
