"""tf.layers (only ``dense`` is on the reference's pose path: nets/attention_module.py:37-50, 89-101)."""
from . import _dense as dense  # noqa: F401
