"""tf.contrib: slim and layers, as far as the reference's pose graph uses them."""
from . import layers, slim  # noqa: F401
