"""Empty stand-in: imported at the top of reference modules, never called on the pose path."""
from . import cm, colors, pyplot  # noqa: F401
