"""Empty stand-in: imported at the top of reference modules, never called on the pose path."""
