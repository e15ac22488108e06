"""Package version (the engine tracks the gradslam release whose API it mirrors: 0.1.0)."""
__version__ = "0.1.0"
