"""Odometry provider interface (mirror of gradslam/odometry/base.py:6-19)."""
from abc import ABC, abstractmethod

__all__ = ["OdometryProvider"]


class OdometryProvider(ABC):
    """A provider turns two observations into per-element rigid transforms of shape (B, 1, 4, 4)."""

    @abstractmethod
    def __init__(self, *params):
        pass

    @abstractmethod
    def provide(self, *args, **kwargs):
        raise NotImplementedError
