from .base import OdometryProvider

__all__ = ["ICPOdometryProvider"]


class ICPOdometryProvider(OdometryProvider):
    def __init__(self, numiters=20, damp=1e-8, dist_thresh=None):
        self.numiters, self.damp, self.dist_thresh = numiters, damp, dist_thresh

    def provide(self, maps_pointclouds, frames_pointclouds):
        raise NotImplementedError
