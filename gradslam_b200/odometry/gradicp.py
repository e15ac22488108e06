from .base import OdometryProvider

__all__ = ["GradICPOdometryProvider"]


class GradICPOdometryProvider(OdometryProvider):
    def __init__(self, numiters=20, damp=1e-8, dist_thresh=None, lambda_max=2.0, B=1.0, B2=1.0, nu=200.0):
        self.numiters, self.damp, self.dist_thresh = numiters, damp, dist_thresh
        self.lambda_max, self.B, self.B2, self.nu = lambda_max, B, B2, nu

    def provide(self, maps_pointclouds, frames_pointclouds):
        raise NotImplementedError
