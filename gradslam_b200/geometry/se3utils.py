"""SO(3)/SE(3) hat and exponential maps (mirror of gradslam/geometry/se3utils.py).

These are host-callable, differentiable torch helpers for API parity.  Inside the ICP loop the same maps
are evaluated on the device by csrc/gsx_icp.cu (one warp per batch element), with the same small-angle
branch: for ||omega|| < 1e-6 BOTH R and V are I + hat(omega) (se3utils.py:91-93).
"""
import torch

_eps = 1e-6

__all__ = ["so3_hat", "se3_hat", "so3_exp", "se3_exp"]


def so3_hat(omega: torch.Tensor) -> torch.Tensor:
    assert torch.is_tensor(omega), "Input must be of type torch.tensor."
    o = omega.reshape(-1)
    z = torch.zeros((), dtype=o.dtype, device=o.device)
    return torch.stack((torch.stack((z, -o[2], o[1])), torch.stack((o[2], z, -o[0])), torch.stack((-o[1], o[0], z))))


def se3_hat(xi: torch.Tensor) -> torch.Tensor:
    assert torch.is_tensor(xi), "Input must be of type torch.tensor."
    x = xi.reshape(-1)
    top = torch.cat((so3_hat(x[3:]), x[:3].view(3, 1)), dim=1)
    return torch.cat((top, torch.zeros(1, 4, dtype=x.dtype, device=x.device)), dim=0)


def _rodrigues(omega):
    """Returns (R, V) of the SE(3) exponential for rotation vector omega."""
    W = so3_hat(omega)
    I = torch.eye(3, dtype=omega.dtype, device=omega.device)
    theta = omega.norm()
    if theta < _eps:
        return I + W, I + W
    s, c = theta.sin(), theta.cos()
    W2 = W.mm(W)
    A = s / theta
    B = (1 - c) / torch.pow(theta, 2)
    C = (theta - s) / torch.pow(theta, 3)
    return I + A * W + B * W2, I + B * W + C * W2


def so3_exp(omega: torch.Tensor) -> torch.Tensor:
    assert torch.is_tensor(omega), "Input must be of type torch.Tensor."
    return _rodrigues(omega.reshape(-1))[0]


def se3_exp(xi: torch.Tensor) -> torch.Tensor:
    assert torch.is_tensor(xi), "Input must be of type torch.tensor."
    x = xi.reshape(-1)
    R, V = _rodrigues(x[3:])
    t = V.mm(x[:3].view(3, 1))
    bottom = torch.tensor([[0, 0, 0, 1]], dtype=x.dtype, device=x.device)
    return torch.cat((torch.cat((R, t), dim=1), bottom), dim=0)
