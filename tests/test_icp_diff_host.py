"""CPU check of the scalar-generic K7 arithmetic (gradslam_b200/csrc/gsx_icp_diff.cuh): the header is compiled for
the host (g++, tests/host/icp_diff_host.cpp) and its forward values and dual-number Jacobians - the exact code the
backward kernels run, one lane per input - are compared with float64 autograd of the reference formulas
(solve_linear_system icputils.py:22-90, se3_exp se3utils.py:77-115, LM / gradLM update icputils.py:356-365, 519-543)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
F64 = torch.float64


@pytest.fixture(scope="module")
def host():
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "icp_diff_host.so")
    src = os.path.join(HERE, "host", "icp_diff_host.cpp")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src], check=True)
    lib = ctypes.CDLL(so)
    lib.host_update.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    lib.host_solve.argtypes = [ctypes.c_void_p] * 3
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _se3_exp(xi):
    v, w = xi[:3], xi[3:]
    z = torch.zeros((), dtype=xi.dtype)
    W = torch.stack([torch.stack([z, -w[2], w[1]]), torch.stack([w[2], z, -w[0]]), torch.stack([-w[1], w[0], z])])
    th = torch.sqrt((w * w).sum())
    eye = torch.eye(3, dtype=xi.dtype)
    if th < 1e-6:
        R = V = eye + W
    else:
        s, c, W2 = th.sin(), th.cos(), W @ W
        R = eye + s / th * W + (1 - c) / th ** 2 * W2
        V = eye + (1 - c) / th ** 2 * W + (th - s) / th ** 3 * W2
    top = torch.cat([R, V @ v.view(3, 1)], 1)
    return torch.cat([top, torch.tensor([[0, 0, 0, 1.0]], dtype=xi.dtype)], 0)


def _solve(inp):
    iu = torch.triu_indices(6, 6)
    U = torch.zeros(6, 6, dtype=inp.dtype).index_put((iu[0], iu[1]), inp[:21])
    M = U + U.t() - torch.diag(torch.diagonal(U))
    xi = torch.inverse(M + torch.eye(6, dtype=inp.dtype) * inp[28]) @ inp[21:27]
    return torch.cat([xi, _se3_exp(xi).reshape(-1)])


def _update(inp, mode, lmax, B, B2, nu):
    xi, err, nerr, damp, T = inp[:6], inp[6], inp[7], inp[8], inp[9:].view(4, 4)
    if mode == 0:
        if nerr < err:
            dT, d = _se3_exp(xi), damp / 2
            Tn = dT @ T
        else:
            dT, d, Tn = torch.eye(4, dtype=inp.dtype), damp * 2, T
    else:
        diff = (nerr - err).clamp(-70, 70)
        d = damp * (1 / lmax + (lmax - 1 / lmax) / (1 + torch.exp(-B * diff)))
        sig = 1 / ((1 + torch.exp(-B2 * diff)) ** (1 / nu))
        dT = _se3_exp(sig * xi)
        Tn = dT @ T
    return torch.cat([d.view(1), dT.reshape(-1), Tn.reshape(-1)])


@pytest.mark.parametrize("seed,damp", [(0, 1e-3), (1, 1e-8), (2, 0.5)])
def test_solve_step_values_and_jacobian(host, seed, damp):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(200, 6, dtype=F64, generator=g)
    b = torch.randn(200, dtype=F64, generator=g) * 0.05
    M = A.t() @ A
    iu = torch.triu_indices(6, 6)
    inp = torch.cat([M[iu[0], iu[1]], A.t() @ b, (b * b).sum().view(1), torch.tensor([damp], dtype=F64)])
    J = torch.autograd.functional.jacobian(_solve, inp).numpy()  # (22, 29)
    i32 = inp.float().numpy().copy()
    out, jac = np.zeros(22, np.float32), np.zeros((29, 22), np.float32)
    host.host_solve(_p(i32), _p(out), _p(jac))
    np.testing.assert_allclose(out, _solve(inp).numpy(), atol=2e-6)
    assert np.abs(jac.T - J).max() <= 2e-4 * np.abs(J).max()


@pytest.mark.parametrize("mode,err,new_err", [(1, 0.5, 0.3), (1, 0.3, 0.5), (0, 0.5, 0.3), (0, 0.3, 0.5),
                                              (1, 0.0, 100.0), (1, 100.0, 0.0)])
def test_update_step_values_and_jacobian(host, mode, err, new_err):
    g = torch.Generator().manual_seed(int(err * 10 + new_err * 100) + mode)
    xi = torch.randn(6, dtype=F64, generator=g) * 0.05
    T = _se3_exp(torch.randn(6, dtype=F64, generator=g) * 0.3)
    inp = torch.cat([xi, torch.tensor([err, new_err, 1e-3], dtype=F64), T.reshape(-1)])

    def f(x):
        return _update(x, mode, 2.0, 1.0, 1.0, 200.0)

    J = torch.autograd.functional.jacobian(f, inp).numpy()  # (33, 25)
    i32 = inp.float().numpy().copy()
    out, jac = np.zeros(33, np.float32), np.zeros((25, 33), np.float32)
    host.host_update(_p(i32), mode, 2.0, 1.0, 1.0, 200.0, _p(out), _p(jac))
    np.testing.assert_allclose(out, f(inp).numpy(), atol=2e-6)
    np.testing.assert_allclose(jac.T, J, atol=2e-5)


def test_small_angle_branch(host):
    """||omega|| < 1e-6: R = V = I + hat(omega) (se3utils.py:91-93), and the Jacobian of that branch."""
    inp = torch.tensor([0.01, -0.02, 0.03, 1e-8, -2e-8, 1e-8, 0.5, 0.3, 1e-3] + torch.eye(4).reshape(-1).tolist(),
                       dtype=F64)
    J = torch.autograd.functional.jacobian(lambda x: _update(x, 0, 2.0, 1.0, 1.0, 200.0), inp).numpy()
    i32 = inp.float().numpy().copy()
    out, jac = np.zeros(33, np.float32), np.zeros((25, 33), np.float32)
    host.host_update(_p(i32), 0, 2.0, 1.0, 1.0, 200.0, _p(out), _p(jac))
    np.testing.assert_allclose(jac.T, J, atol=1e-6)
