"""Exhaustive CPU check of the reduced-range exponential K4 uses for the confidence weight
(gradslam_b200/csrc/gsx_exp.cuh; alpha = clamp(exp(-|v|^2 / 2 sigma^2), 1e-7, 1.01), fusionutils.py:69-72).

The header is compiled for the host (g++ -mfma: fma() is the same IEEE operation as the GPU's DFMA) and evaluated for
EVERY float32 argument in [-17, -0]: results that are not flagged "undecided" must round to the same float32 as libm's
float64 exp - the canonical value the oracle uses - and the public wrapper (fast path + library fallback) must agree
for every argument."""
import ctypes
import os
import struct
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _bits(f):
    return struct.unpack("<I", struct.pack("<f", f))[0]


@pytest.fixture(scope="module")
def lib():
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "exp_host.so")
    src = os.path.join(HERE, "host", "exp_host.cpp")
    subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src], check=True)
    lib = ctypes.CDLL(so)
    lib.exp_scan.argtypes = [ctypes.c_uint32, ctypes.c_uint32] + [ctypes.c_void_p] * 4
    return lib


def _scan(lib, lo, hi):
    und, bad, wbad, err = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_double()
    lib.exp_scan(lo, hi, ctypes.byref(und), ctypes.byref(bad), ctypes.byref(wbad), ctypes.byref(err))
    return und.value, bad.value, wbad.value, err.value


def test_every_float32_argument_down_to_minus_17(lib):
    # negative floats: bit patterns grow with magnitude; [-17, -2^-40] is 370 M arguments (~10 s)
    und, bad, wbad, err = _scan(lib, _bits(-2.0 ** -40), _bits(-17.0) + 1)
    assert bad == 0 and wbad == 0
    assert err < 1.5  # double ulps against libm
    assert und < 1000  # the library fallback is taken for a handful of arguments only


def test_tiny_arguments_and_zero(lib):
    # a sample of the 1.4 G arguments in [-2^-40, -0]: exp(x) rounds to 1 or its predecessor
    for lo, hi in ((_bits(-0.0), _bits(-0.0) + 2_000_000), (_bits(-2.0 ** -41), _bits(-2.0 ** -40)),
                   (_bits(-2.0 ** -60), _bits(-2.0 ** -60) + 2_000_000)):
        und, bad, wbad, err = _scan(lib, lo, hi)
        assert bad == 0 and wbad == 0 and und == 0
    assert _scan(lib, _bits(0.0), _bits(0.0) + 1)[2] == 0  # +0


def test_sqrt_threshold(lib):
    """gsx_thresholds.h: for every threshold t the decision `sqrtf(x) < t` equals `x <= sqrt_lt_threshold(t)` - scanned
    2000 floats either side of the boundary and 2 M random non-negative floats per threshold (host sqrtf is the same
    correctly rounded IEEE operation as the device's)."""
    lib.sqrt_threshold_scan.argtypes = [ctypes.c_float, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p]
    lib.sqrt_threshold_scan.restype = ctypes.c_uint64
    thr = ctypes.c_float()
    for t in (0.05, 0.02, 0.2, 1.0, 1e-3, 0.3, 7.0, 1e-20, 1e19, 3.0000002, 0.049999997):
        assert lib.sqrt_threshold_scan(t, 2000, 2_000_000, ctypes.byref(thr)) == 0, t
        assert thr.value >= 0
    for t in (0.0, -1.0, float("nan")):
        assert lib.sqrt_threshold_scan(t, 10, 1000, ctypes.byref(thr)) == 0
        assert thr.value == -1.0
    assert lib.sqrt_threshold_scan(float("inf"), 10, 1000, ctypes.byref(thr)) == 0
