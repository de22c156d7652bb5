"""Accuracy of csrc/kem_math.cuh (branch-free exp / rcp / div), host build of the SAME
header against long-double libm.  The device build differs only in the reciprocal seed
(MUFU.RCP64H instead of a truncated host division); tests/test_gpu_math.py covers it."""
import ctypes as C
import math
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = tmp_path_factory.mktemp("kem_math") / "libkem_math_host.so"
    subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC",
                    "-I", os.path.join(ROOT, "knp-emi-fenics-x_b200", "csrc"),
                    "-o", str(out), os.path.join(ROOT, "tests", "native", "kem_math_host.cpp")], check=True)
    L = C.CDLL(str(out))
    L.kem_check_exp.restype = C.c_double
    L.kem_check_exp.argtypes = [C.c_long, C.c_double, C.c_double, C.c_uint64, C.POINTER(C.c_double)]
    L.kem_check_div.restype = C.c_double
    L.kem_check_div.argtypes = [C.c_long, C.c_int, C.c_uint64, C.POINTER(C.c_double)]
    for f in ("kem_host_exp", "kem_host_rcp"):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [C.c_double]
    L.kem_host_div.restype = C.c_double
    L.kem_host_div.argtypes = [C.c_double, C.c_double]
    L.kem_check_unary.restype = C.c_double
    L.kem_check_unary.argtypes = [C.c_int, C.c_long, C.c_int, C.c_uint64]
    for f in ("kem_host_log", "kem_host_sqrt", "kem_host_pow15"):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [C.c_double]
    return L


@pytest.mark.parametrize("lo,hi", [(-1e-5, 1e-5), (-1, 1), (-10, 10), (-50, 50), (-708, 709)])
def test_exp_within_1p1_ulp(hostlib, lo, hi):
    mean = C.c_double()
    worst = hostlib.kem_check_exp(400000, lo, hi, 7, C.byref(mean))
    # table-assisted exp: 0.5 ulp (table entry) + 0.5 ulp (final rounding) + 0.09 ulp (polynomial)
    assert worst < 1.1, worst
    assert mean.value < 0.36


def test_exp_special_values(hostlib):
    assert hostlib.kem_host_exp(0.0) == 1.0
    assert hostlib.kem_host_exp(1.0) == pytest.approx(math.e, rel=2.3e-16)
    assert math.isnan(hostlib.kem_host_exp(float("nan")))
    assert not math.isfinite(hostlib.kem_host_exp(float("inf")))
    # exponent field clamped (documented in the header): overflow -> +inf, underflow -> 0
    assert hostlib.kem_host_exp(1000.0) == float("inf")
    assert hostlib.kem_host_exp(-1000.0) == 0.0
    assert hostlib.kem_host_exp(709.0) == pytest.approx(math.exp(709.0), rel=2e-16)
    assert hostlib.kem_host_exp(-708.0) == pytest.approx(math.exp(-708.0), rel=2e-16)


@pytest.mark.parametrize("emax", [1, 30, 300])
def test_div_correctly_rounded_and_rcp_faithful_in_domain(hostlib, emax):
    """div: residual-corrected, 0.5 ulp; rcp: one cubic step from a 2^-20 seed, <= 0.51 ulp."""
    rworst = C.c_double()
    worst = hostlib.kem_check_div(400000, emax, 3, C.byref(rworst))
    assert worst <= 0.5 + 1e-9 and rworst.value <= 0.51


def test_removable_singularity_form_stays_comparable(hostlib):
    """x/(exp(x)-1) near x=0 (alpha_m at V=-40 mV, mm_hh.py:163): the kernel keeps the
    reference's exp(x)-1 form, so its error is the exp error amplified by 1/|x| exactly
    as in the reference."""
    for x in (1e-3, -1e-3, 1e-6, -1e-6):
        ref = x / (math.exp(x) - 1.0)
        got = hostlib.kem_host_div(x, hostlib.kem_host_exp(x) - 1.0)
        assert abs(got - ref) / abs(ref) < 2.3e-16 / abs(x) * 2


@pytest.mark.parametrize("which,name,bound", [(0, "log", 1.7), (1, "sqrt", 0.5 + 1e-9), (2, "pow15", 1.3)])
@pytest.mark.parametrize("emax", [1, 60, 600])
def test_log_sqrt_pow15_accuracy(hostlib, which, name, bound, emax):
    """table-assisted log <= 1.6 ulp (one rounding each for the table entry, k ln2 + logc and the
    final sum; also relative to itself as x -> 1); sqrt correctly rounded; x**1.5 = x*sqrt(x) <= 1.3 ulp
    (CUDA pow: 2 ulp)."""
    assert hostlib.kem_check_unary(which, 300000, emax, 11) <= bound


def test_log_special_values(hostlib):
    assert hostlib.kem_host_log(1.0) == 0.0
    assert hostlib.kem_host_log(0.0) == -math.inf
    assert math.isnan(hostlib.kem_host_log(-1.0))
    assert math.isnan(hostlib.kem_host_log(float("nan")))
    assert hostlib.kem_host_log(math.inf) == math.inf
    assert hostlib.kem_host_log(2.2250738585072014e-308) == math.log(2.2250738585072014e-308)
    assert hostlib.kem_host_log(1.7976931348623157e308) == math.log(1.7976931348623157e308)
    assert hostlib.kem_host_log(-0.0) == -math.inf
    assert math.isnan(hostlib.kem_host_sqrt(-1.0))
    assert hostlib.kem_host_sqrt(4.0) == 2.0
    assert hostlib.kem_host_sqrt(0.0) == 0.0 and math.copysign(1.0, hostlib.kem_host_sqrt(-0.0)) == -1.0
