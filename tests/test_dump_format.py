"""The device "%g" / "%d" formatter of the dump tap (lammps-ucg-dev_b200/csrc/dump_format.cuh), built for the host,
against the C library's snprintf — what DumpCustom::convert_string calls (dump_custom.cpp:1388-1421).  CPU only."""
import ctypes as C
import os
import subprocess

import pytest

import __graft_entry__ as g


@pytest.fixture(scope="module")
def fmt(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("fmt") / "libfmt.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(g.ROOT, "tests", "fmt_harness.cpp")])
    l = C.CDLL(so)
    l.fmt_check.restype = C.c_longlong
    return l


KNOWN = [0.0, -0.0, 1.0, -1.0, 0.5, 0.1, 10.0, 100000.0, 999999.0, 999999.5, 1000000.0, 1000005.0, 1000015.0, 123456.5, 1234565.0,
         0.0001, 0.00001, 0.000123456789, 9.9999949e-5, 9.9999951e-5, 0.99999949, 0.99999951, 123456789.0, 1e22, 1e23, 1e-300,
         5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, float("inf"), float("-inf"), float("nan"), 5.0387885741475218,
         -4.60194e+07, 1.05213e-08, 2.5, 0.15, 0.125, 1e100, 1e-100, 12.5]


def test_known_values(fmt):
    buf = C.create_string_buffer(64)
    for v in KNOWN:
        fmt.fmt_g(C.c_double(v), buf)
        assert buf.value.decode() == "%g" % v, v
    for v in (0, 1, -1, 7, 10, 99, 100, 2147483647, -2147483647, -2147483648, 1000188):
        fmt.fmt_d(C.c_int(v), buf)
        assert buf.value.decode() == "%d" % v


@pytest.mark.parametrize("mode,count", [(0, 1_000_000), (1, 1_000_000), (2, 500_000), (3, 500_000), (4, 500_000)])
def test_against_snprintf(fmt, mode, count):
    """0 random bit patterns, 1 dump-like magnitudes, 2 seven-digit decimals ending in 5 (+-1 ulp), 3 exact binary
    ties, 4 doubles around every (D + 0.5) * 10^k boundary"""
    bad = C.c_double()
    n = fmt.fmt_check(mode, C.c_ulonglong(99 + mode), C.c_longlong(count), C.byref(bad))
    assert n == 0, (n, repr(bad.value))
