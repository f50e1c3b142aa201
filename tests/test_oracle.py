"""Pins the oracle (oracle/p3dfft_oracle.py, oracle/cfft/cfft.c) before anything is checked against it:
  * golden vectors produced by the REFERENCE's own host code (tests/golden/make_golden.py -> oracle/_ref/ref_driver):
    per-rank output arrays, Ldims and GlobStart of 129 cases (3D R2C/C2R/C2C on 1, 3 and 4 ranks, all 36 memory-order
    pairs, fused derivative, config-4 DCT stage, empty types, the 1D transplan API with eight r2r kinds, stand-alone
    compute_deriv incl. its inverse-permutation choice of the storage dimension, the DCT4 = DCT-I registration quirk);
  * the known answers the reference's samples test (test3D_r2c.C:281-331, test1D_cos.C:254-306, test1D_sin.C,
    test_deriv2.C:354-431);
  * the plain-C FFTW restatement against numpy.fft / scipy.fft on every kind and awkward lengths.
CPU only."""
import ctypes
import os
import subprocess
from ctypes import byref, c_int, c_void_p

import numpy as np
import pytest
import scipy.fft as sfft

from util import TOL, golden, golden_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return golden()


def test_golden_file_is_complete(gold):
    index, z, mg = gold
    assert len(index) == 129
    assert [c["name"] for c in mg.cases()] == index, "tests/golden/index.json is stale: rerun tests/golden/make_golden.py"


def test_oracle_matches_reference_golden_vectors(gold, orc):
    index, z, mg = gold
    worst = {4: 0.0, 8: 0.0}
    for name in index:
        c = golden_case(z, name)
        pd = c["procdims"]
        G = mg.global_input(c)
        prec = mg.case_types(c)[2]
        for r in range(pd[0] * pd[1] * pd[2]):
            og1, og2, want = mg.oracle_output(c, G, r)
            ref = z[f"{name}/out_{r}"]
            # geometry: bit-exact (reference build/init.C:1699-1863)
            assert og1.Ldims + og1.GlobStart + og2.Ldims + og2.GlobStart == list(z[f"{name}/meta_{r}"]), (name, r)
            assert ref.size == want.size, (name, r)
            if ref.size:
                err = orc.rel_l2(want.ravel(), ref)
                worst[prec] = max(worst[prec], err)
                assert err < (1e-6 if prec == 4 else 1e-13), (name, r, err)
            if c["mode"] == "deriv":  # multiplication by an integer wavenumber: bit-exact
                assert np.array_equal(want.ravel(), ref), name
    assert worst[8] < 1e-14


def test_oracle_matches_reference_golden_vectors_at_kernel_sizes(orc):
    """the 53 kernel-size cases (128 x 64 x 64 on 1 / 2 / 4 ranks, also stored with y or z fastest, 512...4096-point stages, 768 / 640 / 896 / 1536-point
    mixed-radix lengths, r2r kinds with 128...1024-point extensions): sampled elements + norms of the reference's outputs"""
    from util import compare_with_kernel_golden, golden_kernels
    index, z, mg = golden_kernels()
    assert len(index) == 53 and [c["name"] for c in mg.large_cases()] == index, "index_kernels.json is stale: rerun make_golden.py"
    for name in index:
        c = golden_case(z, name)
        pd = c["procdims"]
        G = mg.global_input(c)
        prec = mg.case_types(c)[2]
        for r in range(pd[0] * pd[1] * pd[2]):
            og1, og2, want = mg.oracle_output(c, G, r)
            assert og1.Ldims + og1.GlobStart + og2.Ldims + og2.GlobStart == list(z[f"{name}/meta_{r}"]), (name, r)
            err = compare_with_kernel_golden(orc, z, mg, name, r, want.astype(mg.np_dtype(mg.case_types(c)[1], prec)), prec)
            assert err < (1e-6 if prec == 4 else 1e-13), (name, r, err)


def test_sample_known_answer_sine_spectrum(orc):
    """sample/C++/test3D_r2c.C:281-331: forward/N^3 of sin*sin*sin is +-0.125i at wavenumbers (1|N-1)^3"""
    n = (16, 12, 10)
    F = orc.transform_global(orc.sine_field(n), ["R2CFFT_D", "CFFT_FORWARD_D", "CFFT_FORWARD_D"]) / np.prod(n)
    expect = np.zeros_like(F)
    for sy, iy in ((1, 1), (-1, n[1] - 1)):
        for sz, iz in ((1, 1), (-1, n[2] - 1)):
            expect[1, iy, iz] = 0.125j * sy * sz
    assert np.abs(F - expect).max() < 1e-14 * n[0] * 0.25  # the sample's own tolerance


def test_sample_known_answer_dct1_dst1(orc):
    """sample/C++/test1D_cos.C:254-306: DCT-I of cos(j pi/(N-1)) scaled by 0.5/(N-1) is 0.5 at k=1;
    sample/C++/test1D_sin.C: DST-I of sin((j+1) pi/(N+1)) scaled by 0.5/(N+1) is 0.5 at k=0"""
    N = 129
    x = np.cos(np.arange(N) * np.pi / (N - 1))
    y = orc.transform_1d(x, "dct1", 0) * 0.5 / (N - 1)
    e = np.zeros(N)
    e[1] = 0.5
    assert np.abs(y - e).max() < 1e-14
    x = np.sin((np.arange(N) + 1) * np.pi / (N + 1))
    y = orc.transform_1d(x, "dst1", 0) * 0.5 / (N + 1)
    e = np.zeros(N)
    e[0] = 0.5
    assert np.abs(y - e).max() < 1e-14


@pytest.mark.parametrize("idir", [0, 1, 2])
def test_sample_known_answer_derivative(orc, idir):
    """sample/C++/test_deriv2.C:354-431: forward with derivative in idir, normalise, backward -> cos in idir, sin elsewhere"""
    n = (16, 12, 10)
    F = orc.transform_global(orc.sine_field(n), ["R2CFFT_D", "CFFT_FORWARD_D", "CFFT_FORWARD_D"], deriv_dim=idir) / np.prod(n)
    back = orc.transform_global(F, ["C2RFFT_D", "CFFT_BACKWARD_D", "CFFT_BACKWARD_D"], n)
    f = [np.sin(2 * np.pi * np.arange(m) / m) for m in n]
    f[idir] = np.cos(2 * np.pi * np.arange(n[idir]) / n[idir])
    expect = f[0][:, None, None] * f[1][None, :, None] * f[2][None, None, :]
    assert np.abs(back - expect).max() < 1e-13


def test_dct4_ids_follow_reference_registration(orc):
    assert orc.type_info("DCT4_REAL_D")[0] == "dct1" and orc.type_info("DST4_REAL_D")[0] == "dst4"


# ------------------------------------------------------------------------------------------------ C restatement of FFTW
@pytest.fixture(scope="module")
def cfft():
    path = os.path.join(ROOT, "oracle", "_build", "libcfft.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_build/libcfft.so"])
    L = ctypes.CDLL(path)
    for f in ("fftw_plan_many_dft", "fftw_plan_many_dft_r2c", "fftw_plan_many_dft_c2r", "fftw_plan_many_r2r", "fftwf_plan_many_dft",
              "fftwf_plan_many_r2r"):
        getattr(L, f).restype = c_void_p
    return L


def _p(a):
    return c_void_p(a.ctypes.data)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 12, 16, 30, 58, 64, 139, 199, 512, 1022])
def test_cfft_matches_numpy(cfft, n):
    rng = np.random.default_rng(n)
    hm = 3
    x = rng.standard_normal((hm, n)) + 1j * rng.standard_normal((hm, n))
    y = np.zeros_like(x)
    for sign, ref in ((-1, np.fft.fft(x, axis=1)), (1, np.fft.ifft(x, axis=1) * n)):
        pl = cfft.fftw_plan_many_dft(1, byref(c_int(n)), hm, None, None, 1, n, None, None, 1, n, sign, 0)
        cfft.fftw_execute_dft(c_void_p(pl), _p(x), _p(y))
        assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-14
    xr = rng.standard_normal((hm, n))
    h = n // 2 + 1
    yc = np.zeros((hm, h), complex)
    pl = cfft.fftw_plan_many_dft_r2c(1, byref(c_int(n)), hm, None, None, 1, n, None, None, 1, h, 0)
    cfft.fftw_execute_dft_r2c(c_void_p(pl), _p(xr), _p(yc))
    ref = np.fft.rfft(xr, axis=1)
    assert np.linalg.norm(yc - ref) / np.linalg.norm(ref) < 1e-14
    back = np.zeros_like(xr)
    pl = cfft.fftw_plan_many_dft_c2r(1, byref(c_int(n)), hm, None, None, 1, h, None, None, 1, n, 0)
    cfft.fftw_execute_dft_c2r(c_void_p(pl), _p(yc), _p(back))
    assert np.linalg.norm(back / n - xr) / np.linalg.norm(xr) < 1e-14
    if n >= 2:
        kinds = {3: ("dct", 1), 5: ("dct", 2), 4: ("dct", 3), 6: ("dct", 4), 7: ("dst", 1), 9: ("dst", 2), 8: ("dst", 3), 10: ("dst", 4)}
        for kind, (fam, typ) in kinds.items():
            out = np.zeros_like(xr)
            pl = cfft.fftw_plan_many_r2r(1, byref(c_int(n)), hm, None, None, 1, n, None, None, 1, n, byref(c_int(kind)), 0)
            cfft.fftw_execute_r2r(c_void_p(pl), _p(xr), _p(out))
            ref = (sfft.dct if fam == "dct" else sfft.dst)(xr, type=typ, axis=1, norm=None)
            assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 2e-14, (n, kind)


def test_cfft_strided_single_precision(cfft):
    rng = np.random.default_rng(0)
    n = 16
    x = (rng.standard_normal((n, 5)) + 1j * rng.standard_normal((n, 5))).astype(np.complex64)
    y = np.zeros_like(x)
    pl = cfft.fftwf_plan_many_dft(1, byref(c_int(n)), 5, None, None, 5, 1, None, None, 5, 1, -1, 0)
    cfft.fftwf_execute_dft(c_void_p(pl), _p(x), _p(y))
    assert np.linalg.norm(y - np.fft.fft(x, axis=0)) / np.linalg.norm(y) < 1e-6
