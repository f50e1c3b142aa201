"""Kernel + planner + executor parity against the oracle, run on the CPU-thread emulation build.

The emulation build (tools/cuda_emu) compiles the SAME kernel sources with g++ and runs one OS thread per
CUDA thread; it is a development tool that lets index arithmetic be checked without a GPU.  The product
library is exercised by tests/test_gpu_parity.py (-m gpu)."""
import numpy as np
import pytest

from cases import CCC, CCC_B, CCC_S, CCR, CCR_S, FASTCORE_CASES, PERMS, R2R_KINDS, RCC, RCC_S, TLOAD_CASES, half
from util import TOL, check_golden, check_golden_kernels, run_1d, run_3d


def fast_variant(name):
    """pipelined (pow2_pipe.cuh) or plain (pow2_stage.cuh) register-radix kernel, not the generic one"""
    return name.startswith("pipe<") or name.startswith("pow2<")


@pytest.fixture(params=["pipe", "nopipe"])
def both_pow2_kernels(request, monkeypatch):
    """run a test once with the pipelined kernel (default) and once with the plain pow2 kernel (the fallback for
    unaligned pointers); the choice is read from the environment when a stage is created"""
    if request.param == "nopipe":
        monkeypatch.setenv("P3DFFT_B200_NO_PIPE", "1")
    return request.param


@pytest.mark.parametrize("n", [(16, 12, 10), (8, 9, 7), (6, 5, 4), (30, 3, 14)])
def test_r2c_c2r_generic(emu, orc, n):
    assert run_3d(emu, orc, n, half(n), RCC, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[8]
    assert run_3d(emu, orc, half(n), n, CCR, (1, 2, 0), (0, 1, 2), cs1=0) < TOL[8]


@pytest.mark.parametrize("n", [(11, 13, 7), (58, 3, 2), (2, 139, 2)])
def test_odd_and_prime_lengths(emu, orc, n):
    assert run_3d(emu, orc, n, n, CCC, (0, 1, 2), (0, 1, 2)) < TOL[8]
    assert run_3d(emu, orc, n, n, CCC_B, (0, 1, 2), (2, 1, 0)) < TOL[8]
    assert run_3d(emu, orc, n, half(n), RCC, (0, 1, 2), (0, 1, 2), cs2=0) < TOL[8]
    assert run_3d(emu, orc, half(n), n, CCR, (0, 1, 2), (0, 1, 2), cs1=0) < TOL[8]


@pytest.mark.parametrize("mo1", PERMS)
@pytest.mark.parametrize("mo2", PERMS)
def test_all_memory_order_pairs(emu, orc, mo1, mo2):
    """the 36 (mo1, mo2) pairs of sample/C/test3D_r2c_memord.c on one rank"""
    n = (8, 6, 10)
    assert run_3d(emu, orc, n, half(n), RCC, mo1, mo2, cs2=0) < TOL[8]
    assert run_3d(emu, orc, half(n), n, CCR, mo2, mo1, cs1=0) < TOL[8]


@pytest.mark.parametrize("m", [64, 128, 256, 512, 1024, 2048, 4096])
def test_pow2_c2c_sizes(emu, orc, m, both_pow2_kernels):
    """every size of the register-radix fast path, transform dimension leading"""
    n = (m, 3, 2)
    err, out, want, desc = run_3d(emu, orc, n, n, ["CFFT_FORWARD_D", "EMPTY_TYPE_DOUBLE_COMPLEX", "EMPTY_TYPE_DOUBLE_COMPLEX"],
                                  (0, 1, 2), (0, 1, 2), return_all=True)
    assert any(fast_variant(s["variant"]) for s in desc["stages"]), desc
    assert err < TOL[8]


@pytest.mark.parametrize("m", [128, 512, 2048])
def test_pow2_real_sizes(emu, orc, m, both_pow2_kernels):
    n = (m, 2, 3)
    err, _, _, desc = run_3d(emu, orc, n, half(n), ["R2CFFT_D", "EMPTY_TYPE_DOUBLE_COMPLEX", "EMPTY_TYPE_DOUBLE_COMPLEX"], (0, 1, 2),
                             (0, 1, 2), cs2=0, return_all=True)
    assert fast_variant(desc["stages"][0]["variant"])
    assert err < TOL[8]
    err, _, _, desc = run_3d(emu, orc, half(n), n, ["C2RFFT_D", "EMPTY_TYPE_DOUBLE_COMPLEX", "EMPTY_TYPE_DOUBLE_COMPLEX"], (0, 1, 2),
                             (0, 1, 2), cs1=0, return_all=True)
    assert fast_variant(desc["stages"][-1]["variant"])
    assert err < TOL[8]


def test_r2c_1024_symmetric_last_pass(emu, orc):
    """the headline's first stage: 1024-point R2C (512-point core whose radix-2 last pass pairs column j with M/2 - j so
    that the Hermitian split needs no exchange): contiguous and transposed stores, partial tiles, fused derivative
    (segment-table store path), single precision"""
    e3 = ["R2CFFT_D", "EMPTY_TYPE_DOUBLE_COMPLEX", "EMPTY_TYPE_DOUBLE_COMPLEX"]
    for n, mo2 in (((1024, 3, 2), (0, 1, 2)), ((1024, 9, 3), (1, 0, 2)), ((1024, 4, 16), (2, 1, 0))):
        err, _, _, desc = run_3d(emu, orc, n, half(n), e3, (0, 1, 2), mo2, cs2=0, return_all=True)
        assert desc["stages"][0]["variant"].startswith("pipe<f64,M=512"), desc["stages"][0]["variant"]
        assert err < TOL[8]
    n = (1024, 16, 4)
    assert run_3d(emu, orc, n, half(n), RCC, (0, 1, 2), (1, 2, 0), cs2=0, deriv=0) < TOL[8]
    assert run_3d(emu, orc, n, half(n), RCC_S, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[4]
    # radix-4 and radix-8 last passes (2048- and 4096-point real transforms), double and single
    e3s = ["R2CFFT_S", "EMPTY_TYPE_SINGLE_COMPLEX", "EMPTY_TYPE_SINGLE_COMPLEX"]
    for nx in (2048, 4096):
        assert run_3d(emu, orc, (nx, 3, 2), half((nx, 3, 2)), e3, (0, 1, 2), (0, 1, 2), cs2=0) < TOL[8]
        assert run_3d(emu, orc, (nx, 9, 2), half((nx, 9, 2)), e3, (0, 1, 2), (1, 0, 2), cs2=0) < TOL[8]
        assert run_3d(emu, orc, (nx, 2, 9), half((nx, 2, 9)), e3s, (0, 1, 2), (2, 1, 0), cs2=0) < TOL[4]


@pytest.mark.parametrize("mo1,mo2", [((0, 1, 2), (1, 2, 0)), ((1, 2, 0), (0, 1, 2)), ((2, 1, 0), (1, 0, 2)), ((0, 2, 1), (2, 0, 1))])
def test_pow2_strided_and_transposing(emu, orc, mo1, mo2, both_pow2_kernels):
    """fast path with the transform dimension not leading on one or both sides (64 and 128 points)"""
    n = (64, 128, 12)
    assert run_3d(emu, orc, n, n, CCC, mo1, mo2) < TOL[8]
    assert run_3d(emu, orc, n, half(n), RCC, mo1, mo2, cs2=0) < TOL[8]
    assert run_3d(emu, orc, half(n), n, CCR, mo2, mo1, cs1=0) < TOL[8]


def test_single_precision(emu, orc):
    n = (64, 20, 6)
    assert run_3d(emu, orc, n, n, CCC_S, (0, 1, 2), (0, 1, 2)) < TOL[4]
    assert run_3d(emu, orc, n, half(n), RCC_S, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[4]
    assert run_3d(emu, orc, half(n), n, CCR_S, (1, 2, 0), (0, 1, 2), cs1=0) < TOL[4]


@pytest.mark.parametrize("kind", R2R_KINDS)
@pytest.mark.parametrize("variant", ["REAL_D", "COMPLEX_D", "REAL_S"])
def test_r2r_kinds_1d(emu, orc, kind, variant, monkeypatch):
    """all eight FFTW r2r kinds through the 1D API, transform along each dimension (test1D_cos.C / test1D_sin.C)"""
    name = f"{kind}_{variant}"
    if kind == "DCT4":
        pytest.skip("DCT4 IDs follow the reference's registration (DCT-I planner); see test_dct4_quirk")
    tol = TOL[4] if variant.endswith("_S") else TOL[8]
    for dim, n in ((0, (9, 4, 3)), (1, (3, 12, 2)), (2, (2, 3, 7))):
        assert run_1d(emu, orc, n, name, dim, (0, 1, 2), (0, 1, 2)) < tol
    assert run_1d(emu, orc, (5, 6, 129 if kind == "DCT1" else 16), name, 2, (0, 1, 2), (2, 0, 1)) < tol


@pytest.mark.parametrize("name,n", FASTCORE_CASES)
def test_fastcore_kinds_and_bluestein(emu, orc, name, n):
    """r2r kinds and non-power-of-two lengths on the register FFT core (fastcore_stage.cuh), transform dimension leading,
    strided, and with a transposing output order"""
    tol = TOL[4] if name.endswith("_S") else TOL[8]
    fc = ("fastcore", "pipe<")  # (r2r pencils a bulk copy can take go to the TMA-fed kernel: test_r2r_kinds_on_pipe_kernel)
    assert run_1d(emu, orc, (n, 3, 2), name, 0, (0, 1, 2), (0, 1, 2), expect_variant=fc) < tol
    assert run_1d(emu, orc, (3, n, 2), name, 1, (0, 1, 2), (0, 1, 2), expect_variant="fastcore") < tol
    assert run_1d(emu, orc, (2, 5, n), name, 2, (0, 1, 2), (2, 0, 1), expect_variant=fc) < tol


@pytest.mark.parametrize("kind", [k for k in R2R_KINDS if k != "DCT4"])
def test_r2r_kinds_on_pipe_kernel(emu, orc, kind, monkeypatch):
    """every r2r kind whose symmetric extension has a power-of-two length L on the TMA-fed kernel (KIND = kPipeR2R): complex and
    real data, double and single, contiguous and transposed stores, L = 64 (8 values per thread), 128 and 512 (three passes);
    then the same transform with P3DFFT_B200_NO_PIPE_R2R=1 (fastcore, the fallback for unaligned pointers)"""
    n_of = {"DCT1": lambda L: L // 2 + 1, "DST1": lambda L: L // 2 - 1}.get(kind, lambda L: L // 2)
    for L in (64, 128, 512):
        n = n_of(L)
        for variant in ("COMPLEX_D", "REAL_D", "COMPLEX_S"):
            name = f"{kind}_{variant}"
            tol = TOL[4] if variant.endswith("_S") else TOL[8]
            bytes_ = n * (16 if variant == "COMPLEX_D" else 8)
            pipe = "pipe<" if bytes_ % 16 == 0 else "fastcore"  # dense user arrays: no room for a rounded-up bulk copy
            assert run_1d(emu, orc, (n, 3, 2), name, 0, (0, 1, 2), (0, 1, 2), expect_variant=pipe) < tol, (name, L)
            if L <= 128:
                assert run_1d(emu, orc, (n, 9, 2), name, 0, (0, 1, 2), (1, 0, 2), expect_variant=pipe) < tol, (name, L, "transposed")
    monkeypatch.setenv("P3DFFT_B200_NO_PIPE_R2R", "1")
    assert run_1d(emu, orc, (n_of(128), 3, 2), f"{kind}_COMPLEX_D", 0, (0, 1, 2), (0, 1, 2), expect_variant="fastcore") < TOL[8]


def test_fastcore_in_3d_with_derivative(emu, orc):
    """config C4's shape in small: R2C(x), C2C(y), DCT-I(z) with z = 2^k+1 (core directly; the TMA-fed r2r kernel when the
    padded intermediate rows leave room for the rounded-up bulk copy) and 2^k (Bluestein)"""
    t = ["R2CFFT_D", "CFFT_FORWARD_D", "DCT1_COMPLEX_D"]
    for nz in (33, 32):
        n = (16, 12, nz)
        assert run_3d(emu, orc, n, half(n), t, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[8]
        assert run_3d(emu, orc, n, half(n), t, (0, 1, 2), (1, 2, 0), cs2=0, deriv=2) < TOL[8]
    ts = ["R2CFFT_S", "CFFT_FORWARD_S", "DCT1_COMPLEX_S"]  # 65 x 8 bytes: the copy takes 8 bytes of the row padding
    err, _, _, desc = run_3d(emu, orc, (16, 12, 65), half((16, 12, 65)), ts, (0, 1, 2), (1, 2, 0), cs2=0, return_all=True)
    assert err < TOL[4] and any("r2r" in s["variant"] for s in desc["stages"]), [s["variant"] for s in desc["stages"]]
    err, _, _, desc = run_3d(emu, orc, (16, 12, 65), half((16, 12, 65)), t, (0, 1, 2), (1, 2, 0), cs2=0, deriv=2, return_all=True)
    assert err < TOL[8] and any("r2r" in s["variant"] for s in desc["stages"]), [s["variant"] for s in desc["stages"]]
    assert run_3d(emu, orc, (100, 12, 10), (100, 12, 10), CCC, (0, 1, 2), (2, 1, 0)) < TOL[8]
    n = (90, 40, 26)  # Bluestein R2C / C2R (Hermitian extension) round the 3D transform
    assert run_3d(emu, orc, n, half(n), RCC, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[8]
    assert run_3d(emu, orc, half(n), n, CCR, (1, 2, 0), (0, 1, 2), cs1=0) < TOL[8]


@pytest.mark.parametrize("mo1", PERMS)
@pytest.mark.parametrize("mo2", PERMS)
def test_1d_r2c_all_orders(emu, orc, mo1, mo2):
    """sample/C++/test_transplan.C matrix: 36 order pairs, each transform dimension"""
    for dim in range(3):
        assert run_1d(emu, orc, (8, 6, 4), "R2CFFT_D", dim, mo1, mo2) < TOL[8]


@pytest.mark.parametrize("idir", [0, 1, 2])
def test_fused_derivative(emu, orc, idir):
    n = (16, 12, 10)
    assert run_3d(emu, orc, n, half(n), RCC, (0, 1, 2), (1, 2, 0), cs2=0, deriv=idir) < TOL[8]
    n = (64, 9, 5)
    assert run_3d(emu, orc, n, half(n), RCC, (0, 1, 2), (1, 2, 0), cs2=0, deriv=idir) < TOL[8]
    assert run_3d(emu, orc, n, n, CCC, (0, 1, 2), (0, 1, 2), deriv=idir) < TOL[8]


def test_in_place(emu, orc):
    """exec(AR, AR, true) as in sample/C++/test3D_c2c_inplace.C:204-236, plus R2C in place"""
    n = (16, 12, 10)
    assert run_3d(emu, orc, n, n, CCC, (0, 1, 2), (1, 2, 0), inplace=True) < TOL[8]
    assert run_3d(emu, orc, n, half(n), RCC, (0, 1, 2), (0, 1, 2), cs2=0, inplace=True) < TOL[8]
    assert run_3d(emu, orc, n, n, ["EMPTY_TYPE_DOUBLE_COMPLEX"] * 3, (0, 1, 2), (2, 0, 1), inplace=True) < TOL[8]


def test_empty_types_are_pure_permutations(emu, orc):
    """empty transform = bit-exact reorder (reorder_out / reorder_in, exec.C:1858-2266)"""
    n = (7, 5, 6)
    for mo1 in PERMS:
        for mo2 in PERMS:
            err, out, want, _ = run_3d(emu, orc, n, n, ["EMPTY_TYPE_DOUBLE"] * 3, mo1, mo2, return_all=True)
            assert np.array_equal(out, want)


def test_2d_plus_empty(emu, orc):
    """sample/C/test2D+empty.c: FFT in two dimensions, empty type in the middle one"""
    n = (16, 6, 8)
    t = ["R2CFFT_D", "EMPTY_TYPE_DOUBLE_COMPLEX", "CFFT_FORWARD_D"]
    assert run_3d(emu, orc, n, half(n), t, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[8]


def test_dct_in_3d(emu, orc):
    """config C4 shape: R2C(x), C2C(y), DCT-I(z) on complex data, non-default output order"""
    n = (16, 8, 9)
    t = ["R2CFFT_D", "CFFT_FORWARD_D", "DCT1_COMPLEX_D"]
    assert run_3d(emu, orc, n, half(n), t, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[8]
    assert run_3d(emu, orc, n, half(n), t, (0, 1, 2), (1, 2, 0), cs2=0, deriv=1) < TOL[8]


@pytest.mark.parametrize("mo", PERMS)
@pytest.mark.parametrize("idir", [0, 1, 2])
def test_compute_deriv_standalone(emu, orc, mo, idir):
    """p3dfft_compute_deriv_double against deriv.C:85-185 (including its choice of storage dimension)"""
    n = (9, 6, 5)
    pg = emu.init_proc_grid([1, 1, 1])
    g = emu.init_data_grid(n, 0, pg, [0, 1, 2], list(mo))
    og = orc.OGrid(n, [0, 1, 2], mo, [1, 1, 1], 0, 0)
    a = orc.local_of(orc.random_field(n, complex_=True), og)
    out = np.zeros_like(a)
    emu.compute_deriv(a, out, g, idir)
    want = orc.compute_deriv_local(a, og, idir, mode="reference")
    assert orc.rel_l2(out, want) < 1e-15
    emu.free_data_grid(g)


def test_dct4_quirk(pkg, orc):
    """reference build/init.C:640-676 registers the DCT4 IDs with the DCT-I planner; the default build reproduces
    that, P3DFFT_B200_TRUE_DCT4=1 gives FFTW_REDFT11.  Checked in subprocesses because setup() reads the env once."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import __graft_entry__ as ge
from util import run_1d
mod = ge.load_package(); orc = ge.load_oracle()
lib = mod.load(emulated=True).setup()
n = (9, 4, 3)
pg = lib.init_proc_grid([1, 1, 1])
g = lib.init_data_grid(n, -1, pg, [0, 1, 2], [0, 1, 2])
plan = lib.plan_1Dtrans(g, g, "DCT4_REAL_D", 0)
G = orc.random_field(n)
og = orc.OGrid(n, [0, 1, 2], [0, 1, 2], [1, 1, 1], 0)
a = orc.local_of(G, og); out = np.zeros_like(a)
lib.exec_1Dtrans(plan, a, out, 0)
e1 = orc.rel_l2(out, orc.local_of(orc.transform_1d(G, "dct1", 0), og))
e4 = orc.rel_l2(out, orc.local_of(orc.transform_1d(G, "dct4", 0), og))
print("RESULT", e1 < 1e-12, e4 < 1e-12)
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    for env_val, want in (("0", "RESULT True False"), ("1", "RESULT False True")):
        env = dict(os.environ, P3DFFT_B200_TRUE_DCT4=env_val)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert want in out.stdout, (out.stdout, out.stderr)


def test_reference_golden_vectors_single_rank(emu, orc):
    """every single-rank golden case (tests/golden: arrays written by the reference's own host code): all 36 memory-order
    pairs, C2R, the 1D API with the r2r kinds, stand-alone compute_deriv, the DCT4 registration quirk"""
    assert check_golden(emu, orc, None, rank=0, world=1) >= 100


@pytest.mark.parametrize("q,mc", [(3, 128), (5, 128), (7, 128), (3, 256), (3, 512), (9, 128), (15, 128), (3, 64), (15, 64)])  # (the gpu test runs every served length)
def test_smooth_lengths_on_mixed_radix_kernel(emu, orc, q, mc, monkeypatch):
    """lengths M = q * 2^k (q = 3, 5, 7, 9, 15) on the TMA-fed mixed-radix kernel (mixed_pipe.cuh): C2C forward / backward, R2C / C2R of
    2M points, contiguous and transposed stores, partial tiles, double and single; then Bluestein (P3DFFT_B200_NO_MIXED=1)"""
    M = q * mc
    e2 = ["EMPTY_TYPE_DOUBLE_COMPLEX"] * 2
    small = mc <= 128
    for types, n, n2, kw in ((["CFFT_FORWARD_D"] + e2, (M, 3, 2), (M, 3, 2), {}),
                             (["CFFT_BACKWARD_D"] + e2, (M, 3, 2), (M, 3, 2), {}),
                             (["R2CFFT_D"] + e2, (2 * M, 3, 2), (M + 1, 3, 2), dict(cs2=0)),
                             (["C2RFFT_D"] + e2, (M + 1, 3, 2), (2 * M, 3, 2), dict(cs1=0))):
        for mo2 in ((0, 1, 2), (1, 0, 2)) if small else ((0, 1, 2),):
            err, _, _, desc = run_3d(emu, orc, n, n2, types, (0, 1, 2), mo2, return_all=True, **kw)
            assert f",{q}x{mc}>" in desc["stages"][0]["variant"] or types[0].startswith("C2R"), desc["stages"][0]["variant"]
            assert err < TOL[8], (types[0], mo2, err)
    if small:
        err, _, _, desc = run_3d(emu, orc, (M, 9, 2), (M, 9, 2), ["CFFT_FORWARD_S", "EMPTY_TYPE_SINGLE_COMPLEX", "EMPTY_TYPE_SINGLE_COMPLEX"],
                                 (0, 1, 2), (1, 0, 2), return_all=True)
        assert err < TOL[4] and f",{q}x{mc}>" in desc["stages"][0]["variant"], desc["stages"][0]["variant"]
        n = (2 * M, 5, 3)
        assert run_3d(emu, orc, n, half(n), RCC_S, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[4]
        assert run_3d(emu, orc, half(n), n, CCR_S, (1, 2, 0), (0, 1, 2), cs1=0) < TOL[4]
        monkeypatch.setenv("P3DFFT_B200_NO_MIXED", "1")
        err, _, _, desc = run_3d(emu, orc, (M, 3, 2), (M, 3, 2), ["CFFT_FORWARD_D"] + e2, (0, 1, 2), (0, 1, 2), return_all=True)
        assert err < TOL[8] and "bluestein" in desc["stages"][0]["variant"], desc["stages"][0]["variant"]


def test_reference_golden_vectors_at_kernel_sizes_emulated(emu, orc):
    """a subset of the kernel-size golden cases on the emulation (the whole set runs on the GPU): 128 x 64 x 64 forward,
    mixed-radix 768, DCT-I of 513 points, run-time r2r kinds, user arrays stored with y or z fastest (tensor-load kernel)"""
    names = ["k_fwd_128x64x64", "k_fwd_128x64x64_mo102", "k_fwd_128x64x64_mo210", "k_fwd_deriv0_128x64x64_mo120", "k_fwd_single_128x64x64_mo210",
             "k_t1d_strided_CFFT_BACKWARD_D_1024_d2_012_210", "k_t1d_strided_R2CFFT_D_512_d0_102_012", "k_fwd_768x6x4", "k_t1d_CFFT_FORWARD_D_768_d0_012_120", "k_t1d_DCT1_COMPLEX_D_513_d0_012_012",
             "k_t1d_DCT2_REAL_D_512_d0_012_120", "k_t1d_DST1_COMPLEX_D_255_d0_012_012", "k_c4_dct_deriv0_64x16x129"]
    assert check_golden_kernels(emu, orc, names, rank=0, world=1) == len(names)


@pytest.mark.parametrize("case", TLOAD_CASES, ids=lambda c: "%s-%s-d%d" % ("x".join(map(str, c[0])), c[1], c[2]))
def test_tensor_load_kernel(emu, orc, case):
    """pow2_tload.cuh: power-of-two stages whose input is unit-stride along another dimension than the transform's (the
    emulation restates the TMA tensor copy as a box gather with zero fill: it pins the tile / box / coordinate arithmetic)"""
    g, t, dim, mo1, mo2, want = case
    assert run_1d(emu, orc, g, t, dim, mo1, mo2, expect_variant=want) < TOL[4 if t.endswith("_S") else 8]


def test_tensor_load_kernel_in_3d_and_switch(emu, orc, monkeypatch):
    """the tensor-load kernel as the first stage of a 3D transform of a user array in a non-default storage order (with the
    fused derivative), and P3DFFT_B200_NO_TLOAD=1 (the plain-load kernel it replaces)"""
    n = (128, 24, 16)  # the real-to-complex dimension goes first whatever its stride
    for mo1 in ((1, 0, 2), (2, 1, 0), (1, 2, 0), (2, 0, 1)):
        for mo2 in ((0, 1, 2), (1, 2, 0)):
            err, _, _, desc = run_3d(emu, orc, n, half(n), RCC, mo1, mo2, cs2=0, return_all=True)
            assert desc["stages"][0]["variant"].startswith("tload<"), desc["stages"][0]["variant"]
            assert err < TOL[8]
    err, _, _, desc = run_3d(emu, orc, n, half(n), RCC, (1, 0, 2), (1, 2, 0), cs2=0, deriv=0, return_all=True)
    assert desc["stages"][0]["variant"].startswith("tload<"), desc["stages"][0]["variant"]
    assert err < TOL[8]
    assert run_3d(emu, orc, n, half(n), RCC_S, (2, 1, 0), (1, 2, 0), cs2=0) < TOL[4]
    monkeypatch.setenv("P3DFFT_B200_NO_TLOAD", "1")
    err, _, _, desc = run_3d(emu, orc, n, half(n), RCC, (1, 0, 2), (0, 1, 2), cs2=0, return_all=True)
    assert desc["stages"][0]["variant"].startswith("pow2<"), desc["stages"][0]["variant"]
    assert err < TOL[8]
