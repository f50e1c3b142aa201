"""GPU parity tests (-m gpu): the product library on a real device against the oracle, all through the C ABI.

Tolerances are the ones BASELINE.json:north_star states: relative L2 error <= 1e-12 (double), <= 1e-5 (single);
pure permutations must be bit-exact."""
import numpy as np
import pytest

from cases import FASTCORE_CASES, CCC, CCC_B, CCC_S, CCR, CCR_S, PERMS, R2R_KINDS, RCC, RCC_S, TLOAD_CASES, half
from util import TOL, check_golden, check_golden_kernels, run_1d, run_3d

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["pipe", "nopipe"])
def both_pow2_kernels(request, monkeypatch):
    """pipelined kernel (default) and the plain pow2 kernel it falls back to for unaligned pointers"""
    if request.param == "nopipe":
        monkeypatch.setenv("P3DFFT_B200_NO_PIPE", "1")
    return request.param


def test_library_is_the_cuda_build(gpu):
    assert "sm_100a" in gpu.version()
    assert gpu.have_device()


@pytest.mark.parametrize("n", [(16, 12, 10), (8, 9, 7), (30, 3, 14), (58, 139, 199)])
def test_r2c_c2r_any_length(gpu, orc, n):
    """includes the reference matrix's uneven grid 58x139x199 (extra/makejob.py:132)"""
    assert run_3d(gpu, orc, n, half(n), RCC, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[8]
    assert run_3d(gpu, orc, half(n), n, CCR, (1, 2, 0), (0, 1, 2), cs1=0) < TOL[8]


@pytest.mark.parametrize("mo1", PERMS)
@pytest.mark.parametrize("mo2", PERMS)
def test_all_memory_order_pairs(gpu, orc, mo1, mo2):
    n = (128, 64, 96)
    assert run_3d(gpu, orc, n, half(n), RCC, mo1, mo2, cs2=0) < TOL[8]
    assert run_3d(gpu, orc, half(n), n, CCR, mo2, mo1, cs1=0) < TOL[8]


@pytest.mark.parametrize("m", [64, 128, 256, 512, 1024, 2048, 4096])
@pytest.mark.parametrize("types", [CCC, CCC_B, CCC_S])
def test_pow2_c2c_sizes(gpu, orc, m, types, both_pow2_kernels):
    n = (m, 24, 20)
    prec = 4 if types is CCC_S else 8
    for mo1, mo2 in (((0, 1, 2), (0, 1, 2)), ((1, 0, 2), (1, 0, 2)), ((0, 1, 2), (1, 2, 0)), ((2, 1, 0), (0, 2, 1))):
        assert run_3d(gpu, orc, n, n, types, mo1, mo2) < TOL[prec]


@pytest.mark.parametrize("m", [128, 256, 512, 1024, 2048, 4096])
def test_pow2_real_sizes(gpu, orc, m, both_pow2_kernels):
    n = (m, 12, 10)
    for mo1, mo2 in (((0, 1, 2), (0, 1, 2)), ((0, 1, 2), (1, 2, 0)), ((1, 2, 0), (0, 1, 2)), ((2, 0, 1), (1, 0, 2))):
        assert run_3d(gpu, orc, n, half(n), RCC, mo1, mo2, cs2=0) < TOL[8]
        assert run_3d(gpu, orc, half(n), n, CCR, mo2, mo1, cs1=0) < TOL[8]
        assert run_3d(gpu, orc, n, half(n), RCC_S, mo1, mo2, cs2=0) < TOL[4]
        assert run_3d(gpu, orc, half(n), n, CCR_S, mo2, mo1, cs1=0) < TOL[4]


def test_config_c1_single_rank_128(gpu, orc):
    """BASELINE config 1 grid (128^3 double R2C, X-pencil -> Z-pencil orders) on one rank + sine-wave known answer
    of sample/C++/test3D_r2c.C:281-331"""
    n = (128, 128, 128)
    err, out, want, _ = run_3d(gpu, orc, n, half(n), RCC, (0, 1, 2), (1, 2, 0), cs2=0, G=orc.sine_field(n), return_all=True)
    spec = orc.from_storage(out, (1, 2, 0)) / np.prod(n)
    expect = np.zeros_like(spec)
    for sy, iy in ((1, 1), (-1, n[1] - 1)):
        for sz, iz in ((1, 1), (-1, n[2] - 1)):
            expect[1, iy, iz] = 0.125j * sy * sz
    assert np.abs(spec - expect).max() < 1e-14 * n[0] * 0.25  # the reference's own gate
    assert err < TOL[8]


def test_config_c2_512_single(gpu, orc):
    """BASELINE config 2: 512^3 C2C single precision, default memory ordering, checked against the oracle"""
    n = (512, 512, 512)
    assert run_3d(gpu, orc, n, n, CCC_S, (0, 1, 2), (0, 1, 2)) < TOL[4]


def test_config_c4_dct_deriv(gpu, orc):
    """BASELINE config 4 (reduced to 128x128x129): R2C(x), C2C(y), DCT-I(z), non-default order, derivative"""
    n = (128, 128, 129)
    t = ["R2CFFT_D", "CFFT_FORWARD_D", "DCT1_COMPLEX_D"]
    assert run_3d(gpu, orc, n, half(n), t, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[8]
    for idir in (0, 1):
        assert run_3d(gpu, orc, n, half(n), t, (0, 1, 2), (1, 2, 0), cs2=0, deriv=idir) < TOL[8]
    n = (64, 64, 128)  # awkward length: 2*(128-1) = 254 = 2*127
    assert run_3d(gpu, orc, n, half(n), t, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[8]


@pytest.mark.parametrize("kind", [k for k in R2R_KINDS if k != "DCT4"])
@pytest.mark.parametrize("variant", ["REAL_D", "COMPLEX_D", "REAL_S", "COMPLEX_S"])
def test_r2r_kinds(gpu, orc, kind, variant):
    name = f"{kind}_{variant}"
    tol = TOL[4] if variant.endswith("_S") else TOL[8]
    for dim, n in ((0, (129, 14, 6)), (1, (12, 64, 5)), (2, (6, 9, 100))):
        for mo1, mo2 in (((0, 1, 2), (0, 1, 2)), ((0, 1, 2), (2, 0, 1)), ((1, 2, 0), (0, 1, 2))):
            assert run_1d(gpu, orc, n, name, dim, mo1, mo2) < tol


@pytest.mark.parametrize("name,n", FASTCORE_CASES + [("DCT1_COMPLEX_D", 513), ("DCT1_COMPLEX_D", 512), ("DST1_REAL_D", 1023),
                                                     ("DCT2_COMPLEX_D", 1024), ("DST3_REAL_S", 600), ("CFFT_FORWARD_D", 1000),
                                                     ("CFFT_BACKWARD_D", 768), ("R2CFFT_D", 1000), ("CFFT_FORWARD_S", 2000),
                                                     ("CFFT_FORWARD_D", 2048 - 1)])
def test_fastcore_kinds_and_bluestein(gpu, orc, name, n):
    """r2r kinds and non-power-of-two lengths on the register FFT core (fastcore_stage.cuh): core sizes 64...4096,
    directly (L a power of two) and through Bluestein; leading, strided and transposing layouts"""
    tol = TOL[4] if name.endswith("_S") else TOL[8]
    fc = ("fastcore", "pipe<")  # (r2r pencils a bulk copy can take go to the TMA-fed kernel: test_r2r_kinds_on_pipe_kernel)
    assert run_1d(gpu, orc, (n, 5, 3), name, 0, (0, 1, 2), (0, 1, 2), expect_variant=fc) < tol
    assert run_1d(gpu, orc, (9, n, 2), name, 1, (0, 1, 2), (0, 1, 2), expect_variant="fastcore") < tol
    assert run_1d(gpu, orc, (2, 17, n), name, 2, (0, 1, 2), (2, 0, 1), expect_variant=fc) < tol


@pytest.mark.parametrize("kind", [k for k in R2R_KINDS if k != "DCT4"])
@pytest.mark.parametrize("L", [64, 256, 1024, 4096])
def test_r2r_kinds_on_pipe_kernel(gpu, orc, kind, L, monkeypatch):
    """every r2r kind whose symmetric extension has a power-of-two length L on the TMA-fed kernel (pow2_pipe.cuh, KIND =
    kPipeR2R): complex and real data, double and single, contiguous and transposed stores, partial tiles; then the fallback"""
    n = {"DCT1": L // 2 + 1, "DST1": L // 2 - 1}.get(kind, L // 2)
    for variant in ("COMPLEX_D", "REAL_D", "COMPLEX_S", "REAL_S"):
        name = f"{kind}_{variant}"
        tol = TOL[4] if variant.endswith("_S") else TOL[8]
        esz = {"COMPLEX_D": 16, "REAL_D": 8, "COMPLEX_S": 8, "REAL_S": 4}[variant]
        want = "pipe<" if (n * esz) % 16 == 0 else "fastcore"  # dense user arrays: no room for a rounded-up bulk copy
        assert run_1d(gpu, orc, (n, 21, 3), name, 0, (0, 1, 2), (0, 1, 2), expect_variant=want) < tol, name
        assert run_1d(gpu, orc, (n, 21, 3), name, 0, (0, 1, 2), (1, 0, 2), expect_variant=want) < tol, (name, "transposed")
        assert run_1d(gpu, orc, (5, 6, n), name, 2, (1, 2, 0), (0, 1, 2), expect_variant=want) < tol, (name, "v-transposed")
    monkeypatch.setenv("P3DFFT_B200_NO_PIPE_R2R", "1")
    assert run_1d(gpu, orc, (n, 21, 3), f"{kind}_COMPLEX_D", 0, (0, 1, 2), (0, 1, 2), expect_variant="fastcore") < TOL[8]


@pytest.mark.parametrize("M", [384, 640, 768, 896, 1280, 1536, 1792, 2560, 3072, 3584, 1152, 1920, 2304, 192, 320, 448, 576, 960])
def test_smooth_lengths_on_mixed_radix_kernel(gpu, orc, M, monkeypatch):
    """lengths M = q * 2^k (q = 3, 5, 7, 9, 15; 2^k = 64...1024) on the TMA-fed mixed-radix kernel (mixed_pipe.cuh) instead of
    Bluestein: C2C forward / backward, R2C / C2R of 2M points, contiguous and transposed stores, partial tiles, both precisions"""
    e2 = ["EMPTY_TYPE_DOUBLE_COMPLEX"] * 2
    tag = [f",{q}x{M // q}>" for q in (3, 5, 7, 9, 15) if M % q == 0 and (M // q) & (M // q - 1) == 0]
    for types, n, n2, kw in ((["CFFT_FORWARD_D"] + e2, (M, 21, 3), (M, 21, 3), {}),
                             (["CFFT_BACKWARD_D"] + e2, (M, 21, 3), (M, 21, 3), {}),
                             (["R2CFFT_D"] + e2, (2 * M, 21, 3), (M + 1, 21, 3), dict(cs2=0)),
                             (["C2RFFT_D"] + e2, (M + 1, 21, 3), (2 * M, 21, 3), dict(cs1=0))):
        for mo2 in ((0, 1, 2), (1, 0, 2), (2, 0, 1)):
            err, _, _, desc = run_3d(gpu, orc, n, n2, types, (0, 1, 2), mo2, return_all=True, **kw)
            assert any(t in desc["stages"][0]["variant"] for t in tag), desc["stages"][0]["variant"]
            assert err < TOL[8], (types[0], mo2, err)
    n = (2 * M, 10, 6)
    assert run_3d(gpu, orc, n, half(n), RCC_S, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[4]
    assert run_3d(gpu, orc, half(n), n, CCR_S, (1, 2, 0), (0, 1, 2), cs1=0) < TOL[4]
    assert run_3d(gpu, orc, (M, 10, 6), (M, 10, 6), CCC_S, (0, 1, 2), (1, 2, 0)) < TOL[4]


def test_768_cubed_r2c_against_oracle(gpu, orc):
    """768^3 double R2C forward (extra/makejob.py:131-134 sizes): 768-point real stage (3 x 128 core) and 768-point complex
    stages (3 x 256) on the mixed-radix kernel, against the oracle"""
    n = (768, 768, 192)
    err, _, _, desc = run_3d(gpu, orc, n, half(n), RCC, (0, 1, 2), (1, 2, 0), cs2=0, return_all=True)
    assert err < TOL[8]
    assert all("pipe<" in s["variant"] for s in desc["stages"][:2]), [s["variant"] for s in desc["stages"]]
    assert run_3d(gpu, orc, half(n), n, CCR, (1, 2, 0), (0, 1, 2), cs1=0) < TOL[8]


def test_config_c4_full_size_against_oracle(gpu, orc):
    """BASELINE config 4 at its full size, 512 x 512 x 513 double: R2C(x) C2C(y) DCT-I(z) with the fused d/dx, against the
    oracle (the DCT stage reads padded 513-element rows by bulk copy: the r2r form of the TMA-fed kernel), and backward"""
    n = (512, 512, 513)
    t = ["R2CFFT_D", "CFFT_FORWARD_D", "DCT1_COMPLEX_D"]
    err, _, _, desc = run_3d(gpu, orc, n, half(n), t, (0, 1, 2), (1, 2, 0), cs2=0, deriv=0, return_all=True)
    assert err < TOL[8]
    assert any("r2r" in s["variant"] for s in desc["stages"]), [s["variant"] for s in desc["stages"]]
    tb = ["C2RFFT_D", "CFFT_BACKWARD_D", "DCT1_COMPLEX_D"]
    G = orc.transform_global(orc.random_field(n), t)  # a spectrum whose backward transform is real
    err, _, _, desc = run_3d(gpu, orc, half(n), n, tb, (1, 2, 0), (0, 1, 2), cs1=0, G=G, return_all=True)
    assert err < TOL[8]


def test_config_c5_shape_2048_point_stages(gpu, orc):
    """BASELINE config 5's stage shapes: 2048-point single-precision R2C(x) and C2C(y) stages on a 2048 x 2048 x 8 slab, and
    the 2048-point C2C(z) stage on 16 x 8 x 2048, forward and backward against the oracle"""
    for n in ((2048, 2048, 8), (16, 8, 2048)):
        assert run_3d(gpu, orc, n, half(n), RCC_S, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[4]
        assert run_3d(gpu, orc, half(n), n, CCR_S, (1, 2, 0), (0, 1, 2), cs1=0) < TOL[4]


@pytest.mark.parametrize("mode", ["ring", "register", "plain"])
def test_host_staging_modes(gpu, orc, mode):
    """pageable host arrays through every staging mode of the library (pinned ring with threaded CPU copies, page-locking
    the user's array, bare cudaMemcpy): 256^3 double (134 MB real, 135 MB complex: several ring chunks), the same arrays
    twice (register mode: the second call finds the ranges page-locked), released before they are freed as
    include/p3dfft_b200.h asks"""
    import ctypes
    n = (256, 256, 256)
    pg = gpu.init_proc_grid([1, 1, 1])
    g1 = gpu.init_data_grid(n, -1, pg, [0, 1, 2], [0, 1, 2])
    g2 = gpu.init_data_grid(half(n), 0, pg, [1, 2, 0], [1, 2, 0])
    pf = gpu.plan_3Dtrans(g1, g2, gpu.init_3Dtype(RCC))
    pb = gpu.plan_3Dtrans(g2, g1, gpu.init_3Dtype(CCR))
    G = orc.random_field(n)
    og1 = orc.OGrid(n, [0, 1, 2], [0, 1, 2], [1, 1, 1], 0)
    og2 = orc.OGrid(half(n), [1, 2, 0], [1, 2, 0], [1, 1, 1], 0, 0)
    a = np.ascontiguousarray(orc.local_of(G, og1))
    want = orc.local_of(orc.transform_global(G, RCC), og2)
    out = np.empty(og2.storage_shape(), dtype=np.complex128)
    back = np.empty_like(a)
    gpu.dll.p3dfft_b200_set_host_staging(mode.encode())
    try:
        for _ in range(2):
            out[...] = np.nan
            back[...] = np.nan
            gpu.exec_3Dtrans(pf, a, out, 0)
            gpu.exec_3Dtrans(pb, out, back, 0)
            assert orc.rel_l2(out, want) < TOL[8]
            assert orc.rel_l2(back / np.prod(n), a) < TOL[8]
    finally:
        for arr in (a, out, back):
            gpu.dll.p3dfft_b200_host_release(ctypes.c_void_p(arr.ctypes.data))
        gpu.dll.p3dfft_b200_set_host_staging(b"ring")
        gpu.free_data_grid(g1)
        gpu.free_data_grid(g2)


def test_known_answer_at_bench_size_1024(gpu, orc):
    """the reference sample's own known-answer test (sine field -> +-N/8 i at the modes (1, +-1, +-1), zero elsewhere;
    sample/C++/test3D_r2c.C:281-331, 343-369) at the headline size 1024^3 double, device resident -- a consistent
    wavenumber permutation passes the round-trip and Parseval properties but not this"""
    torch = pytest.importorskip("torch")
    import bench
    prob = bench.Problem(gpu, "r2c", (1024, 1024, 1024), False, [1, 1, 1])
    x = torch.empty(prob.n1e, device="cuda", dtype=torch.float64)
    X = torch.empty(prob.n2e, device="cuda", dtype=torch.complex128)
    err = bench.known_answer_check(torch, prob, x, X, torch.float64, torch.complex128, TOL[8])
    assert err < 1e-14 * 1024 * 0.25, err  # the reference's own gate, relative to the peak N/8
    del x, X
    torch.cuda.empty_cache()
    prob.free()


def test_fastcore_3d_c4_literal_and_bluestein_real(gpu, orc):
    """config C4 with z = 128 (L = 254 = 2*127: Bluestein) and 129 (L = 256), derivative; R2C/C2R of length 90"""
    t = ["R2CFFT_D", "CFFT_FORWARD_D", "DCT1_COMPLEX_D"]
    for nz in (129, 128):
        n = (64, 48, nz)
        assert run_3d(gpu, orc, n, half(n), t, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[8]
        assert run_3d(gpu, orc, n, half(n), t, (0, 1, 2), (1, 2, 0), cs2=0, deriv=2) < TOL[8]
    n = (90, 40, 26)
    assert run_3d(gpu, orc, n, half(n), RCC, (0, 1, 2), (1, 2, 0), cs2=0) < TOL[8]
    assert run_3d(gpu, orc, half(n), n, CCR, (1, 2, 0), (0, 1, 2), cs1=0) < TOL[8]


@pytest.mark.parametrize("mo1", PERMS)
@pytest.mark.parametrize("mo2", PERMS)
def test_1d_r2c_all_orders(gpu, orc, mo1, mo2):
    for dim in range(3):
        assert run_1d(gpu, orc, (64, 48, 32), "R2CFFT_D", dim, mo1, mo2) < TOL[8]


@pytest.mark.parametrize("idir", [0, 1, 2])
def test_fused_derivative(gpu, orc, idir):
    n = (128, 64, 32)
    assert run_3d(gpu, orc, n, half(n), RCC, (0, 1, 2), (1, 2, 0), cs2=0, deriv=idir) < TOL[8]
    assert run_3d(gpu, orc, n, half(n), RCC_S, (0, 1, 2), (1, 2, 0), cs2=0, deriv=idir) < TOL[4]
    assert run_3d(gpu, orc, (30, 20, 18), (16, 20, 18), RCC, (0, 1, 2), (1, 2, 0), cs2=0, deriv=idir) < TOL[8]


def test_in_place_and_empty(gpu, orc):
    n = (64, 32, 48)
    assert run_3d(gpu, orc, n, n, CCC, (0, 1, 2), (1, 2, 0), inplace=True) < TOL[8]
    assert run_3d(gpu, orc, n, half(n), RCC, (0, 1, 2), (0, 1, 2), cs2=0, inplace=True) < TOL[8]
    for mo1 in PERMS:
        for mo2 in PERMS:
            err, out, want, _ = run_3d(gpu, orc, (17, 9, 12), (17, 9, 12), ["EMPTY_TYPE_DOUBLE"] * 3, mo1, mo2, return_all=True)
            assert np.array_equal(out, want)


@pytest.mark.parametrize("mo", PERMS)
@pytest.mark.parametrize("idir", [0, 1, 2])
def test_compute_deriv_standalone(gpu, orc, mo, idir):
    n = (65, 24, 20)
    pg = gpu.init_proc_grid([1, 1, 1])
    g = gpu.init_data_grid(n, 0, pg, [0, 1, 2], list(mo))
    og = orc.OGrid(n, [0, 1, 2], mo, [1, 1, 1], 0, 0)
    a = orc.local_of(orc.random_field(n, complex_=True), og)
    out = np.zeros_like(a)
    gpu.compute_deriv(a, out, g, idir)
    assert orc.rel_l2(out, orc.compute_deriv_local(a, og, idir, mode="reference")) < 1e-15
    gpu.free_data_grid(g)


def test_device_pointers_roundtrip_1024_properties(gpu, orc):
    """full-size config (1024^3 double R2C+C2R, device resident): size-independent properties -
    round trip returns N*input and Parseval - since the oracle cannot hold this size in seconds"""
    torch = pytest.importorskip("torch")
    n = (1024, 1024, 1024)
    pg = gpu.init_proc_grid([1, 1, 1])
    g1 = gpu.init_data_grid(n, -1, pg, [0, 1, 2], [0, 1, 2])
    g2 = gpu.init_data_grid(half(n), 0, pg, [0, 1, 2], [1, 2, 0])
    pf = gpu.plan_3Dtrans(g1, g2, gpu.init_3Dtype(RCC))
    pb = gpu.plan_3Dtrans(g2, g1, gpu.init_3Dtype(CCR))
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n[2], n[1], n[0], device="cuda", dtype=torch.float64, generator=gen)
    X = torch.empty(n[1] * n[2] * (n[0] // 2 + 1), device="cuda", dtype=torch.complex128)
    y = torch.empty_like(x)
    gpu.exec_3Dtrans(pf, x, X, 0)
    gpu.exec_3Dtrans(pb, X, y, 1)
    gpu.sync()
    N = float(np.prod(n))
    y.div_(N).sub_(x)
    rt = (torch.linalg.vector_norm(y) / torch.linalg.vector_norm(x)).item()
    assert rt < TOL[8], rt
    # Parseval with Hermitian weights: sum |x|^2 = (1/N) (2 sum |X|^2 - sum_{kx=0} |X|^2 - sum_{kx=N/2} |X|^2)
    gpu.exec_3Dtrans(pf, x, X, 0)
    gpu.sync()
    Xv = torch.view_as_real(X).view(n[1], n[0] // 2 + 1, n[2], 2)  # storage (y, x, z) for memory order {1,2,0}
    p2 = (Xv ** 2).sum(dim=-1)
    tot = 2 * p2.sum() - p2[:, 0, :].sum() - p2[:, n[0] // 2, :].sum()
    lhs = (x ** 2).sum()
    assert abs((tot / N - lhs) / lhs).item() < 1e-12
    del x, X, y, Xv, p2
    torch.cuda.empty_cache()
    gpu.free_data_grid(g1)
    gpu.free_data_grid(g2)


def test_reference_golden_vectors_single_rank(gpu, orc):
    """every single-rank golden case (tests/golden: arrays written by the reference's own host code): all 36 memory-order
    pairs, C2R, the 1D API with the r2r kinds, stand-alone compute_deriv, the DCT4 registration quirk"""
    assert check_golden(gpu, orc, None, rank=0, world=1) >= 100


def test_reference_golden_vectors_at_kernel_sizes(gpu, orc):
    """the single-rank kernel-size golden cases (tests/golden/reference_golden_kernels.npz: outputs of the reference's own host
    code at M >= 64): the TMA-fed power-of-two, r2r and mixed-radix kernels pinned to the reference, not only to the oracle"""
    assert check_golden_kernels(gpu, orc, None, rank=0, world=1) >= 48


@pytest.mark.parametrize("case", TLOAD_CASES, ids=lambda c: "%s-%s-d%d" % ("x".join(map(str, c[0])), c[1], c[2]))
def test_tensor_load_kernel(gpu, orc, case):
    """pow2_tload.cuh on hardware: TMA tensor copies (cp.async.bulk.tensor, 2-D boxes of a 3-D tensor map) feed power-of-two
    stages whose input is unit-stride along another dimension than the transform's; zero-filled partial tiles"""
    g, t, dim, mo1, mo2, want = case
    assert run_1d(gpu, orc, g, t, dim, mo1, mo2, expect_variant=want) < TOL[4 if t.endswith("_S") else 8]


def test_tensor_load_kernel_3d_orders_and_fallback(gpu, orc, monkeypatch):
    """user arrays of 256 x 96 x 64 stored with y or z fastest: the R2C first stage takes the tensor-load kernel for every
    such order (many tiles per CTA: the prefetch of the next tile and the barrier phases), also with the fused derivative and
    in single precision; a device pointer that is not 16-byte aligned takes the plain-load kernel; P3DFFT_B200_NO_TLOAD=1"""
    n = (256, 96, 64)
    for mo1 in ((1, 0, 2), (2, 1, 0), (1, 2, 0), (2, 0, 1)):
        for mo2 in ((0, 1, 2), (1, 2, 0), (2, 1, 0)):
            err, _, _, desc = run_3d(gpu, orc, n, half(n), RCC, mo1, mo2, cs2=0, return_all=True)
            assert desc["stages"][0]["variant"].startswith("tload<"), desc["stages"][0]["variant"]
            assert err < TOL[8], (mo1, mo2, err)
    assert run_3d(gpu, orc, n, half(n), RCC, (1, 0, 2), (1, 2, 0), cs2=0, deriv=0) < TOL[8]
    assert run_3d(gpu, orc, n, half(n), RCC_S, (2, 1, 0), (1, 2, 0), cs2=0) < TOL[4]
    # device pointers: aligned -> tensor maps on the user's own array; offset by 8 bytes -> the plain-load fallback
    torch = pytest.importorskip("torch")
    pg = gpu.init_proc_grid([1, 1, 1])
    g1 = gpu.init_data_grid(n, -1, pg, [0, 1, 2], [1, 0, 2])
    g2 = gpu.init_data_grid(half(n), 0, pg, [0, 1, 2], [1, 2, 0])
    plan = gpu.plan_3Dtrans(g1, g2, gpu.init_3Dtype(RCC))
    og1 = orc.OGrid(n, [0, 1, 2], [1, 0, 2], [1, 1, 1], 0)
    og2 = orc.OGrid(half(n), [0, 1, 2], [1, 2, 0], [1, 1, 1], 0, 0)
    G = orc.random_field(n, complex_=False, key=99)
    a = np.ascontiguousarray(orc.local_of(G, og1), dtype=np.float64).ravel()
    want = orc.local_of(orc.transform_global(G, RCC, half(n)), og2)
    buf = torch.zeros(a.size + 2, device="cuda", dtype=torch.float64)
    outs = []
    for off in (0, 1):  # element offsets: 0 = 16-byte aligned (cudaMalloc), 1 = 8 bytes past it
        buf[off:off + a.size] = torch.from_numpy(a).cuda()
        X = torch.empty(want.size, device="cuda", dtype=torch.complex128)
        gpu.exec_3Dtrans(plan, buf[off:off + a.size], X, 0)
        gpu.sync()
        outs.append(X.cpu().numpy().reshape(want.shape))
        assert orc.rel_l2(outs[-1], want) < TOL[8], off
    del buf, X
    gpu.free_data_grid(g1)
    gpu.free_data_grid(g2)
    monkeypatch.setenv("P3DFFT_B200_NO_TLOAD", "1")
    err, _, _, desc = run_3d(gpu, orc, n, half(n), RCC, (1, 0, 2), (0, 1, 2), cs2=0, return_all=True)
    assert desc["stages"][0]["variant"].startswith("pow2<"), desc["stages"][0]["variant"]
    assert err < TOL[8]
