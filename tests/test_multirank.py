"""N>1 path on CPU: planner + executor + fused exchange (peer stores, flag barrier) across real processes.

Each rank is a process of the mini-MPI world (include/compat/mpi.h) running the CPU-thread emulation build, in
which "device" buffers are POSIX shared memory and the CUDA-IPC mapping / peer barrier run unchanged.  Every rank
compares its local output with the oracle's slice of the global transform (tests/mp_worker.py).  One test uses
torch.distributed.run + gloo, the launcher and environment bench.py runs under on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

from cases import CCC, CCR, RCC, RCC_S, half

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
XP = dict(dmap1=[0, 1, 2], mo1=[0, 1, 2], dmap2=[1, 2, 0], mo2=[1, 2, 0])  # X-pencil -> Z-pencil (test3D_r2c.C:153-179)


def fwd(n, pd, types=RCC, **kw):
    c = dict(types=types, procdims=pd, gdims1=list(n), gdims2=list(half(n)), cs2=0, **XP)
    c.update(kw)
    return c


def bwd(n, pd, types=CCR, **kw):
    c = dict(types=types, procdims=pd, gdims1=list(half(n)), gdims2=list(n), cs1=0, dmap1=XP["dmap2"], mo1=XP["mo2"],
             dmap2=XP["dmap1"], mo2=XP["mo1"])
    c.update(kw)
    return c


def c2c(n, pd, **kw):
    c = dict(types=CCC, procdims=pd, gdims1=list(n), gdims2=list(n), **XP)
    c.update(kw)
    return c


def launch(nranks, cases, mode="emu", gloo=False, timeout=900, env_extra=None):
    arg = cases if isinstance(cases, str) else json.dumps(cases)
    if gloo:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 2000), os.path.join(HERE, "mp_worker.py"), mode, arg, "--gloo"]
    else:
        cmd = [sys.executable, os.path.join(ROOT, "tools", "mpirun.py"), "-np", str(nranks), sys.executable,
               os.path.join(HERE, "mp_worker.py"), mode, arg]
    env = dict(os.environ)
    env.pop("P3DFFT_B200_PLAN_ONLY", None)
    env.setdefault("P3DFFT_B200_PEER_TIMEOUT_S", "120")  # a rank that dies must fail the others instead of hanging them
    env.update(env_extra or {})
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    ok = out.returncode == 0 and out.stdout.count(" OK worst") == nranks
    assert ok, (out.stdout[-3000:], out.stderr[-3000:])


@pytest.fixture(scope="module", autouse=True)
def _emu_built(emu):
    return emu


def test_pencil_2x2_config_c1_shape():
    """BASELINE config 1 shape (2x2 pencil grid on 4 ranks), reduced to 32x24x20 + the uneven split 9 = 4|5"""
    n = (32, 24, 20)
    launch(4, [fwd(n, [1, 2, 2]), bwd(n, [1, 2, 2]), fwd((16, 12, 10), [1, 2, 2]), bwd((16, 12, 10), [1, 2, 2])])


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_reference_golden_vectors_multirank(nranks):
    """per-rank outputs of the reference's own host code (tests/golden) on 2, 3 and 4 ranks, including the kernel-size cases
    (128 x 64 x 64 on slab and pencil grids: the TMA-fed kernels storing through the per-peer segment tables)"""
    launch(nranks, "golden")


def test_slab_and_row_grids():
    """slab {1,1,P} (one exchange) and {1,P,1}; uneven blocks (10 = 3|3|4 over 3 ranks does not occur with P=2,4: use 4)"""
    n = (16, 14, 10)
    launch(4, [fwd(n, [1, 1, 4]), bwd(n, [1, 1, 4]), fwd(n, [1, 4, 1]), bwd(n, [1, 4, 1]), c2c(n, [1, 1, 4])])


def test_world_size_2_gloo_torchrun():
    """world_size 2 under torch.distributed.run with a gloo group next to the library's own rendezvous"""
    n = (16, 12, 10)
    launch(2, [fwd(n, [1, 1, 2]), bwd(n, [1, 1, 2]), fwd(n, [1, 2, 1])], gloo=True)


def test_uneven_and_prime_sizes_3_ranks():
    n = (14, 7, 11)
    launch(3, [fwd(n, [1, 1, 3]), bwd(n, [1, 1, 3]), c2c((5, 7, 11), [1, 3, 1])])


def test_pow2_kernels_with_exchange_segments():
    """64-point stages take the register-radix kernel; its stores go through the per-peer segment table"""
    n = (128, 64, 64)
    launch(2, [fwd(n, [1, 1, 2], reps=1), bwd(n, [1, 1, 2], reps=1)], timeout=1500)
    launch(4, [fwd((128, 64, 4), [1, 2, 2], reps=1), bwd((128, 64, 4), [1, 2, 2], reps=1)], timeout=1500)


def test_more_ranks_than_planes():
    """ragged edge: a distributed dimension shorter than the processor grid leaves some ranks with EMPTY local blocks
    (block distribution init.C:1834-1862 gives the first P - N%P ranks floor(N/P) = 0 planes); those ranks still take part in
    every barrier and receive their share of the output"""
    cs = [fwd((8, 6, 3), [1, 1, 4]), bwd((8, 6, 3), [1, 1, 4]), fwd((8, 3, 5), [1, 4, 1]), bwd((8, 3, 5), [1, 4, 1]),
          c2c((4, 3, 3), [1, 2, 2]), c2c((5, 1, 3), [1, 2, 2])]
    launch(4, cs)
    launch(4, cs[:2], env_extra={"P3DFFT_B200_OVERLAP_ALIGN": "1", "P3DFFT_B200_OVERLAP_CHUNKS": "2"})


def test_in_place_multirank_and_odd_grids():
    """in == out with the overwrite flag across ranks (also with overlapped pairs, whose exchange-first form must not let the
    local stage overwrite the exchange stage's input), and processor grids that are not powers of two"""
    n = (16, 12, 10)
    ip = [fwd(n, [1, 1, 2], inplace=True), bwd(n, [1, 1, 2], inplace=True), c2c(n, [1, 1, 2], inplace=True),
          c2c(n, [1, 1, 2], inplace=True, dmap2=[0, 1, 2], mo2=[0, 1, 2], expect_pairs=False)]
    launch(2, ip)
    launch(2, ip, env_extra=FORCE_PAIRS)
    launch(6, [fwd((12, 9, 10), [1, 3, 2]), bwd((12, 9, 10), [1, 3, 2]), c2c((7, 9, 10), [1, 2, 3]), fwd((12, 9, 10), [1, 2, 3], deriv=2)])


def test_r2c_1024_with_exchange_segments():
    """pencil grid: the 1024-point R2C stage (symmetric-column last pass) is itself an exchange stage, so its out-of-order
    rows go through the per-peer segment table; also as the local stage of an overlapped pair on a slab grid"""
    n = (1024, 8, 8)
    launch(4, [fwd(n, [1, 2, 2], reps=1), bwd(n, [1, 2, 2], reps=1)], timeout=1500)
    launch(2, [fwd(n, [1, 1, 2], reps=1), fwd(n, [1, 1, 2], reps=1, deriv=0)], timeout=1500,
           env_extra={"P3DFFT_B200_OVERLAP_ALIGN": "1", "P3DFFT_B200_OVERLAP_CHUNKS": "2"})


def test_memory_orders_and_derivative_multirank():
    n = (16, 12, 10)
    cs = []
    for mo1, mo2 in (([1, 0, 2], [2, 1, 0]), ([2, 0, 1], [0, 2, 1]), ([0, 2, 1], [1, 0, 2])):
        cs.append(fwd(n, [1, 2, 2], mo1=mo1, mo2=mo2))
        cs.append(bwd(n, [1, 2, 2], mo1=mo2, mo2=mo1))
    for idir in (0, 1, 2):
        cs.append(fwd(n, [1, 2, 2], deriv=idir))
    cs.append(fwd(n, [1, 2, 2], types=RCC_S))
    launch(4, cs)


def test_other_distributions_and_empty_types():
    """same distribution in and out, Y-pencil outputs, and MPI-only redistribution of an untransformed array"""
    n = (12, 10, 8)
    e = ["EMPTY_TYPE_DOUBLE"] * 3
    cs = [
        c2c(n, [1, 2, 2], dmap2=[0, 1, 2], mo2=[0, 1, 2]),                 # in and out both X-pencils
        c2c(n, [1, 2, 2], dmap2=[1, 0, 2], mo2=[1, 0, 2]),                 # Y-pencil output
        c2c(n, [1, 2, 2], dmap1=[2, 1, 0], mo1=[2, 1, 0], dmap2=[0, 1, 2], mo2=[0, 1, 2]),
        dict(types=e, procdims=[1, 2, 2], gdims1=list(n), gdims2=list(n), dmap1=[0, 1, 2], mo1=[0, 1, 2], dmap2=[1, 2, 0], mo2=[2, 0, 1]),
        dict(types=["R2CFFT_D", "EMPTY_TYPE_DOUBLE_COMPLEX", "CFFT_FORWARD_D"], procdims=[1, 2, 2], gdims1=list(n), gdims2=list(half(n)),
             cs2=0, **XP),                                                 # sample/C/test2D+empty.c
    ]
    launch(4, cs)


def test_smooth_lengths_with_exchange_segments():
    """mixed-radix kernel (192 = 3 x 64, 640 = 5 x 128 points) as an exchange stage: its stores go through the per-peer segment
    table; R2C of 384 points (192-point core) in front of an exchange"""
    T = dict(reps=1)
    launch(2, [c2c((192, 12, 8), [1, 2, 1], expect_variant="pipe<f64,M=192", **T), fwd((384, 12, 8), [1, 1, 2], expect_variant="pipe<f64,M=192", **T),
               bwd((384, 12, 8), [1, 2, 1], **T), c2c((640, 4, 6), [1, 1, 2], expect_variant="pipe<f64,M=640", **T)], timeout=1500)


def test_tensor_load_first_stage_multirank():
    """user arrays stored with y or z fastest: the first stage (R2C along the strided x) runs the tensor-load kernel
    (pow2_tload.cuh) -- as a plain stage, as the exchange stage of a pencil grid and as a member of chunked pairs"""
    n = (128, 64, 16)
    T = dict(expect_variant="tload<", reps=1)
    launch(2, [fwd(n, [1, 1, 2], mo1=[1, 0, 2], **T), fwd(n, [1, 1, 2], mo1=[2, 1, 0], deriv=1, **T), fwd(n, [1, 2, 1], mo1=[1, 2, 0], **T)],
           timeout=1500)
    launch(2, [fwd(n, [1, 1, 2], mo1=[1, 0, 2], **T)], timeout=1500, env_extra={"P3DFFT_B200_OVERLAP_ALIGN": "1", "P3DFFT_B200_OVERLAP_CHUNKS": "2"})
    launch(4, [fwd(n, [1, 2, 2], mo1=[2, 0, 1], **T)], timeout=1500)


FORCE_PAIRS = {"P3DFFT_B200_OVERLAP_ALIGN": "1", "P3DFFT_B200_OVERLAP_CHUNKS": "3", "P3DFFT_TEST_EXPECT_PAIRS": "1"}


def test_overlapped_exchange_pairs():
    """an exchange stage and its neighbouring local stage cut into chunks on two streams (per-chunk peer barriers):
    slab forward (local stage first) and backward (exchange first), pencil grids, uneven blocks with empty chunks,
    fused derivative; small grids are cut by lifting the chunk alignment"""
    n = (16, 12, 10)
    launch(2, [fwd(n, [1, 1, 2]), bwd(n, [1, 1, 2]), c2c(n, [1, 1, 2]), fwd(n, [1, 1, 2], deriv=1), fwd(n, [1, 1, 2], deriv=0)],
           env_extra=FORCE_PAIRS)
    launch(3, [fwd((14, 7, 11), [1, 1, 3]), bwd((14, 7, 11), [1, 1, 3])], env_extra=FORCE_PAIRS)
    launch(4, [fwd((32, 24, 20), [1, 2, 2]), bwd((32, 24, 20), [1, 2, 2]), fwd(n, [1, 2, 2], deriv=2), c2c(n, [1, 4, 1])],
           env_extra=FORCE_PAIRS)
    launch(2, [fwd(n, [1, 1, 2]), bwd(n, [1, 1, 2])], env_extra={"P3DFFT_B200_OVERLAP": "0"})


SYNC_PAIRS = {"P3DFFT_TEST_EXPECT_PAIRS": "1", "P3DFFT_TEST_EXPECT_SYNC": "1", "P3DFFT_TEST_EXPECT_TRIPLE": "1"}


def test_persistent_pair_kernels_with_flags():
    """overlapped pairs as ONE persistent launch per stage: the chunks are tile groups, "chunk complete" is a flag word written
    inside the producing kernel (local stage -> exchange stage on this rank; exchange stage -> every peer's local stage) and
    awaited inside the consuming one.  Slab forward (L then X) and backward (X then L), uneven blocks over 3 ranks, pencil
    grid, fused derivative in either member, in-place.  Forward slab plans run as a TRIPLE (asserted): local stage -> exchange
    stage -> last local stage as three persistent kernels, the exchange stage publishing the second half of its work piece by
    piece to the peers' last stages"""
    n = (128, 64, 64)
    T = dict(expect_triple=True)
    launch(2, [fwd(n, [1, 1, 2], **T), bwd(n, [1, 1, 2]), fwd(n, [1, 1, 2], deriv=1, **T), fwd(n, [1, 1, 2], deriv=0, **T),
               fwd(n, [1, 1, 2], deriv=2, **T), c2c((64, 64, 64), [1, 1, 2], inplace=True, **T)], env_extra=SYNC_PAIRS, timeout=1500)
    launch(4, [fwd((128, 64, 128), [1, 1, 4], reps=1, **T)], env_extra=SYNC_PAIRS, timeout=1500)
    launch(2, [fwd(n, [1, 1, 2], reps=1)], env_extra={"P3DFFT_B200_TRIPLE": "0", "P3DFFT_TEST_EXPECT_SYNC": "1"}, timeout=1500)
    small = dict(SYNC_PAIRS, P3DFFT_B200_OVERLAP_ALIGN="1", P3DFFT_B200_OVERLAP_CHUNKS="3")
    launch(3, [fwd((128, 64, 20), [1, 1, 3], reps=2), bwd((128, 64, 64), [1, 1, 3], reps=2)], env_extra=small, timeout=1500)
    launch(4, [fwd((128, 64, 64), [1, 2, 2], reps=1), bwd((128, 64, 64), [1, 2, 2], reps=1)], env_extra=small, timeout=1500)
    launch(2, [fwd(n, [1, 1, 2], reps=1), bwd(n, [1, 1, 2], reps=1)], env_extra={"P3DFFT_B200_PAIR_SYNC": "0", "P3DFFT_TEST_EXPECT_PAIRS": "1"},
           timeout=1500)


# ------------------------------------------------------------------------------------------------ real GPUs
def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_gpu_multirank_parity(nranks):
    """the same cases on N real GPUs (one process per GPU, peer stores over NVLink).  On a box with fewer GPUs the ranks
    share the devices (rank r -> device r mod count; CUDA IPC maps buffers between processes of one device as well, the
    contexts are time-sliced), so the exchange kernels, peer barriers and overlapped pairs run on hardware whatever the box"""
    if _ngpu() < 1:
        pytest.skip("needs a GPU")
    n = (128, 96, 64)
    grids = [[1, 1, nranks]] + ([[1, 2, nranks // 2]] if nranks >= 4 else [[1, nranks, 1]])
    cs = []
    for pd in grids:
        cs += [fwd(n, pd), bwd(n, pd), fwd(n, pd, types=RCC_S), fwd(n, pd, deriv=1), c2c((64, 50, 36), pd)]
    cs.append(fwd((58, 139, 199), grids[0]))
    cs.append(bwd((58, 139, 199), grids[-1]))
    launch(nranks, cs, mode="gpu", timeout=1200)
    if nranks in (2, 4):
        launch(nranks, "golden", mode="gpu")


@pytest.mark.gpu
@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_gpu_overlapped_pairs_parity(nranks):
    """shapes large enough to plan overlapped pairs WITHOUT lifting the chunk alignment (the production configuration of the
    1024^3 runs): persistent pair kernels with tile-group flags (asserted), then the same cases with one launch + one peer
    barrier per chunk (P3DFFT_B200_PAIR_SYNC=0); slab and pencil grids, single precision, fused derivative"""
    if _ngpu() < 1:
        pytest.skip("needs a GPU")
    n = (256, 128, 32 * nranks)
    pd = [1, 1, nranks]
    T = dict(expect_triple=True)  # forward slab plans: L -> X -> Z as three persistent kernels
    cs = [fwd(n, pd, **T), bwd(n, pd), fwd(n, pd, types=RCC_S, **T), fwd(n, pd, deriv=1, **T), fwd(n, pd, deriv=0, **T),
          fwd(n, pd, deriv=2, **T), c2c((128, 128, 32 * nranks), pd, **T), c2c((128, 128, 32 * nranks), pd, inplace=True, **T)]
    launch(nranks, cs, mode="gpu", timeout=1200, env_extra=SYNC_PAIRS)
    launch(nranks, cs[:2], mode="gpu", timeout=1200, env_extra={"P3DFFT_B200_TRIPLE": "0", "P3DFFT_TEST_EXPECT_SYNC": "1"})
    launch(nranks, cs[:2], mode="gpu", timeout=1200, env_extra={"P3DFFT_B200_PAIR_SYNC": "0", "P3DFFT_TEST_EXPECT_PAIRS": "1"})
    if nranks >= 4:
        pp = [1, 2, nranks // 2]
        m = (256, 128, 64 * (nranks // 2))
        launch(nranks, [fwd(m, pp), bwd(m, pp)], mode="gpu", timeout=1200, env_extra={"P3DFFT_TEST_EXPECT_PAIRS": "1"})


@pytest.mark.gpu
def test_gpu_overlapped_pairs_512x512x256():
    """4 ranks, 512 x 512 x 256 double: eight chunks per pair as in the headline runs, against the oracle"""
    if _ngpu() < 1:
        pytest.skip("needs a GPU")
    n = (512, 512, 256)
    launch(4, [fwd(n, [1, 1, 4], reps=2, expect_triple=True), bwd(n, [1, 1, 4], reps=2)], mode="gpu", timeout=1500, env_extra=SYNC_PAIRS)


@pytest.mark.parametrize("switch", ["P3DFFT_B200_NO_PAD", "P3DFFT_B200_NO_PIPE", "P3DFFT_B200_NO_FASTCORE", "P3DFFT_B200_FORCE_GENERIC"])
def test_kernel_ladder_switches(switch):
    """the A/B switches of INTEGRATION.md section 7 keep working: dense intermediates, plain pow2 kernel instead of the TMA one,
    generic kernel instead of fastcore / of everything (run in worker processes because they are read at plan creation)"""
    cs = [fwd((128, 64, 16), [1, 1, 2], reps=1), bwd((128, 64, 16), [1, 1, 2], reps=1), c2c((100, 30, 8), [1, 2, 1], reps=1)]
    launch(2, cs, env_extra={switch: "1"}, timeout=1500)


def test_random_multirank_cases_fixed_seed():
    """two batches of tools/fuzz_multirank.py (random grids / processor grids / orders / kinds / chunking) with a fixed seed"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_multirank.py"), "3", "2"], capture_output=True, text=True,
                         timeout=1500)
    assert out.returncode == 0 and "failures 0" in out.stdout, (out.stdout[-3000:], out.stderr[-2000:])
