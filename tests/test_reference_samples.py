"""The reference's own sample programs (sample/C++/*.C, sample/C/*.c), compiled UNMODIFIED against this library
(samples/Makefile), must pass their own checks: the drop-in test a P3DFFT++ user would run first.

CPU: built on the emulation library when /root/reference is present (the authoring container) and run on 1 and 4 ranks.
GPU (-m gpu): the prebuilt samples/_build binaries (they travel to the box with the library) on the real device.
The samples print "Results are correct" / "Results are incorrect" (sample/C++/test3D_r2c.C:247-256, 281-331)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SAMPLES = os.path.join(ROOT, "samples")
REFERENCE = os.environ.get("REFERENCE", "/root/reference")

# name -> (input file, extra leading fields after "nx ny nz")
THREE_D = ["test3D_r2c_cpp", "test3D_c2c_cpp", "test3D_c2c_inplace_cpp", "test3D_r2c_single_cpp", "test_deriv_cpp",
           "test3D_r2c_c", "test3D_c2c_c", "test3D_c2c_inplace_c", "test3D_r2c_single_c", "test_deriv_c", "test2D+empty_c"]
ONE_D = ["test1D_cos_cpp", "test1D_cos_complex_cpp", "test1D_sin_cpp", "test_transplan_cpp", "test1D_cos_c", "test1D_cos_complex_c",
         "test1D_r2c_c"]


def run_sample(bindir, name, n, pdims, tmp_path, nranks, mo1=(0, 1, 2), mo2=(0, 1, 2), dim=0):
    exe = os.path.join(bindir, name)
    if not os.path.exists(exe):
        pytest.skip(f"{name} not built")
    d = tmp_path / (name.replace("+", "_") + f"_{nranks}")
    d.mkdir()
    idir = " 1" if name.startswith("test_deriv") else ""
    (d / "stdin").write_text(f"{n[0]} {n[1]} {n[2]} 2 1{idir}\n")
    (d / "trans.in").write_text(f"{n[0]} {n[1]} {n[2]} {dim} 1\n{mo1[0]} {mo1[1]} {mo1[2]}\n{mo2[0]} {mo2[1]} {mo2[2]}\n")
    (d / "memord3d").write_text(f"{n[0]} {n[1]} {n[2]} 2 1\n{mo1[0]} {mo1[1]} {mo1[2]}\n{mo2[0]} {mo2[1]} {mo2[2]}\n")
    (d / "dims").write_text(f"{pdims[0]} {pdims[1]}\n")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mpirun.py"), "-np", str(nranks), exe], cwd=d,
                         capture_output=True, text=True, timeout=900)
    text = out.stdout + out.stderr
    assert out.returncode == 0, text[-3000:]
    assert "Results are correct" in text and "incorrect" not in text, text[-3000:]
    return text


@pytest.fixture(scope="module")
def emu_samples(emu):
    if not os.path.exists(os.path.join(REFERENCE, "sample", "C++", "test3D_r2c.C")):
        pytest.skip("reference sources not present (samples are built from where they lie)")
    subprocess.check_call(["make", "-s", "-j8", "EMU=1", f"REFERENCE={REFERENCE}"], cwd=SAMPLES)
    return os.path.join(SAMPLES, "_build_emu")


@pytest.mark.parametrize("name", THREE_D)
def test_reference_3d_samples_emulated(emu_samples, name, tmp_path):
    """cubic 16^3 grid (the samples' own known-answer checks assume it), one rank and the 2x2 pencil grid of config C1"""
    run_sample(emu_samples, name, (16, 16, 16), (1, 1), tmp_path, 1)
    run_sample(emu_samples, name, (16, 16, 16), (2, 2), tmp_path, 4)


@pytest.mark.parametrize("name", ONE_D)
def test_reference_1d_samples_emulated(emu_samples, name, tmp_path):
    run_sample(emu_samples, name, (16, 16, 16), (1, 1), tmp_path, 1, mo1=(0, 1, 2), mo2=(1, 0, 2), dim=0)


def test_reference_memord_sample_emulated(emu_samples, tmp_path):
    run_sample(emu_samples, "test3D_r2c_memord_c", (16, 16, 16), (1, 1), tmp_path, 1, mo1=(1, 0, 2), mo2=(2, 1, 0))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", THREE_D + ["test3D_r2c_memord_c"])
def test_reference_3d_samples_gpu(gpu, name, tmp_path):
    """the same unmodified programs linked to the product library, on the device; 128^3 = config C1's grid"""
    bindir = os.path.join(SAMPLES, "_build")
    run_sample(bindir, name, (128, 128, 128), (1, 1), tmp_path, 1, mo1=(1, 0, 2), mo2=(2, 1, 0))
    if _ngpu() >= 4:
        run_sample(bindir, name, (128, 128, 128), (2, 2), tmp_path, 4, mo1=(1, 0, 2), mo2=(2, 1, 0))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ONE_D)
def test_reference_1d_samples_gpu(gpu, name, tmp_path):
    run_sample(os.path.join(SAMPLES, "_build"), name, (128, 128, 128), (1, 1), tmp_path, 1, mo1=(0, 1, 2), mo2=(1, 0, 2), dim=0)
