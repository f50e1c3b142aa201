"""The drop-in boundary: the shared library loads and exports every function include/*.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for h in ("Cwrap.h", "Fwrap.h", "p3dfft_b200.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
        text = re.sub(r"//[^\n]*", " ", text)
        for m in re.finditer(r"\b((?:p3dfft|p3dfftcu)_\w+|find_grid)\s*\(", text):
            names.add(m.group(1))
    return sorted(names)


def test_headers_declare_the_reference_entry_points():
    names = declared_functions()
    for must in ("p3dfft_setup", "p3dfft_cleanup", "p3dfft_init_proc_grid", "p3dfft_init_data_grid", "p3dfft_init_3Dtype",
                 "p3dfft_plan_3Dtrans", "p3dfft_plan_1Dtrans", "p3dfft_exec_3Dtrans_double", "p3dfft_exec_3Dtrans_single",
                 "p3dfft_exec_3Dderiv_double", "p3dfft_exec_3Dderiv_single", "p3dfft_compute_deriv_double",
                 "p3dfft_compute_deriv_single", "p3dfft_exec_1Dtrans_double", "p3dfft_plan_3Dtrans_f", "p3dfft_exec_3Dtrans_double_f",
                 "p3dfftcu_stage_exec"):
        assert must in names, must


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.Library()  # the product library (CUDA build); loading needs no GPU
    missing = [n for n in declared_functions() if not hasattr(lib.dll, n)]
    assert not missing, missing
    for n in pkg.TYPE_NAMES:  # the 44 type-ID globals (reference build/init.C:84-89)
        import ctypes
        ctypes.c_int.in_dll(lib.dll, "P3DFFT_" + n)


def test_exec_without_device_fails_loudly(pkg):
    """no CPU fallback: with no usable device the library reports it and exec aborts (checked via have_device only)"""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import __graft_entry__ as ge; m = ge.load_package(); "
            "l = m.Library().setup(); print('HAVE', l.have_device())" % ROOT)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert "HAVE False" in out.stdout, (out.stdout, out.stderr)
