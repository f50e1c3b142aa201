"""The Fortran-callable boundary without a Fortran compiler (none in the image): the bind(C) names of fortran/*.f90 resolve in
the library, and the `_f` entry points (reference build/wrap.C:575-790, build/fwrap.f90:81-154) behave as a Fortran caller
expects -- every argument by reference, grids and plans as integer handles, idir 1-based -- on the CPU emulation, against the
oracle.  What stays untested is the compilation of the module source itself."""
import ctypes
import os
import re
from ctypes import byref, c_int

import numpy as np

from cases import RCC, half
from util import TOL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bind_c_names():
    names = []
    for f in ("p3dfft_plus_plus.f90", "fwrap.f90"):
        text = open(os.path.join(ROOT, "fortran", f)).read()
        names += re.findall(r"bind\s*\(\s*C\s*,\s*name\s*=\s*'([^']+)'\s*\)", text, flags=re.I)
    return names


def test_bind_c_names_resolve_in_the_library(pkg):
    names = bind_c_names()
    assert len([n for n in names if n.startswith("P3DFFT_")]) == 44 and len([n for n in names if n.startswith("p3dfft_")]) == 16, names
    lib = pkg.Library()
    for n in names:
        if n.startswith("P3DFFT_"):
            ctypes.c_int.in_dll(lib.dll, n)  # type-ID global (reference build/init.C:84-89)
        else:
            assert hasattr(lib.dll, n), n


def _i3(v):
    return (c_int * 3)(*v)


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def test_fortran_style_calls_by_reference(emu, orc):
    """sample/FORTRAN/test3D_r2c.f90's call sequence, by reference: processor grid, two data grids (handles + Ldims +
    GlobStart returned through arguments), 3D type, plan, exec; then exec_3Dderiv and compute_deriv with 1-based idir, and
    the 1D API"""
    d = emu.dll
    n, n2 = (16, 12, 10), half((16, 12, 10))
    pg = d.p3dfft_init_proc_grid_f(_i3([1, 1, 1]), byref(c_int(0)))  # MPI_Comm_f2c of the world handle
    g1, g2 = c_int(-1), c_int(-1)
    l1, s1, l2, s2 = _i3([0] * 3), _i3([0] * 3), _i3([0] * 3), _i3([0] * 3)
    d.p3dfft_init_data_grid_f(byref(g1), l1, s1, _i3(n), byref(c_int(-1)), byref(c_int(pg)), _i3([0, 1, 2]), _i3([0, 1, 2]))
    d.p3dfft_init_data_grid_f(byref(g2), l2, s2, _i3(n2), byref(c_int(0)), byref(c_int(pg)), _i3([1, 2, 0]), _i3([1, 2, 0]))
    assert g1.value >= 0 and g2.value >= 0 and g1.value != g2.value
    og1 = orc.OGrid(n, [0, 1, 2], [0, 1, 2], [1, 1, 1], 0)
    og2 = orc.OGrid(n2, [1, 2, 0], [1, 2, 0], [1, 1, 1], 0, 0)
    assert list(l1) == og1.Ldims and list(s1) == og1.GlobStart and list(l2) == og2.Ldims and list(s2) == og2.GlobStart
    # the same request again returns the same handle (find_grid, wrap.C:727)
    g1b = c_int(-1)
    d.p3dfft_init_data_grid_f(byref(g1b), l1, s1, _i3(n), byref(c_int(-1)), byref(c_int(pg)), _i3([0, 1, 2]), _i3([0, 1, 2]))
    assert g1b.value == g1.value
    t3 = c_int(-1)
    d.p3dfft_init_3Dtype_f(byref(t3), _i3([emu.types[t] for t in RCC]))
    plan = c_int(-1)
    d.p3dfft_plan_3Dtrans_f(byref(plan), byref(g1), byref(g2), byref(t3))
    assert plan.value >= 0
    G = orc.random_field(n, key=31)
    a = np.ascontiguousarray(orc.local_of(G, og1), dtype=np.float64)
    out = np.full(og2.storage_shape(), np.nan, dtype=np.complex128)
    d.p3dfft_exec_3Dtrans_double_f(byref(plan), _dptr(a), _dptr(out), byref(c_int(0)))
    want = orc.local_of(orc.transform_global(G, RCC, n2), og2)
    assert orc.rel_l2(out, want) < TOL[8]
    for idir_f in (1, 2, 3):  # Fortran counts dimensions from 1 (wrap.C:773-781)
        od = np.full(og2.storage_shape(), np.nan, dtype=np.complex128)
        d.p3dfft_exec_3Dderiv_double_f(byref(plan), _dptr(a), _dptr(od), byref(c_int(idir_f)), byref(c_int(0)))
        wd = orc.local_of(orc.transform_global(G, RCC, n2, deriv_dim=idir_f - 1), og2)
        assert orc.rel_l2(od, wd) < TOL[8], idir_f
        cd = np.full(og2.storage_shape(), np.nan, dtype=np.complex128)
        d.p3dfft_compute_deriv_double_f(_dptr(out), _dptr(cd), byref(g2), byref(c_int(idir_f)))
        assert orc.rel_l2(cd, orc.compute_deriv_local(out, og2, idir_f - 1, mode="reference")) < 1e-15, idir_f
    # 1D API: C2C along y (0-based dimension, as in the reference's wrapper: wrap.C:584-622)
    gc = c_int(-1)
    d.p3dfft_init_data_grid_f(byref(gc), l1, s1, _i3(n), byref(c_int(-1)), byref(c_int(pg)), _i3([0, 1, 2]), _i3([1, 0, 2]))
    p1 = c_int(-1)
    d.p3dfft_plan_1Dtrans_f(byref(p1), byref(gc), byref(gc), byref(c_int(emu.types["CFFT_FORWARD_D"])), byref(c_int(1)))
    assert p1.value >= 0
    ogc = orc.OGrid(n, [0, 1, 2], [1, 0, 2], [1, 1, 1], 0)
    Gc = orc.random_field(n, complex_=True, key=32)
    ac = np.ascontiguousarray(orc.local_of(Gc, ogc), dtype=np.complex128)
    oc = np.full(ogc.storage_shape(), np.nan, dtype=np.complex128)
    d.p3dfft_exec_1Dtrans_double_f(byref(p1), _dptr(ac), _dptr(oc), byref(c_int(0)))
    assert orc.rel_l2(oc, orc.local_of(orc.transform_1d(Gc, "fwd", 1), ogc)) < TOL[8]
    imo = _i3([9, 9, 9])
    d.p3dfft_inv_mo(_i3([1, 2, 0]), imo)  # bound directly by the module (fp3dfft++mod.f90:87): arrays by reference on both sides
    assert list(imo) == [2, 0, 1]
