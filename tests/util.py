"""helpers shared by the parity tests: run one transform through the C ABI and compare with the oracle"""
import numpy as np

TOL = {8: 1e-12, 4: 1e-5}  # relative L2 error bounds stated by BASELINE.json:north_star


def np_dtype(dt, prec):
    if dt == 2:
        return np.complex128 if prec == 8 else np.complex64
    return np.float64 if prec == 8 else np.float32


def run_3d(lib, orc, gdims1, gdims2, types, mo1, mo2, dmap1=(0, 1, 2), dmap2=(0, 1, 2), cs1=-1, cs2=-1, procdims=(1, 1, 1),
           rank=0, G=None, deriv=-1, inplace=False, key=20240, return_all=False):
    """single-rank 3D transform of a random field; returns rel-L2 error against the oracle"""
    k0, dt_in, _, prec = orc.type_info(types[orc.transform_order(types)[0]])
    _, _, dt_out, _ = orc.type_info(types[orc.transform_order(types)[-1]])
    for t in types:  # empty types keep the datatype
        kk, a, b, _ = orc.type_info(t)
    has_r2c = any(orc.type_info(t)[0] == "r2c" for t in types)
    has_c2r = any(orc.type_info(t)[0] == "c2r" for t in types)
    if has_r2c:
        dt_in, dt_out = 1, 2
    elif has_c2r:
        dt_in, dt_out = 2, 1
    single = prec == 4
    pg = lib.init_proc_grid(list(procdims))
    g1 = lib.init_data_grid(gdims1, cs1, pg, list(dmap1), list(mo1))
    g2 = lib.init_data_grid(gdims2, cs2, pg, list(dmap2), list(mo2))
    t3 = lib.init_3Dtype(list(types))
    plan = lib.plan_3Dtrans(g1, g2, t3)
    desc = lib.describe_plan3d(plan)
    assert desc["ok"], desc
    if G is None:
        if has_c2r:  # Hermitian-consistent input: forward transform of a real field
            fwd = [("R2CFFT_D" if orc.type_info(t)[0] == "c2r" else
                    ("CFFT_FORWARD_D" if orc.type_info(t)[0] == "bwd" else "EMPTY_TYPE_DOUBLE_COMPLEX")) for t in types]
            G = orc.transform_global(orc.random_field(gdims2, key=key), fwd)
        else:
            G = orc.random_field(gdims1, complex_=(dt_in == 2), key=key)
    og1 = orc.OGrid(gdims1, dmap1, mo1, procdims, rank, cs1)
    og2 = orc.OGrid(gdims2, dmap2, mo2, procdims, rank, cs2)
    a = orc.local_of(G, og1).astype(np_dtype(dt_in, prec))
    want = orc.local_of(orc.transform_global(G, types, gdims2, deriv_dim=deriv), og2)
    if inplace:
        n1 = a.size * (2 if dt_in == 2 else 1)
        n2 = int(np.prod(og2.storage_shape())) * (2 if dt_out == 2 else 1)
        buf = np.zeros(max(n1, n2), dtype=np.float32 if single else np.float64)
        buf[:n1] = a.view(buf.dtype).ravel()
        if deriv >= 0:
            lib.exec_3Dderiv(plan, buf, buf, deriv, 1, single=single)
        else:
            lib.exec_3Dtrans(plan, buf, buf, 1, single=single)
        out = buf[:n2].view(np_dtype(dt_out, prec)).reshape(og2.storage_shape())
    else:
        out = np.full(og2.storage_shape(), np.nan, dtype=np_dtype(dt_out, prec))
        a0 = a.copy()
        if deriv >= 0:
            lib.exec_3Dderiv(plan, a, out, deriv, 0, single=single)
        else:
            lib.exec_3Dtrans(plan, a, out, 0, single=single)
        assert np.array_equal(a, a0), "input was modified although OW == 0"
    lib.free_data_grid(g1)
    lib.free_data_grid(g2)
    err = orc.rel_l2(out, want)
    if return_all:
        return err, out, want, desc
    return err


def run_1d(lib, orc, gdims, type_name, dim, mo1, mo2, key=7, expect_variant=None):
    """transplan-style 1D transform (p3dfft_plan_1Dtrans) on one rank"""
    kind, dt_in, dt_out, prec = orc.type_info(type_name)
    single = prec == 4
    gd2 = list(gdims)
    if kind == "r2c":
        gd2[dim] = gdims[dim] // 2 + 1
    pg = lib.init_proc_grid([1, 1, 1])
    g1 = lib.init_data_grid(gdims, -1, pg, [0, 1, 2], list(mo1))
    g2 = lib.init_data_grid(gd2, dim if kind == "r2c" else -1, pg, [0, 1, 2], list(mo2))
    plan = lib.plan_1Dtrans(g1, g2, type_name, dim)
    desc = lib.describe_plan1d(plan)
    assert desc["ok"], desc
    if expect_variant:
        assert any(s["variant"].startswith(expect_variant) for s in desc["stages"]), [s["variant"] for s in desc["stages"]]
    G = orc.random_field(gdims, complex_=(dt_in == 2), key=key)
    og1 = orc.OGrid(gdims, [0, 1, 2], mo1, [1, 1, 1], 0)
    og2 = orc.OGrid(gd2, [0, 1, 2], mo2, [1, 1, 1], 0)
    a = orc.local_of(G, og1).astype(np_dtype(dt_in, prec))
    out = np.full(og2.storage_shape(), np.nan, dtype=np_dtype(dt_out, prec))
    lib.exec_1Dtrans(plan, a, out, 0, single=single)
    want = orc.local_of(orc.transform_1d(G, kind, dim), og2)
    lib.free_data_grid(g1)
    lib.free_data_grid(g2)
    return orc.rel_l2(out, want)


# ------------------------------------------------------------------------------------------------ golden vectors
def golden():
    """(index, npz, make_golden module) of tests/golden/reference_golden.npz (produced by the reference's host code)"""
    import importlib.util
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    with open(os.path.join(here, "index.json")) as f:
        index = json.load(f)
    return index, np.load(os.path.join(here, "reference_golden.npz")), mg


def golden_case(z, name):
    import json
    return json.loads(str(z[name + "/case"]))


def exec_case(lib, orc, mg, c, rank):
    """run one golden-format case on this rank through the C ABI; returns (output array, grid1, grid2 geometry lists)"""
    dt_in, dt_out, prec = mg.case_types(c)
    single = prec == 4
    pd = c["procdims"]
    pg = lib.init_proc_grid(pd)
    g1 = lib.init_data_grid(c["gdims1"], c["cs1"], pg, c["dmap1"], c["mo1"])
    g2 = lib.init_data_grid(c["gdims2"], c["cs2"], pg, c["dmap2"], c["mo2"])
    og1 = orc.OGrid(c["gdims1"], c["dmap1"], c["mo1"], pd, rank, c["cs1"])
    og2 = orc.OGrid(c["gdims2"], c["dmap2"], c["mo2"], pd, rank, c["cs2"])
    geo = list(g1.contents.Ldims) + list(g1.contents.GlobStart) + list(g2.contents.Ldims) + list(g2.contents.GlobStart)
    a = orc.local_of(mg.global_input(c), og1).astype(np_dtype(dt_in, prec))
    if c["mode"] == "deriv":
        out = np.full(og1.storage_shape(), np.nan, dtype=np_dtype(dt_out, prec))
        lib.compute_deriv(a, out, g1, c["idir"], single=single)
    else:
        out = np.full(og2.storage_shape(), np.nan, dtype=np_dtype(dt_out, prec))
        if c["mode"] == "3d":
            plan = lib.plan_3Dtrans(g1, g2, lib.init_3Dtype(c["types"]))
            assert lib.describe_plan3d(plan)["ok"]
            if c["idir"] >= 0:
                lib.exec_3Dderiv(plan, a, out, c["idir"], 0, single=single)
            else:
                lib.exec_3Dtrans(plan, a, out, 0, single=single)
        else:
            plan = lib.plan_1Dtrans(g1, g2, c["types"][0], c["dim"])
            assert lib.describe_plan1d(plan)["ok"]
            lib.exec_1Dtrans(plan, a, out, 0, single=single)
    lib.free_data_grid(g1)
    lib.free_data_grid(g2)
    return out, geo


def check_golden(lib, orc, names, rank=0, world=1):
    """run the named golden cases whose rank count equals `world`; compare with the reference's arrays"""
    index, z, mg = golden()
    done = 0
    for name in names or index:
        c = golden_case(z, name)
        pd = c["procdims"]
        if pd[0] * pd[1] * pd[2] != world:
            continue
        out, geo = exec_case(lib, orc, mg, c, rank)
        ref = z[f"{name}/out_{rank}"]
        assert geo == list(z[f"{name}/meta_{rank}"]), (name, geo)
        prec = mg.case_types(c)[2]
        if ref.size:
            err = orc.rel_l2(out.ravel(), ref)
            assert err < TOL[prec], (name, rank, err)
        done += 1
    return done


# ------------------------------------------------------------------------------------------------ kernel-size golden vectors
def golden_kernels():
    """(index, npz, make_golden module) of tests/golden/reference_golden_kernels.npz: cases at the sizes the production kernels
    serve (M >= 64), produced by the reference's host code; every STRIDE-th element of each rank's output + its L2 norm"""
    import importlib.util
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    with open(os.path.join(here, "index_kernels.json")) as f:
        index = json.load(f)
    return index, np.load(os.path.join(here, "reference_golden_kernels.npz")), mg


def compare_with_kernel_golden(orc, z, mg, name, rank, out, prec):
    """out: this rank's full output array; compares the sampled elements and the norm with the reference's"""
    sub = z[f"{name}/sub_{rank}"]
    norm, size = z[f"{name}/norm_{rank}"]
    flat = np.asarray(out).ravel()
    assert flat.size == int(size), (name, rank, flat.size, size)
    if flat.size == 0:
        return 0.0
    err = orc.rel_l2(flat[::mg.STRIDE], sub)
    mine = np.linalg.norm(flat.astype(np.complex128 if np.iscomplexobj(flat) else np.float64))
    assert abs(mine - norm) <= (1e-5 if prec == 4 else 1e-12) * max(norm, 1e-300), (name, rank, mine, norm)
    return err


def check_golden_kernels(lib, orc, names=None, rank=0, world=1, expect_variants=None):
    """run the kernel-size golden cases whose rank count equals `world` through the C ABI and compare with the reference"""
    index, z, mg = golden_kernels()
    done = 0
    for name in names or index:
        c = golden_case(z, name)
        pd = c["procdims"]
        if pd[0] * pd[1] * pd[2] != world:
            continue
        out, geo = exec_case(lib, orc, mg, c, rank)
        assert geo == list(z[f"{name}/meta_{rank}"]), (name, geo)
        prec = mg.case_types(c)[2]
        err = compare_with_kernel_golden(orc, z, mg, name, rank, out, prec)
        assert err < TOL[prec], (name, rank, err)
        done += 1
    return done
