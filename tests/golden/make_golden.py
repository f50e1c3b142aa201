#!/usr/bin/env python
"""make_golden.py -- produce the golden vectors under tests/golden/ by running the REFERENCE's own host code.

Run in the authoring container (needs /root/reference): `make -C oracle` builds oracle/_ref/ref_driver = the
reference's unmodified build/{init,templ,exec,deriv,wrap}.C linked against a single-host MPI subset and the plain-C
FFTW restatement (oracle/cfft).  For every case below each rank's input is the oracle's slice of a seeded global
random field; the reference plans and executes the transform; its per-rank output arrays, Ldims and GlobStart are
stored in tests/golden/<name>.npz.  tests/test_oracle.py then pins the NumPy oracle against these files, and the
GPU parity tests compare the CUDA path with them directly.  Nothing here runs on the GPU box.

    python tests/golden/make_golden.py            # regenerate everything
"""
import itertools
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import p3dfft_oracle as orc  # noqa: E402

DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
PERMS = list(itertools.permutations((0, 1, 2)))
RCC = ["R2CFFT_D", "CFFT_FORWARD_D", "CFFT_FORWARD_D"]
CCR = ["C2RFFT_D", "CFFT_BACKWARD_D", "CFFT_BACKWARD_D"]
XP = dict(dmap1=[0, 1, 2], mo1=[0, 1, 2], dmap2=[1, 2, 0], mo2=[1, 2, 0])


def half(n, d=0):
    n = list(n)
    n[d] = n[d] // 2 + 1
    return n


def fwd(name, n, pd, types=RCC, **kw):
    c = dict(name=name, mode="3d", types=types, procdims=pd, gdims1=list(n), gdims2=half(n), cs1=-1, cs2=0, idir=-1, **XP)
    c.update(kw)
    return c


def bwd(name, n, pd, types=CCR, **kw):
    c = dict(name=name, mode="3d", types=types, procdims=pd, gdims1=half(n), gdims2=list(n), cs1=0, cs2=-1, idir=-1,
             dmap1=XP["dmap2"], mo1=XP["mo2"], dmap2=XP["dmap1"], mo2=XP["mo1"])
    c.update(kw)
    return c


def cases():
    cs = []
    n = (16, 12, 10)
    # BASELINE config 1 shape: 2x2 pencil grid on 4 ranks, X-pencil -> Z-pencil, uneven split 9 = 4|5
    cs += [fwd("c1_fwd_2x2", n, [1, 2, 2]), bwd("c1_bwd_2x2", n, [1, 2, 2])]
    cs += [fwd("slab4_fwd", n, [1, 1, 4]), bwd("slab4_bwd", n, [1, 1, 4]), fwd("row4_fwd", n, [1, 4, 1])]
    cs += [fwd("uneven3_fwd", (14, 7, 11), [1, 1, 3]), bwd("uneven3_bwd", (14, 7, 11), [1, 1, 3])]
    for idir in (0, 1, 2):
        cs.append(fwd(f"deriv{idir}_2x2", n, [1, 2, 2], idir=idir))
    cs.append(fwd("single_fwd_2x2", n, [1, 2, 2], types=["R2CFFT_S", "CFFT_FORWARD_S", "CFFT_FORWARD_S"]))
    cs.append(dict(name="c2c_fwd_2x2", mode="3d", types=["CFFT_FORWARD_D"] * 3, procdims=[1, 2, 2], gdims1=list(n), gdims2=list(n),
                   cs1=-1, cs2=-1, idir=-1, **XP))
    cs.append(dict(name="c2c_bwd_1", mode="3d", types=["CFFT_BACKWARD_D"] * 3, procdims=[1, 1, 1], gdims1=list(n), gdims2=list(n),
                   cs1=-1, cs2=-1, idir=-1, dmap1=[0, 1, 2], mo1=[0, 1, 2], dmap2=[0, 1, 2], mo2=[0, 1, 2]))
    # config 4 shape: R2C(x) C2C(y) DCT-I(z) on complex data, non-default order, with and without derivative
    t4 = ["R2CFFT_D", "CFFT_FORWARD_D", "DCT1_COMPLEX_D"]
    cs += [fwd("c4_dct_2x2", (16, 8, 9), [1, 2, 2], types=t4), fwd("c4_dct_deriv1_2x2", (16, 8, 9), [1, 2, 2], types=t4, idir=1)]
    # sample/C/test2D+empty.c
    cs.append(fwd("empty_mid_2x2", n, [1, 2, 2], types=["R2CFFT_D", "EMPTY_TYPE_DOUBLE_COMPLEX", "CFFT_FORWARD_D"]))
    # all 36 memory-order pairs on one rank (sample/C/test3D_r2c_memord.c), forward; backward for the diagonal
    m = (8, 6, 10)
    for mo1 in PERMS:
        for mo2 in PERMS:
            tag = "".join(map(str, mo1)) + "_" + "".join(map(str, mo2))
            cs.append(fwd(f"memord_fwd_{tag}", m, [1, 1, 1], mo1=list(mo1), mo2=list(mo2), dmap2=[0, 1, 2]))
        tag = "".join(map(str, mo1))
        cs.append(bwd(f"memord_bwd_{tag}", m, [1, 1, 1], mo1=list(mo1), mo2=[0, 1, 2], dmap1=[0, 1, 2]))
    # memory orders on the 2x2 grid
    for mo1, mo2 in (([1, 0, 2], [2, 1, 0]), ([2, 0, 1], [0, 2, 1])):
        tag = "".join(map(str, mo1)) + "_" + "".join(map(str, mo2))
        cs.append(fwd(f"memord_2x2_fwd_{tag}", n, [1, 2, 2], mo1=mo1, mo2=mo2))
    # 1D transplan API (sample/C++/test_transplan.C, test1D_cos.C, test1D_cos_complex.C, test1D_sin.C)
    g = (9, 6, 5)
    for ty, dim in (("R2CFFT_D", 0), ("R2CFFT_D", 1), ("R2CFFT_D", 2), ("DCT1_REAL_D", 0), ("DCT1_REAL_D", 2), ("DCT1_COMPLEX_D", 1),
                    ("DST1_REAL_D", 0), ("DST1_COMPLEX_D", 2), ("DCT2_REAL_D", 0), ("DCT3_REAL_D", 1), ("DST2_REAL_D", 2),
                    ("DST3_REAL_D", 0), ("DST4_REAL_D", 1), ("DCT4_REAL_D", 0), ("CFFT_FORWARD_D", 1), ("DCT1_REAL_S", 0)):
        # Only pairs where the transform dimension is the unit-stride one in the input or in the output.  For the
        # remaining pairs the reference's two-step path (build/exec.C:594-621) reads the input with the strides of the
        # swap0 order instead of mo1 and returns values that are not the transform of the input along `dim`
        # (established by brute force, see DESIGN.md "Reference defects"); there is nothing meaningful to pin there.
        lead = {0: [0, 1, 2], 1: [1, 0, 2], 2: [1, 2, 0]}[dim]   # dim is storage rank 0
        other = {0: [1, 2, 0], 1: [0, 1, 2], 2: [0, 1, 2]}[dim]  # dim is not storage rank 0
        for mo1, mo2 in ((lead, lead), (lead, other), (other, lead)):
            kind = orc.type_info(ty)[0]
            g2 = half(g, dim) if kind == "r2c" else list(g)
            tag = f"{ty}_d{dim}_" + "".join(map(str, mo1)) + "_" + "".join(map(str, mo2))
            cs.append(dict(name=f"t1d_{tag}", mode="1d", types=[ty], dim=dim, procdims=[1, 1, 1], gdims1=list(g), gdims2=g2, cs1=-1,
                           cs2=dim if kind == "r2c" else -1, idir=-1, dmap1=[0, 1, 2], mo1=mo1, dmap2=[0, 1, 2], mo2=mo2))
    # stand-alone compute_deriv (build/deriv.C): every storage order and direction on one rank, plus a distributed grid
    for mo in PERMS:
        for idir in (0, 1, 2):
            tag = "".join(map(str, mo)) + f"_i{idir}"
            cs.append(dict(name=f"cderiv_{tag}", mode="deriv", types=["EMPTY_TYPE_DOUBLE_COMPLEX"], procdims=[1, 1, 1], gdims1=[9, 6, 5],
                           gdims2=[9, 6, 5], cs1=0, cs2=0, idir=idir, dmap1=[0, 1, 2], mo1=list(mo), dmap2=[0, 1, 2], mo2=list(mo)))
    for idir in (0, 1, 2):
        cs.append(dict(name=f"cderiv_2x2_i{idir}", mode="deriv", types=["EMPTY_TYPE_DOUBLE_COMPLEX"], procdims=[1, 2, 2], gdims1=[9, 12, 10],
                       gdims2=[9, 12, 10], cs1=0, cs2=0, idir=idir, dmap1=[1, 2, 0], mo1=[1, 2, 0], dmap2=[1, 2, 0], mo2=[1, 2, 0]))
    return cs


STRIDE = 61  # kernel-size cases keep every 61st element of each rank's output (+ its L2 norm) instead of the whole array


def large_cases():
    """cases at the sizes the production kernels serve (M >= 64: TMA-fed power-of-two, r2r and mixed-radix kernels, exchange
    segments) -- the small cases above pin the index mapping, these pin the kernels themselves to the reference's output"""
    cs = []
    n = (128, 64, 64)
    cs += [fwd("k_fwd_128x64x64", n, [1, 1, 1]), bwd("k_bwd_128x64x64", n, [1, 1, 1]),
           fwd("k_fwd_128x64x64_2x2", n, [1, 2, 2]), bwd("k_bwd_128x64x64_2x2", n, [1, 2, 2]),
           fwd("k_fwd_128x64x64_slab4", n, [1, 1, 4]), fwd("k_fwd_deriv1_128x64x64_slab2", n, [1, 1, 2], idir=1),
           fwd("k_fwd_single_128x64x64", n, [1, 1, 1], types=["R2CFFT_S", "CFFT_FORWARD_S", "CFFT_FORWARD_S"])]
    # user arrays stored with y or z fastest (test3D_r2c_memord.c's matrix at kernel size): the strided first stage takes the
    # tensor-load kernel (pow2_tload.cuh); 1 rank and 2x2
    for mo1 in ([1, 0, 2], [2, 1, 0], [1, 2, 0], [2, 0, 1]):
        tag = "".join(map(str, mo1))
        cs.append(fwd(f"k_fwd_128x64x64_mo{tag}", n, [1, 1, 1], mo1=mo1))
    cs.append(fwd("k_fwd_128x64x64_mo102_2x2", n, [1, 2, 2], mo1=[1, 0, 2]))
    cs.append(fwd("k_fwd_single_128x64x64_mo210", n, [1, 1, 1], types=["R2CFFT_S", "CFFT_FORWARD_S", "CFFT_FORWARD_S"], mo1=[2, 1, 0]))
    cs.append(fwd("k_fwd_deriv0_128x64x64_mo120", n, [1, 1, 1], mo1=[1, 2, 0], idir=0))
    cs.append(fwd("k_fwd_2048x4x3", (2048, 4, 3), [1, 1, 1]))   # 1024-point complex core with the symmetric last pass
    cs.append(bwd("k_bwd_2048x4x3", (2048, 4, 3), [1, 1, 1]))
    cs.append(fwd("k_fwd_768x6x4", (768, 6, 4), [1, 1, 1]))     # mixed radix: 768-point real = 3 x 128 core
    t4 = ["R2CFFT_D", "CFFT_FORWARD_D", "DCT1_COMPLEX_D"]
    cs.append(fwd("k_c4_dct_64x16x129", (64, 16, 129), [1, 1, 1], types=t4))  # DCT-I of 2^k + 1 points on padded rows
    cs.append(fwd("k_c4_dct_deriv0_64x16x129", (64, 16, 129), [1, 1, 1], types=t4, idir=0))
    for ty, g, dim in (("CFFT_FORWARD_D", (1024, 4, 3), 0), ("CFFT_BACKWARD_D", (4096, 2, 2), 0), ("CFFT_FORWARD_S", (512, 4, 3), 0),
                       ("CFFT_FORWARD_D", (768, 4, 3), 0), ("CFFT_BACKWARD_D", (640, 4, 3), 0), ("CFFT_FORWARD_D", (896, 3, 2), 0),
                       ("R2CFFT_D", (1536, 4, 3), 0), ("DCT1_COMPLEX_D", (513, 4, 3), 0), ("DST1_COMPLEX_D", (255, 4, 3), 0),
                       ("DCT2_REAL_D", (512, 4, 3), 0), ("DCT3_COMPLEX_D", (256, 4, 3), 0), ("DST2_REAL_D", (256, 4, 3), 0),
                       ("DST3_REAL_D", (128, 4, 3), 0), ("DST4_REAL_D", (128, 4, 3), 0), ("CFFT_FORWARD_D", (4, 3, 1024), 2)):
        kind = orc.type_info(ty)[0]
        lead = {0: [0, 1, 2], 2: [1, 2, 0]}[dim]
        other = {0: [1, 2, 0], 2: [0, 1, 2]}[dim]
        for mo1, mo2 in ((lead, lead), (lead, other)):
            g2 = half(g, dim) if kind == "r2c" else list(g)
            tag = f"{ty}_{g[dim]}_d{dim}_" + "".join(map(str, mo1)) + "_" + "".join(map(str, mo2))
            cs.append(dict(name=f"k_t1d_{tag}", mode="1d", types=[ty], dim=dim, procdims=[1, 1, 1], gdims1=list(g), gdims2=g2, cs1=-1,
                           cs2=dim if kind == "r2c" else -1, idir=-1, dmap1=[0, 1, 2], mo1=mo1, dmap2=[0, 1, 2], mo2=mo2))
    for ty, g, dim, mo1, mo2 in (("CFFT_FORWARD_D", (8, 256, 6), 1, [0, 1, 2], [1, 0, 2]), ("CFFT_BACKWARD_D", (6, 4, 1024), 2, [0, 1, 2], [2, 1, 0]),
                                 ("R2CFFT_D", (512, 16, 3), 0, [1, 0, 2], [0, 1, 2]), ("CFFT_FORWARD_S", (512, 16, 2), 0, [2, 0, 1], [0, 1, 2])):
        kind = orc.type_info(ty)[0]
        g2 = half(g, dim) if kind == "r2c" else list(g)
        tag = f"{ty}_{g[dim]}_d{dim}_" + "".join(map(str, mo1)) + "_" + "".join(map(str, mo2))
        cs.append(dict(name=f"k_t1d_strided_{tag}", mode="1d", types=[ty], dim=dim, procdims=[1, 1, 1], gdims1=list(g), gdims2=g2, cs1=-1,
                       cs2=dim if kind == "r2c" else -1, idir=-1, dmap1=[0, 1, 2], mo1=mo1, dmap2=[0, 1, 2], mo2=mo2))
    return cs


def case_types(c):
    """(dt_in, dt_out, prec) and the numpy dtypes of a case"""
    if c["mode"] == "deriv":
        return 2, 2, 8
    kinds = [orc.type_info(t)[0] for t in c["types"]]
    prec = orc.type_info(c["types"][0])[3]
    if c["mode"] == "1d":
        _, d1, d2, _ = orc.type_info(c["types"][0])
        return d1, d2, prec
    dt_in = 1 if "r2c" in kinds else 2
    dt_out = 1 if "c2r" in kinds else 2
    if all(k == "empty" for k in kinds):
        dt_in = dt_out = orc.type_info(c["types"][0])[1]
    return dt_in, dt_out, prec


def np_dtype(dt, prec):
    return {(1, 8): np.float64, (2, 8): np.complex128, (1, 4): np.float32, (2, 4): np.complex64}[(dt, prec)]


def global_input(c):
    """the seeded global logical input array of a case (double precision)"""
    dt_in, _, _ = case_types(c)
    if c["mode"] == "3d" and any(orc.type_info(t)[0] == "c2r" for t in c["types"]):
        kinds = [orc.type_info(t)[0] for t in c["types"]]
        f = ["R2CFFT_D" if k == "c2r" else ("CFFT_FORWARD_D" if k == "bwd" else "EMPTY_TYPE_DOUBLE_COMPLEX") for k in kinds]
        return orc.transform_global(orc.random_field(c["gdims2"]), f)
    return orc.random_field(c["gdims1"], complex_=(dt_in == 2))


def oracle_output(c, G, rank):
    """what the oracle says rank `rank` must hold after the case"""
    pd = c["procdims"]
    og1 = orc.OGrid(c["gdims1"], c["dmap1"], c["mo1"], pd, rank, c["cs1"])
    og2 = orc.OGrid(c["gdims2"], c["dmap2"], c["mo2"], pd, rank, c["cs2"])
    if c["mode"] == "3d":
        want = orc.local_of(orc.transform_global(G, c["types"], c["gdims2"], deriv_dim=c["idir"]), og2)
    elif c["mode"] == "1d":
        want = orc.local_of(orc.transform_1d(G, orc.type_info(c["types"][0])[0], c["dim"]), og2)
    else:
        want = orc.compute_deriv_local(orc.local_of(G, og1), og1, c["idir"], mode="reference")
    return og1, og2, want


def run_case(c, td):
    dt_in, dt_out, prec = case_types(c)
    pd = c["procdims"]
    nr = pd[0] * pd[1] * pd[2]
    G = global_input(c)
    inp, outp = os.path.join(td, "in"), os.path.join(td, "out")
    for r in range(nr):
        og1 = orc.OGrid(c["gdims1"], c["dmap1"], c["mo1"], pd, r, c["cs1"])
        orc.local_of(G, og1).astype(np_dtype(dt_in, prec)).tofile(f"{inp}.{r}.bin")
    g = lambda k, cs: " ".join(map(str, list(c["gdims" + k]) + [cs] + list(c["dmap" + k]) + list(c["mo" + k])))  # noqa: E731
    with open(os.path.join(td, "case"), "w") as f:
        f.write(f"mode {c['mode']}\nprocdims {pd[0]} {pd[1]} {pd[2]}\ngrid1 {g('1', c['cs1'])}\ngrid2 {g('2', c['cs2'])}\n")
        if c["mode"] == "3d":
            f.write("types " + " ".join(c["types"]) + "\n")
        elif c["mode"] == "1d":
            f.write(f"type {c['types'][0]}\ndim {c['dim']}\n")
        f.write(f"idir {c['idir']}\nprec {prec}\ndtout {dt_out}\now 0\nin {inp}\nout {outp}\n")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mpirun.py"), "-np", str(nr), DRIVER, os.path.join(td, "case")],
                         capture_output=True, text=True, timeout=300)
    if res.returncode != 0:
        raise RuntimeError(f"reference driver failed for {c['name']}:\n{res.stdout[-2000:]}\n{res.stderr[-2000:]}")
    data = {}
    worst = 0.0
    for r in range(nr):
        og1, og2, want = oracle_output(c, G, r)
        meta = np.loadtxt(f"{outp}.{r}.meta", dtype=np.int64)
        out = np.fromfile(f"{outp}.{r}.bin", dtype=np_dtype(dt_out, prec))
        data[f"out_{r}"] = out
        data[f"meta_{r}"] = meta
        shape = og2.storage_shape() if c["mode"] != "deriv" else og1.storage_shape()
        err = orc.rel_l2(out.reshape(shape), want) if out.size else 0.0
        worst = max(worst, err)
        geo = list(meta) == og1.Ldims + og1.GlobStart + og2.Ldims + og2.GlobStart
        if not geo:
            print(f"  !! geometry mismatch rank {r}: ref {list(meta)} oracle {og1.Ldims + og1.GlobStart + og2.Ldims + og2.GlobStart}")
    data["case"] = np.array(json.dumps(c))
    return data, worst


def main():
    if not os.path.exists(DRIVER):
        sys.exit("oracle/_ref/ref_driver missing: run `make -C oracle` in the authoring container (needs /root/reference)")
    only = sys.argv[1:]
    bundle = {}
    index = []
    for c in cases():
        if only and not any(o in c["name"] for o in only):
            continue
        with tempfile.TemporaryDirectory() as td:
            try:
                data, worst = run_case(c, td)
            except Exception as e:  # the reference rejects or mishandles some requests; record and move on
                print(f"{c['name']:40s} REFERENCE FAILED: {str(e)[:300]}")
                continue
        tol = 1e-5 if case_types(c)[2] == 4 else 1e-12
        print(f"{c['name']:40s} oracle vs reference rel-L2 {worst:.2e} {'ok' if worst < tol else 'MISMATCH'}")
        for k, v in data.items():
            bundle[f"{c['name']}/{k}"] = v
        index.append(c["name"])
    if not only:
        np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **bundle)
        with open(os.path.join(HERE, "index.json"), "w") as f:
            json.dump(index, f, indent=0)
        print(f"{len(index)} cases -> tests/golden/reference_golden.npz")
    # kernel-size cases: every STRIDE-th element of each rank's output and the output's L2 norm
    bundle, index = {}, []
    for c in large_cases():
        if only and not any(o in c["name"] for o in only):
            continue
        with tempfile.TemporaryDirectory() as td:
            try:
                data, worst = run_case(c, td)
            except Exception as e:
                print(f"{c['name']:44s} REFERENCE FAILED: {str(e)[:300]}")
                continue
        tol = 1e-5 if case_types(c)[2] == 4 else 1e-12
        print(f"{c['name']:44s} oracle vs reference rel-L2 {worst:.2e} {'ok' if worst < tol else 'MISMATCH'}")
        for k, v in data.items():
            if k.startswith("out_"):
                bundle[f"{c['name']}/sub_{k[4:]}"] = v[::STRIDE].copy()
                bundle[f"{c['name']}/norm_{k[4:]}"] = np.array([np.linalg.norm(v.astype(np.complex128 if np.iscomplexobj(v) else np.float64)), v.size])
            else:
                bundle[f"{c['name']}/{k}"] = v
        index.append(c["name"])
    if not only:
        np.savez_compressed(os.path.join(HERE, "reference_golden_kernels.npz"), **bundle)
        with open(os.path.join(HERE, "index_kernels.json"), "w") as f:
            json.dump(index, f, indent=0)
        print(f"{len(index)} kernel-size cases -> tests/golden/reference_golden_kernels.npz (every {STRIDE}th element + norms)")


if __name__ == "__main__":
    main()
