#!/usr/bin/env python
"""fuzz_tload.py [seed] [cases] -- random single-rank cases aimed at the tensor-load kernel (pow2_tload.cuh) on the CPU
emulation: power-of-two transform lengths, random other extents (partial tiles, 16-byte stride rule hit or missed), random
storage orders on both sides, 1D and 3D plans, double / single, fused derivative.  Every output is checked against the
oracle; prints how many cases actually took the tensor-load kernel.  Development tool."""
import itertools
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
from cases import RCC, RCC_S, half  # noqa: E402
from util import TOL, run_1d, run_3d  # noqa: E402

rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
ncases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
perms = list(itertools.permutations((0, 1, 2)))
pkg, orc = ge.load_package(), ge.load_oracle()
emu = pkg.load(emulated=True).setup()
bad = used = 0
for it in range(ncases):
    m = rnd.choice([64, 128, 128, 256, 512, 1024])
    single = rnd.random() < 0.35
    dim = rnd.randint(0, 2)
    n = [rnd.randint(1, 40), rnd.randint(1, 40), rnd.randint(1, 12)]
    rnd.shuffle(n)
    mo1, mo2 = rnd.choice(perms), rnd.choice(perms)
    three_d = rnd.random() < 0.4
    if rnd.random() < 0.8:  # the transform dimension is not the input's unit-stride dimension; extents mostly even (16-byte strides)
        mo1 = rnd.choice([p for p in perms if p[0 if three_d else dim] != 0])
        n = [x + (x % 2) * (rnd.random() < 0.8) for x in n]
    try:
        if not three_d:
            t = rnd.choice(["CFFT_FORWARD", "CFFT_BACKWARD", "R2CFFT"]) + ("_S" if single else "_D")
            n[dim] = 2 * m if t.startswith("R2C") and m <= 512 else m
            pg = emu.init_proc_grid([1, 1, 1])
            g1 = emu.init_data_grid(n, -1, pg, [0, 1, 2], list(mo1))
            gd2 = list(n)
            if t.startswith("R2C"):
                gd2[dim] = n[dim] // 2 + 1
            g2 = emu.init_data_grid(gd2, dim if t.startswith("R2C") else -1, pg, [0, 1, 2], list(mo2))
            v = emu.describe_plan1d(emu.plan_1Dtrans(g1, g2, t, dim))["stages"][0]["variant"].split(" ")[0]
            err = run_1d(emu, orc, tuple(n), t, dim, mo1, mo2, key=it)
            what = ("1d", tuple(n), t, dim, mo1, mo2)
        else:
            n[0] = 2 * m if m <= 512 else m
            deriv = rnd.choice([-1, -1, 0, 1, 2])
            err, _, _, desc = run_3d(emu, orc, tuple(n), half(n), RCC_S if single else RCC, mo1, mo2, cs2=0, deriv=deriv, key=it, return_all=True)
            v = desc["stages"][0]["variant"].split(" ")[0]
            what = ("3d", tuple(n), "rcc_s" if single else "rcc", deriv, mo1, mo2)
        ok = err < TOL[4 if single else 8]
    except Exception as e:  # noqa: BLE001
        ok, err, v = False, repr(e), "?"
    used += v.startswith("tload")
    print(it, "ok" if ok else "FAIL", what, v, err, flush=True)
    bad += not ok
print("failures", bad, "tensor-load cases", used, "of", ncases)
sys.exit(1 if bad else 0)
