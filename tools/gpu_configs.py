#!/usr/bin/env python
"""gpu_configs.py -- times the single-GPU shapes of the BASELINE.json configs other than the headline (device-resident,
CUDA events per stage): C2 512^3 C2C single; C4 512x512x513 R2C.C2C.DCT-I double + derivative, order {1,2,0};
C5's per-GPU kernel shapes (2048-point single R2C/C2C on a 2048x2048x16 slab); 1024^3 single R2C+C2R."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def run(lib, name, gd1, gd2, types, mo1, mo2, single, cs2=-1, deriv=-1, reps=5, env=None):
    if ONLY and ONLY not in name:
        return
    if not ONLY and name.split(" ")[0] in ("TLOAD", "SMOOTH", "MIXAB"):  # A/B shapes run only when asked for
        return
    for k, v in (env or {}).items():  # kernel-ladder switches are read when the plan is created
        os.environ[k] = v
    pg = lib.init_proc_grid([1, 1, 1])
    g1 = lib.init_data_grid(gd1, -1, pg, [0, 1, 2], list(mo1))
    g2 = lib.init_data_grid(gd2, cs2, pg, [0, 1, 2], list(mo2))
    plan = lib.plan_3Dtrans(g1, g2, lib.init_3Dtype(types))
    desc = lib.describe_plan3d(plan)
    assert desc["ok"], desc
    for k in (env or {}):
        del os.environ[k]
    rdt = torch.float32 if single else torch.float64
    n1 = int(np.prod(gd1)) * desc["dt_in"]
    n2 = int(np.prod(gd2)) * desc["dt_out"]
    x = torch.randn(n1, device="cuda", dtype=rdt)
    y = torch.empty(n2, device="cuda", dtype=rdt)

    def step():
        if deriv >= 0:
            lib.exec_3Dderiv(plan, x, y, deriv, 0, single=single)
        else:
            lib.exec_3Dtrans(plan, x, y, 0, single=single)
    for _ in range(2):
        step()
    lib.enable_timers(True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = lib.stage_times(plan)
    lib.enable_timers(False)
    prec = 4 if single else 8
    parts = []
    for s, t in zip(desc["stages"], st):
        b = (int(np.prod(s["in_ldims"])) * s["dt_in"] + int(np.prod(s["out_ldims"])) * s["dt_out"]) * prec
        parts.append(f"{s['variant'].split(' ')[0]} {t:.3f}ms {b / t / 1e6:.0f}GB/s")
    print(f"{name}: {ms:.3f} ms | " + " | ".join(parts), flush=True)
    lib.free_data_grid(g1)
    lib.free_data_grid(g2)
    del x, y
    torch.cuda.empty_cache()


ONLY = sys.argv[1] if len(sys.argv) > 1 else ""  # run only the shapes whose label contains this string


def main():
    lib = ge.load_package().load().setup()
    lib.set_stream(torch.cuda.current_stream().cuda_stream)
    S3, D = ["CFFT_FORWARD_S"] * 3, "D"
    n = (512, 512, 512)
    run(lib, "C2 512^3 C2C single mo 012->012", n, n, S3, (0, 1, 2), (0, 1, 2), True)
    run(lib, "C4 512x512x513 R2C.C2C.DCT1 double mo 012->120", (512, 512, 513), (257, 512, 513), ["R2CFFT_D", "CFFT_FORWARD_D", "DCT1_COMPLEX_D"],
        (0, 1, 2), (1, 2, 0), False, cs2=0)
    run(lib, "C4 same + derivative idir=1", (512, 512, 513), (257, 512, 513), ["R2CFFT_D", "CFFT_FORWARD_D", "DCT1_COMPLEX_D"],
        (0, 1, 2), (1, 2, 0), False, cs2=0, deriv=1)
    run(lib, "C4 literal 512^3 (DCT-I of 512 points: L = 1022, Bluestein on 2048)", (512, 512, 512), (257, 512, 512),
        ["R2CFFT_D", "CFFT_FORWARD_D", "DCT1_COMPLEX_D"], (0, 1, 2), (1, 2, 0), False, cs2=0)
    run(lib, "1000x1000x200 C2C double (Bluestein on 2048 in x and y)", (1000, 1000, 200), (1000, 1000, 200), ["CFFT_FORWARD_D"] * 3,
        (0, 1, 2), (0, 1, 2), False)
    run(lib, "768^3 R2C double (3 x 2^k mixed-radix kernel)", (768, 768, 768), (385, 768, 768), ["R2CFFT_D", "CFFT_FORWARD_D", "CFFT_FORWARD_D"], (0, 1, 2), (1, 2, 0),
        False, cs2=0)
    run(lib, "768^3 C2R double", (385, 768, 768), (768, 768, 768), ["C2RFFT_D", "CFFT_BACKWARD_D", "CFFT_BACKWARD_D"], (1, 2, 0), (0, 1, 2), False)
    run(lib, "640^3 C2C double (5 x 128)", (640, 640, 640), (640, 640, 640), ["CFFT_FORWARD_D"] * 3, (0, 1, 2), (0, 1, 2), False)
    n = (1024, 1024, 1024)
    run(lib, "1024^3 R2C single mo 012->120", n, (513, 1024, 1024), ["R2CFFT_S", "CFFT_FORWARD_S", "CFFT_FORWARD_S"], (0, 1, 2), (1, 2, 0), True, cs2=0)
    run(lib, "1024^3 C2R single mo 120->012", (513, 1024, 1024), n, ["C2RFFT_S", "CFFT_BACKWARD_S", "CFFT_BACKWARD_S"], (1, 2, 0), (0, 1, 2), True)
    n = (2048, 2048, 64)
    run(lib, "C5 shape 2048x2048x64 R2C single (x,y stages at 2048 points)", n, (1025, 2048, 64), ["R2CFFT_S", "CFFT_FORWARD_S", "CFFT_FORWARD_S"],
        (0, 1, 2), (1, 2, 0), True, cs2=0)
    # user arrays stored with y fastest: the first stage reads x with a stride -- tensor-load kernel vs the plain-load kernel
    n = (1024, 1024, 1024)
    EC, ES = "EMPTY_TYPE_DOUBLE_COMPLEX", "EMPTY_TYPE_SINGLE_COMPLEX"
    for tag, env in (("tensor loads", None), ("plain loads", {"P3DFFT_B200_NO_TLOAD": "1"})):
        run(lib, f"TLOAD 1024^3 R2C double mo 102->120, {tag}", n, (513, 1024, 1024), ["R2CFFT_D", "CFFT_FORWARD_D", "CFFT_FORWARD_D"], (1, 0, 2),
            (1, 2, 0), False, cs2=0, env=env, reps=3)
        run(lib, f"TLOAD 1024x1024x512 C2C(y) double mo 012->012 (transposed stores), {tag}", (1024, 1024, 512), (1024, 1024, 512),
            [EC, "CFFT_FORWARD_D", EC], (0, 1, 2), (0, 1, 2), False, env=env, reps=3)
        run(lib, f"TLOAD 1024x1024x512 C2C(y) double mo 012->102 (contiguous stores), {tag}", (1024, 1024, 512), (1024, 1024, 512),
            [EC, "CFFT_FORWARD_D", EC], (0, 1, 2), (1, 0, 2), False, env=env, reps=3)
        run(lib, f"TLOAD 1024^3 R2C single mo 102->120, {tag}", n, (513, 1024, 1024), ["R2CFFT_S", "CFFT_FORWARD_S", "CFFT_FORWARD_S"], (1, 0, 2),
            (1, 2, 0), True, cs2=0, env=env, reps=3)
        run(lib, f"TLOAD 512^3 C2C(z) single mo 012->012, {tag}", (512, 512, 512), (512, 512, 512), [ES, ES, "CFFT_BACKWARD_S"], (0, 1, 2),
            (0, 1, 2), True, env=env, reps=3)
    # layout-search A/B (planner.cpp:stage_cost): a row-granular first stage followed by contiguous loads instead of a chain of
    # tensor-load stages; and the headline forward transform with the transposition moved to the load side of stages 2 and 3
    R = ["R2CFFT_D", "CFFT_FORWARD_D", "CFFT_FORWARD_D"]
    RS = ["R2CFFT_S", "CFFT_FORWARD_S", "CFFT_FORWARD_S"]
    n = (1024, 1024, 1024)
    run(lib, "TLOAD 1024^3 R2C double mo 102->120, row-row first stage", n, (513, 1024, 1024), R, (1, 0, 2), (1, 2, 0), False, cs2=0,
        env={"P3DFFT_B200_COST_ROWROW": "1.0"}, reps=3)
    run(lib, "TLOAD 1024^3 R2C double mo 012->120 (headline forward), default plan", n, (513, 1024, 1024), R, (0, 1, 2), (1, 2, 0), False, cs2=0, reps=3)
    run(lib, "TLOAD 1024^3 R2C double mo 012->120 (headline forward), load-side transposition", n, (513, 1024, 1024), R, (0, 1, 2), (1, 2, 0), False,
        cs2=0, env={"P3DFFT_B200_COST_TLOAD": "1.0"}, reps=3)
    run(lib, "TLOAD 1024^3 R2C single mo 012->120, default plan", n, (513, 1024, 1024), RS, (0, 1, 2), (1, 2, 0), True, cs2=0, reps=3)
    run(lib, "TLOAD 1024^3 R2C single mo 012->120, load-side transposition", n, (513, 1024, 1024), RS, (0, 1, 2), (1, 2, 0), True, cs2=0,
        env={"P3DFFT_B200_COST_TLOAD": "1.0"}, reps=3)
    # smooth lengths added last in round 2: 9 x 2^k, 15 x 2^k and the 64-point cores (3 x 64 ...); Bluestein before
    for tag, env in (("mixed-radix kernel", None), ("Bluestein", {"P3DFFT_B200_NO_MIXED": "1"})):
        run(lib, f"SMOOTH 1152^3 R2C double (9 x 128 / 9 x 64 cores), {tag}", (1152, 1152, 1152), (577, 1152, 1152), R, (0, 1, 2), (1, 2, 0), False,
            cs2=0, env=env, reps=2)
        run(lib, f"SMOOTH 960^3 C2C double (15 x 64), {tag}", (960, 960, 960), (960, 960, 960), ["CFFT_FORWARD_D"] * 3, (0, 1, 2), (0, 1, 2), False,
            env=env, reps=2)
        run(lib, f"SMOOTH 192^3 R2C double (3 x 64 / 3 x 32...), {tag}", (192, 192, 192), (97, 192, 192), R, (0, 1, 2), (1, 2, 0), False, cs2=0,
            env=env, reps=10)
    # A/B of the mixed-radix kernel's radix-Q step (by definition vs register butterflies): run once per library build
    C3 = ["CFFT_FORWARD_D"] * 3
    run(lib, "MIXAB 768^3 R2C double (3 x 128, 3 x 256)", (768, 768, 768), (385, 768, 768), R, (0, 1, 2), (1, 2, 0), False, cs2=0, reps=3)
    run(lib, "MIXAB 768^3 C2R double", (385, 768, 768), (768, 768, 768), ["C2RFFT_D", "CFFT_BACKWARD_D", "CFFT_BACKWARD_D"], (1, 2, 0), (0, 1, 2), False, reps=3)
    run(lib, "MIXAB 640^3 C2C double (5 x 128)", (640, 640, 640), (640, 640, 640), C3, (0, 1, 2), (0, 1, 2), False, reps=3)
    run(lib, "MIXAB 896^3 C2C double (7 x 128)", (896, 896, 896), (896, 896, 896), C3, (0, 1, 2), (0, 1, 2), False, reps=3)
    run(lib, "MIXAB 1152^3 R2C double (9 x 64, 9 x 128)", (1152, 1152, 1152), (577, 1152, 1152), R, (0, 1, 2), (1, 2, 0), False, cs2=0, reps=3)
    run(lib, "MIXAB 960^3 C2C double (15 x 64)", (960, 960, 960), (960, 960, 960), C3, (0, 1, 2), (0, 1, 2), False, reps=3)
    run(lib, "MIXAB 768^3 R2C single", (768, 768, 768), (385, 768, 768), RS, (0, 1, 2), (1, 2, 0), True, cs2=0, reps=3)
    n = (256, 256, 256)
    run(lib, "256^3 R2C double", n, (129, 256, 256), ["R2CFFT_D", "CFFT_FORWARD_D", "CFFT_FORWARD_D"], (0, 1, 2), (1, 2, 0), False, cs2=0, reps=20)
    n = (128, 128, 128)
    run(lib, "C1 shape 128^3 R2C double (1 rank)", n, (65, 128, 128), ["R2CFFT_D", "CFFT_FORWARD_D", "CFFT_FORWARD_D"], (0, 1, 2), (1, 2, 0), False, cs2=0, reps=50)
    lib.cleanup()


if __name__ == "__main__":
    main()
