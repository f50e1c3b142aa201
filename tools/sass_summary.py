#!/usr/bin/env python
"""sass_summary.py -- instruction mix of the production kernels from `cuobjdump -sass` of the built objects (no GPU needed):
which data-movement path a kernel uses (UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDG/STG, LDS/STS), its FP64 / FP32 math,
and that no tensor-core instruction occurs (HMMA / DMMA / UTCMMA: a radix-16 butterfly network is not a dense contraction).
Writes profiles/<tag>_sass_summary.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "p3dfft.3_b200", "lib", "obj")
PICK = [  # (object, mangled-name fragment, label)
    ("pipe_8_3.o", "pow2_pipe_kernelIdLi512ELi3ELi8ELi1E", "pow2_pipe_kernel<double,512,R2C,8,transposed>   1024^3 fwd0"),
    ("pipe_8_1.o", "pow2_pipe_kernelIdLi1024ELi1ELi8ELi1E", "pow2_pipe_kernel<double,1024,C2C fwd,8,transposed> 1024^3 fwd1"),
    ("pipe_8_1.o", "pow2_pipe_kernelIdLi1024ELi1ELi4ELi0E", "pow2_pipe_kernel<double,1024,C2C fwd,4,contiguous> 1024^3 fwd2"),
    ("pipe_8_1.o", "pow2_pipe_sync_kernelIdLi1024ELi1ELi8ELi1E", "pow2_pipe_sync_kernel<double,1024,C2C fwd,8,transposed> persistent exchange stage (flags)"),
    ("pipe_8_4.o", "pow2_pipe_kernelIdLi512ELi4ELi8ELi0E", "pow2_pipe_kernel<double,512,C2R,8,contiguous>  1024^3 bwd2"),
    ("pipe_8_5.o", "pow2_pipe_kernelIdLi1024ELi5ELi4ELi0E", "pow2_pipe_kernel<double,1024,DCT-I,4,contiguous> config C4"),
    ("pipe_8_13.o", "pow2_pipe_kernelIdLi1024ELi13ELi4ELi0E", "pow2_pipe_kernel<double,1024,r2r run-time kind,4,contiguous>"),
    ("pipe_4_1.o", "pow2_pipe_kernelIfLi2048ELi1ELi8ELi1E", "pow2_pipe_kernel<float,2048,C2C fwd,8,transposed> config C5"),
    ("mixed_8_1.o", "mixed_pipe_kernelIdLi256ELi3ELi1ELi8ELi1E", "mixed_pipe_kernel<double,256,3,C2C fwd,8,transposed> 768 points"),
    ("mixed_8_1.o", "mixed_pipe_kernelIdLi128ELi9ELi1ELi4ELi1E", "mixed_pipe_kernel<double,128,9,C2C fwd,4,transposed> 1152 points (register butterflies)"),
    ("tload_8_1.o", "pow2_tload_kernelIdLi1024ELi1ELi8ELi0E", "pow2_tload_kernel<double,1024,C2C fwd,8,contiguous> strided input (TMA tensor-map loads)"),
    ("tload_8_3.o", "pow2_tload_kernelIdLi512ELi3ELi16ELi0E", "pow2_tload_kernel<double,512,R2C,16,contiguous> strided real input"),
    ("fastcore_inst.o", "fastcore_stage_kernelIdLi2048ELi1E", "fastcore_stage_kernel<double,2048,bluestein>"),
]
OPS = ["UBLKCP", "SYNCS", "LDG", "STG", "LDS", "STS", "DFMA", "DADD", "DMUL", "FFMA", "FADD", "FMUL", "BAR", "ATOM", "RED", "MEMBAR",
       "HMMA", "DMMA", "IMMA", "UTCMMA", "UTMALDG", "UTMASTG", "LDL", "STL"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    lines = ["instruction counts (static, per kernel) from cuobjdump -sass of p3dfft.3_b200/lib/obj/*.o, sm_100a",
             "kernel | total | " + " ".join(OPS)]
    for obj, frag, label in PICK:
        path = os.path.join(OBJ, obj)
        if not os.path.exists(path):
            lines.append(f"{label}: {obj} not built")
            continue
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        cur, counts, total = False, collections.Counter(), 0
        for ln in sass.splitlines():
            if "Function :" in ln:
                cur = frag in ln
                continue
            if cur:
                m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
                if m:
                    total += 1
                    op = m.group(1)
                    for o in OPS:
                        if op.startswith(o):
                            counts[o] += 1
                            break
        lines.append(f"{label} | {total} | " + " ".join(f"{o}:{counts[o]}" for o in OPS if counts[o]))
    lines.append("(no HMMA / DMMA / IMMA / UTCMMA anywhere: no tensor-core instruction; UBLKCP = cp.async.bulk, the whole-pencil copies of "
                 "the unit-stride kernels; UTMALDG = cp.async.bulk.tensor, the tensor-map loads of pow2_tload_kernel for strided inputs; "
                 "no UTMASTG: stores go from registers)")
    out = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
