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


def test_headline_plan_layouts_plan_only():
    """planner regression at full size without a device (P3DFFT_B200_PLAN_ONLY): 1024^3 R2C / C2R stage list, the storage
    orders of the intermediates (contiguous loads, transposed stores kept within few pages) and the 128-byte row padding of
    library-owned half-complex arrays"""
    import json
    import subprocess
    import sys
    code = r'''
import sys, json
sys.path.insert(0, %r)
import __graft_entry__ as ge
lib = ge.load_package().Library().setup()
n = (1024, 1024, 1024)
pg = lib.init_proc_grid([1, 1, 1])
g1 = lib.init_data_grid(n, -1, pg, [0, 1, 2], [0, 1, 2])
g2 = lib.init_data_grid((513, 1024, 1024), 0, pg, [1, 2, 0], [1, 2, 0])
out = []
for t, a, b in ((["R2CFFT_D", "CFFT_FORWARD_D", "CFFT_FORWARD_D"], g1, g2), (["C2RFFT_D", "CFFT_BACKWARD_D", "CFFT_BACKWARD_D"], g2, g1)):
    out.append(lib.describe_plan3d(lib.plan_3Dtrans(a, b, lib.init_3Dtype(t))))
print("PLANS" + json.dumps(out))
''' % ROOT
    env = dict(os.environ, P3DFFT_B200_PLAN_ONLY="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    line = [x for x in r.stdout.splitlines() if x.startswith("PLANS")]
    assert line, (r.stdout[-2000:], r.stderr[-2000:])
    fwd, bwd = json.loads(line[0][5:])
    assert fwd["ok"] and bwd["ok"]
    assert [s["dim"] for s in fwd["stages"]] == [0, 1, 2] and [s["dim"] for s in bwd["stages"]] == [2, 1, 0]
    assert [s["kind"] for s in fwd["stages"]] == [3, 1, 1] and [s["kind"] for s in bwd["stages"]] == [2, 2, 4]
    # forward: [z][y][x] -> [z][kx][y] -> [kx][ky][z] -> [ky][kx][kz]
    assert [s["out_mo"] for s in fwd["stages"]] == [[1, 0, 2], [2, 1, 0], [1, 2, 0]]
    # backward: [ky][kx][kz] -> [kx][z][ky] -> [z][y][kx] -> [z][y][x]
    assert [s["out_mo"] for s in bwd["stages"]] == [[2, 0, 1], [0, 1, 2], [0, 1, 2]]
    # every stage reads whole pencils (unit stride along its transform dimension)
    for p in (fwd, bwd):
        for s in p["stages"]:
            assert s["in_stride"][s["dim"]] == 1, s
    # the intermediate [z][y][kx] array of the backward transform has 513-element rows: padded to 520 (128-byte multiple)
    assert bwd["stages"][2]["in_stride"] == [1, 520, 520 * 1024]
    assert bwd["stages"][1]["segs"][0]["os_d"] == 520
    # user-visible arrays stay dense
    assert fwd["stages"][0]["in_stride"] == [1, 1024, 1024 * 1024] and bwd["stages"][0]["in_stride"] == [1024, 513 * 1024, 1]


def _bench_dry_run(args, nranks=1, timeout=900):
    """bench.py's own logic (configs, synthetic field, parity block, roofline bookkeeping, host-array legs) on the CPU-thread
    emulation build with tiny grids: P3DFFT_BENCH_DRYRUN=1 prints {"dry_run": true, "would_print": {...}} instead of a metric line"""
    import json
    import subprocess
    import sys
    env = dict(os.environ, P3DFFT_BENCH_DRYRUN="1")
    env.pop("P3DFFT_B200_PLAN_ONLY", None)
    if nranks > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1",
               "--master-port", str(29700 + os.getpid() % 200), os.path.join(ROOT, "bench.py"), "--gpus", str(nranks)] + args
    else:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py")] + args
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=timeout, cwd=ROOT)
    lines = [x for x in out.stdout.splitlines() if x.startswith("{")]
    assert out.returncode == 0 and lines, (out.stdout[-2000:], out.stderr[-3000:])
    d = json.loads(lines[-1])
    assert d["dry_run"] and d["emulated_library"]
    return d["would_print"]


@pytest.mark.parametrize("cfg,edge", [("c3", 32), ("c2", 16), ("c4", 16)])
def test_bench_logic_dry_run(emu, cfg, edge):
    d = _bench_dry_run(["--config", cfg, "--edge", str(edge), "--steps", "1", "--warmup", "1", "--e2e-steps", "1"])
    tol = 1e-5 if cfg == "c2" else 1e-12
    assert d["config"]["config"] == cfg and d["parity"]["roundtrip_rel_l2"] < tol
    assert d["parity"]["oracle"]["fwd_rel_l2"] < tol and d["parity"]["oracle"]["bwd_rel_l2"] < tol
    assert d["parity"]["known_answer"]["max_abs_err_over_peak"] < 10 * tol
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and "pageable" in d["e2e"] and "pageable_registered" in d["e2e"]
    assert d["roofline"]["bound"] == "hbm" and len(d["roofline"]["stages"]) == 6 and d["gpu_launches"] > 0
    if cfg == "c4":
        assert d["parity"]["oracle"]["fwd_deriv_rel_l2"] < tol


def test_bench_logic_dry_run_world_size_2(emu):
    """the N > 1 path of bench.py under torch.distributed.run with gloo (world_size 2): per-rank blocks of the global Philox
    field, per-rank known-answer and oracle checks reduced over the ranks, NVLink roofline bookkeeping of the exchange stages"""
    d = _bench_dry_run(["--config", "c3", "--edge", "32", "--steps", "1", "--warmup", "1", "--e2e-steps", "1"], nranks=2)
    assert d["n_gpus"] == 2 and d["config"]["proc_grid"] == [1, 1, 2]
    assert d["parity"]["roundtrip_rel_l2"] < 1e-12 and d["parity"]["oracle"]["fwd_rel_l2"] < 1e-12
    assert d["parity"]["known_answer"]["max_abs_err_over_peak"] < 1e-11
    assert d["roofline"]["bound"] == "nvlink" and sum(1 for s in d["roofline"]["stages"] if s["exchange"]) == 2
