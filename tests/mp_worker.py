"""Multi-rank worker (not collected by pytest): runs one 3D transform case on every rank of a mini-MPI world and
compares each rank's local output with the oracle's slice of the global transform.

Launched by tests/test_multirank.py through tools/mpirun.py (P3DFFT_RANK/...) or torch.distributed.run
(RANK/WORLD_SIZE/MASTER_PORT, the launcher bench.py runs under).  argv: <emu|gpu> <json case list>"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import __graft_entry__ as ge  # noqa: E402
from util import TOL, check_golden, check_golden_kernels, np_dtype  # noqa: E402


def main():
    mode = sys.argv[1]
    cases = [] if sys.argv[2] == "golden" else json.loads(sys.argv[2])
    rank = int(os.environ.get("P3DFFT_RANK", os.environ.get("RANK", "0")))
    world = int(os.environ.get("P3DFFT_NRANKS", os.environ.get("WORLD_SIZE", "1")))
    use_gloo = "--gloo" in sys.argv
    if use_gloo:
        import torch
        import torch.distributed as dist
        dist.init_process_group("gloo")
    mod = ge.load_package()
    orc = ge.load_oracle()
    lib = mod.load(emulated=(mode == "emu")).setup()
    worst = 0.0
    if sys.argv[2] == "golden":  # every golden case recorded on this many ranks, against the reference's own arrays
        n = check_golden(lib, orc, None, rank=rank, world=world)
        n += check_golden_kernels(lib, orc, None, rank=rank, world=world)  # 128 x 64 x 64 on 2 and 4 ranks (exchange segments)
        assert n > 0, "no golden case for this world size"
    for c in cases:
        types, pd = c["types"], c["procdims"]
        assert pd[0] * pd[1] * pd[2] == world, (pd, world)
        g1d, g2d = c["gdims1"], c["gdims2"]
        prec = orc.type_info(types[0])[3]
        single = prec == 4
        kinds = [orc.type_info(t)[0] for t in types]
        dt_in = 1 if "r2c" in kinds else 2
        dt_out = 1 if "c2r" in kinds else 2
        if all(k == "empty" for k in kinds):
            dt_in = dt_out = orc.type_info(types[0])[1]
        pg = lib.init_proc_grid(pd)
        g1 = lib.init_data_grid(g1d, c.get("cs1", -1), pg, c["dmap1"], c["mo1"])
        g2 = lib.init_data_grid(g2d, c.get("cs2", -1), pg, c["dmap2"], c["mo2"])
        plan = lib.plan_3Dtrans(g1, g2, lib.init_3Dtype(types))
        desc = lib.describe_plan3d(plan)
        assert desc["ok"], desc
        if os.environ.get("P3DFFT_TEST_EXPECT_PAIRS") and c.get("expect_pairs", True):
            assert any(s["pair"] for s in desc["stages"]), ("no overlapped pair planned", desc)
        if os.environ.get("P3DFFT_TEST_EXPECT_TRIPLE") and c.get("expect_triple", False):  # L -> X -> Z as three persistent kernels
            assert any(s.get("triple") for s in desc["stages"]), ("no L-X-Z triple planned", desc)
        if os.environ.get("P3DFFT_TEST_EXPECT_SYNC") and c.get("expect_sync", True):  # persistent pair kernels with tile-group flags
            assert any(s["pair_sync"] for s in desc["stages"]), ("pair without the tile-group form", desc)
        if c.get("expect_variant"):  # some stage runs the named kernel variant
            assert any(s["variant"].startswith(c["expect_variant"]) for s in desc["stages"]), (c["expect_variant"], [s["variant"] for s in desc["stages"]])
        og1 = orc.OGrid(g1d, c["dmap1"], c["mo1"], pd, rank, c.get("cs1", -1))
        og2 = orc.OGrid(g2d, c["dmap2"], c["mo2"], pd, rank, c.get("cs2", -1))
        assert list(g1.contents.Ldims) == og1.Ldims and list(g1.contents.GlobStart) == og1.GlobStart
        assert list(g2.contents.Ldims) == og2.Ldims and list(g2.contents.GlobStart) == og2.GlobStart
        if "c2r" in kinds:
            fwd = ["R2CFFT_D" if k == "c2r" else ("CFFT_FORWARD_D" if k == "bwd" else "EMPTY_TYPE_DOUBLE_COMPLEX") for k in kinds]
            G = orc.transform_global(orc.random_field(g2d), fwd)
        else:
            G = orc.random_field(g1d, complex_=(dt_in == 2))
        a = orc.local_of(G, og1).astype(np_dtype(dt_in, prec))
        deriv = c.get("deriv", -1)
        want = orc.local_of(orc.transform_global(G, types, g2d, deriv_dim=deriv), og2)
        out = np.full(og2.storage_shape(), np.nan, dtype=np_dtype(dt_out, prec))
        reps = c.get("reps", 2)  # run twice: the second pass reuses the peer buffers (barrier epochs)
        if c.get("inplace"):  # in == out with the overwrite flag: one buffer of max(size1, size2) (exec.C:110-113)
            rdt = np.float32 if single else np.float64
            n1, n2 = a.size * dt_in, out.size * dt_out
            buf = np.zeros(max(n1, n2, 1), dtype=rdt)
            for _ in range(reps):
                buf[:n1] = a.view(rdt).ravel()
                if deriv >= 0:
                    lib.exec_3Dderiv(plan, buf, buf, deriv, 1, single=single)
                else:
                    lib.exec_3Dtrans(plan, buf, buf, 1, single=single)
            out = buf[:n2].view(np_dtype(dt_out, prec)).reshape(og2.storage_shape()).copy()
            reps = 0
        for _ in range(reps):
            out[...] = np.nan
            if deriv >= 0:
                lib.exec_3Dderiv(plan, a, out, deriv, 0, single=single)
            else:
                lib.exec_3Dtrans(plan, a, out, 0, single=single)
        err = orc.rel_l2(out, want) if want.size else 0.0
        if all(k == "empty" for k in kinds):
            assert np.array_equal(out, want.astype(out.dtype)), f"rank {rank}: permutation not bit-exact in case {c}"
        assert err < TOL[prec], f"rank {rank}: rel-L2 {err} in case {c}\nplan {json.dumps(desc)[:1500]}"
        worst = max(worst, err)
        lib.free_data_grid(g1)
        lib.free_data_grid(g2)
    if use_gloo:
        t = torch.tensor([worst], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = float(t.item())
        dist.destroy_process_group()
    lib.cleanup()
    print(f"RANK {rank} OK worst {worst:.3e}", flush=True)


if __name__ == "__main__":
    main()
