#!/usr/bin/env python
"""bench.py -- headline benchmark: 1024^3 double-precision R2C + C2R round trip through the P3DFFT++ C API.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 1024] [--grid slab|pencil]
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One step = one forward (R2C, X-pencil/slab layout -> Z-pencil layout) plus one backward (C2R) transform of the
same synthetic random field.  metric = GFLOP/s with the nominal 5*N*log2(N) flops per 3D transform
(SURVEY.md section 8d).  Prints ONE JSON line on rank 0.

  value          arrays resident in HBM, device pointers through p3dfft_exec_3Dtrans_double, CUDA-event timed
  e2e            same calls with pinned HOST arrays: H2D of the input and D2H of the result inside the timed region
  roofline       slowest stage kernel: algorithmic bytes (its input array once + its output array once) / its
                 CUDA-event duration inside the timed region, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline   the CPU restatement (oracle, scipy pocketfft, all host cores) on a bounded sample of the workload
  --impl reference   times the reference's CPU path: oracle/_ref (reference host code on shims) when it was built,
                 else the oracle port; bounded sample per step
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def flops_3d(n):
    N = float(n[0]) * n[1] * n[2]
    return 5.0 * N * math.log2(N)


class ClockSampler(threading.Thread):
    """samples SM clocks, power and throttle reasons while the GPU is under load: NVML in-process every 2 ms (nvidia-smi
    every 100 ms when NVML is unavailable).  Samples are time-stamped so the ones inside the timed region can be told
    from the warm-up ones; a region shorter than the sampling period falls back to all samples taken under load."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []  # (t, sm_mhz, sm_max_mhz, power_w, [reasons])
        self.proc = None
        self.halt = threading.Event()
        self.ready = threading.Event()  # first sample taken (NVML initialisation can take longer than a short timed region)
        self.t0 = self.t1 = None  # timed region

    def _gpu_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip()]
            if self.gpu < len(ids) and ids[self.gpu].strip().isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self._gpu_index())
        smax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))
        while not self.halt.is_set():
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            try:
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                pw = 0.0
            self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), smax, pw,
                                 [n for n, b in bits if r & b]))
            self.ready.set()
            time.sleep(0.002)

    def _run_smi(self):
        self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                      "-i", str(self._gpu_index())], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                rs = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8])
                      if v.lower().startswith("active")]
                self.samples.append((time.perf_counter(), float(f[1]), float(f[2]), float(f[3]), rs))
                self.ready.set()
            except ValueError:
                continue

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            try:
                self._run_smi()
            except Exception:
                pass

    def stop(self):
        self.halt.set()
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        inside = [s for s in self.samples if self.t0 is not None and self.t0 <= s[0] <= self.t1]
        use, where = (inside, "timed region") if len(inside) >= 3 else (self.samples, "warm-up + timed region (region shorter than 3 samples)")
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(s[1] for s in use)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(s[2] for s in use), "reasons": sorted({r for s in use for r in s[4]}),
                "samples": len(sm), "sampled_during": where, "power_w_max": max(s[3] for s in use)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def choose_grid(nranks, mode):
    if nranks == 1:
        return [1, 1, 1]
    if mode == "pencil":
        p1 = 2 if nranks % 2 == 0 and nranks > 2 else 1
        return [1, p1, nranks // p1]
    return [1, 1, nranks]


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_sample_size(n):
    """bounded sample: the same transform on a cube whose edge is halved until it is <= 512 (<= 1/8 of the work)"""
    m = list(n)
    while max(m) > 512:
        m = [x // 2 for x in m]
    return tuple(m)


def cpu_port_roundtrip(n, reps=1):
    """oracle port: scipy pocketfft rfftn + irfftn (unnormalised semantics), all host cores; returns s per round trip"""
    import numpy as np
    import scipy.fft as sfft
    from oracle import p3dfft_oracle as orc
    cores = os.cpu_count() or 1
    x = orc.random_field(n)
    xt = np.ascontiguousarray(x.transpose(2, 1, 0))  # storage order of the X-pencil array: x fastest
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        X = sfft.rfftn(xt, axes=(2, 1, 0), workers=cores)
        y = sfft.irfftn(X, s=xt.shape, axes=(2, 1, 0), workers=cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    err = float(np.abs(y - xt).max())
    assert err < 1e-10, err
    return best, cores


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "test3D_r2c_ref")
    return p if os.path.exists(p) else None


def cpu_reference_roundtrip(n, reps=1):
    """the reference's own host code (oracle/_ref: build/*.C + sample/C++/test3D_r2c.C compiled unmodified against the
    mini-MPI and the plain-C FFT shim), one rank per host core; returns (s per round trip, ranks) or None"""
    exe = ref_binary()
    if not exe:
        return None
    import tempfile
    cores = os.cpu_count() or 1
    ranks = 1
    while ranks * 2 <= min(cores, 8):
        ranks *= 2
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "stdin"), "w") as f:
            f.write(f"{n[0]} {n[1]} {n[2]} 2 {reps}\n")
        p1 = 2 if ranks >= 4 else 1
        with open(os.path.join(td, "dims"), "w") as f:
            f.write(f"{p1} {ranks // p1}\n")
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mpirun.py"), "-np", str(ranks), exe], cwd=td,
                             capture_output=True, text=True, timeout=1800)
    t = None
    ok = "Results are correct" in out.stdout
    for line in out.stdout.splitlines():
        if line.startswith("Transform time"):
            t = float(line.split(":")[1].split()[2])  # max over ranks
    if t is None or not ok:
        sys.stderr.write("reference run failed:\n" + out.stdout[-2000:] + out.stderr[-2000:])
        return None
    return t, ranks


def run_reference_arm(args, n, rank, world):
    if rank != 0:
        return
    sample = cpu_sample_size(n)
    scale = flops_3d(sample) / flops_3d(n)
    times = []
    kind, cores = "port", os.cpu_count() or 1
    for i in range(args.warmup + args.steps):
        r = cpu_reference_roundtrip(sample, reps=1)
        if r is not None:
            dt, cores = r
            kind = "reference"
        else:
            dt, cores = cpu_port_roundtrip(sample, reps=1)
        if i >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    gflops = 2 * flops_3d(sample) / dt / 1e9
    desc = (f"{sample[0]}x{sample[1]}x{sample[2]} double R2C+C2R round trip per step = {scale:.4f} of the {n[0]}^3 workload's "
            f"flops; " + ("reference build/*.C + sample/C++/test3D_r2c.C on mini-MPI + plain-C FFT shim (not FFTW)"
                          if kind == "reference" else "oracle port: scipy pocketfft rfftn/irfftn"))
    line = {"impl": "reference", "metric": "3D R2C+C2R GFLOP/s (5N log2N)", "value": gflops, "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{n[0]}x{n[1]}x{n[2]} double R2C+C2R round trip", "sample": list(sample)},
            "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--grid", default=os.environ.get("P3DFFT_BENCH_GRID", "slab"), choices=["slab", "pencil"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--single", action="store_true", help="single precision (configs 2/5 style)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = (args.n, args.n, args.n)

    if args.impl == "reference":
        run_reference_arm(args, n, rank, world)
        return

    import numpy as np
    import torch
    import __graft_entry__ as ge

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    mod = ge.load_package()
    lib = mod.load().setup()
    assert lib.have_device()

    single = args.single
    rdt, cdt = (torch.float32, torch.complex64) if single else (torch.float64, torch.complex128)
    rb, cb = (4, 8) if single else (8, 16)
    S = "S" if single else "D"
    pdims = choose_grid(world, args.grid)
    pg = lib.init_proc_grid(pdims)
    nh = (n[0] // 2 + 1, n[1], n[2])
    g1 = lib.init_data_grid(n, -1, pg, [0, 1, 2], [0, 1, 2])
    g2 = lib.init_data_grid(nh, 0, pg, [1, 2, 0], [1, 2, 0])
    pf = lib.plan_3Dtrans(g1, g2, lib.init_3Dtype([f"R2CFFT_{S}", f"CFFT_FORWARD_{S}", f"CFFT_FORWARD_{S}"]))
    pb = lib.plan_3Dtrans(g2, g1, lib.init_3Dtype([f"C2RFFT_{S}", f"CFFT_BACKWARD_{S}", f"CFFT_BACKWARD_{S}"]))
    dfw, dbw = lib.describe_plan3d(pf), lib.describe_plan3d(pb)
    assert dfw["ok"] and dbw["ok"], (dfw, dbw)
    n1 = int(np.prod(g1.contents.Ldims[:]))
    n2 = int(np.prod(g2.contents.Ldims[:]))

    stream = torch.cuda.current_stream()
    lib.set_stream(stream.cuda_stream)
    gen = torch.Generator(device="cuda").manual_seed(20240 + rank)
    x = torch.randn(n1, device="cuda", dtype=rdt, generator=gen)
    X = torch.empty(n2, device="cuda", dtype=cdt)
    y = torch.empty(n1, device="cuda", dtype=rdt)

    def step_device():
        lib.exec_3Dtrans(pf, x, X, 0, single=single)
        lib.exec_3Dtrans(pb, X, y, 1, single=single)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- correctness guard inside the bench: the round trip must return N * input
    step_device()
    torch.cuda.synchronize()
    N = float(n[0]) * n[1] * n[2]
    rt_err = float((torch.linalg.vector_norm(y / N - x) / torch.linalg.vector_norm(x)).item())
    tol = 1e-5 if single else 1e-12
    assert rt_err < tol, f"round-trip error {rt_err}"

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.ready.wait(10)
    for _ in range(args.warmup):
        step_device()
    lib.enable_timers(True)
    barrier()
    sampler.t0 = time.perf_counter()
    l0 = lib.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    stage_f = np.zeros(len(dfw["stages"]))
    stage_b = np.zeros(len(dbw["stages"]))
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    sampler.t1 = time.perf_counter()
    ms = e0.elapsed_time(e1) / args.steps
    launches = lib.kernel_launches() - l0
    # per-stage durations averaged over the steps of the timed region (CUDA events recorded by the library on the launch
    # stream around every stage kernel of every exec; read back only now, so nothing synchronises inside the region)
    stage_f += np.array(lib.stage_times(pf))
    stage_b += np.array(lib.stage_times(pb))
    clocks = sampler.stop() if rank == 0 else None
    lib.enable_timers(False)
    ms = max_over_ranks(ms)
    gflops = 2 * flops_3d(n) / (ms * 1e-3) / 1e9

    # ---- roofline of the dominant (slowest) stage kernel
    peak, peak_src = measured_peaks()
    stages = []
    for d, t, name in ((dfw, stage_f, "fwd"), (dbw, stage_b, "bwd")):
        for i, s in enumerate(d["stages"]):
            bytes_in = int(np.prod(s["in_ldims"])) * s["dt_in"] * d["prec"]
            bytes_out = int(np.prod(s["out_ldims"])) * s["dt_out"] * d["prec"]
            st = {"stage": f"{name}{i}", "kind": s["kind"], "dim": s["dim"], "exchange": s["exchange"], "ms": float(t[i]),
                  "alg_bytes": bytes_in + bytes_out, "gbs": (bytes_in + bytes_out) / (t[i] * 1e-3) / 1e9 if t[i] > 0 else None,
                  "variant": s["variant"].split(" ")[0]}
            if s["exchange"]:  # bytes this rank stores into OTHER GPUs' buffers over NVLink (local_bytes*(p-1)/p)
                npen = s["in_ldims"][s["u"]] * s["in_ldims"][s["v"]]
                st["nvlink_bytes"] = sum((g["k1"] - g["k0"]) * npen for g in s["segs"] if g["peer_world"] != rank) * s["dt_out"] * d["prec"]
                st["nvlink_gbs"] = st["nvlink_bytes"] / (t[i] * 1e-3) / 1e9 if t[i] > 0 else None
            stages.append(st)
        # overlapped pairs (exchange stage + neighbouring local stage cut into chunks on two streams) are timed as a whole:
        # the pair's duration is booked on both members, the local one carries no bandwidth figure of its own
        base = len(stages) - len(d["stages"])
        for i, s in enumerate(d["stages"]):
            if s.get("pair"):
                a, b = stages[base + i], stages[base + i + 1]
                pair_ms = a["ms"] + b["ms"]
                for m, o in ((a, b), (b, a)):
                    m["ms"] = pair_ms
                    m["overlapped_with"] = o["stage"]
                    if m["exchange"]:
                        m["nvlink_gbs"] = m["nvlink_bytes"] / (pair_ms * 1e-3) / 1e9
                        m["gbs"] = None
                    else:
                        m["gbs"] = None
    local = [s for s in stages if not s["exchange"] and s["gbs"]]
    dom = max(local or stages, key=lambda s: s["ms"])  # dominant HBM-bound kernel (exchange stages: see "nvlink" below)
    roofline = {"bound": "hbm", "achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": dom["gbs"] / peak if dom["gbs"] else None,
                "traffic": None, "kernel": f"{dom['variant']} ({dom['stage']}, dim {dom['dim']})", "peak_source": peak_src,
                "alg_bytes_per_launch": dom["alg_bytes"], "ms_per_launch": dom["ms"], "stages": stages,
                "whole_step_hbm_frac": sum(s["alg_bytes"] for s in stages) / (ms * 1e-3) / 1e9 / peak}
    if not dom["gbs"]:
        roofline["achieved"] = roofline["frac"] = None
    xs = [s for s in stages if s["exchange"]]
    if xs:  # fused FFT + all-to-all stages: NVLink roofline, 900 GB/s per direction per GPU (rank 0's figures)
        xd = max(xs, key=lambda s: s["ms"])
        tot_b, tot_ms = sum(s["nvlink_bytes"] for s in xs), sum(s["ms"] for s in xs)  # (pair durations include the hidden local stage)
        roofline["nvlink"] = {"bound": "nvlink", "peak": 900.0, "unit": "GB/s", "peak_source": "NVLink 5 nominal, per direction per GPU",
                              "achieved": xd["nvlink_gbs"], "frac": xd["nvlink_gbs"] / 900.0 if xd["nvlink_gbs"] else None,
                              "kernel": f"{xd['variant']} ({xd['stage']}, dim {xd['dim']}, fused exchange)",
                              "bytes_per_launch": xd["nvlink_bytes"], "ms_per_launch": xd["ms"],
                              "all_exchanges": {"bytes": tot_b, "ms": tot_ms, "achieved": tot_b / (tot_ms * 1e-3) / 1e9, "frac": tot_b / (tot_ms * 1e-3) / 1e9 / 900.0},
                              "whole_step_frac_of_exchange_bound": (tot_b / 900e9) / (ms * 1e-3)}
    traffic_file = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            roofline["traffic"] = json.load(f).get(dom["variant"])

    # ---- end to end: pinned host arrays through the same C ABI calls
    e2e = None
    if not args.no_e2e:
        hx = torch.empty(n1, dtype=rdt).pin_memory()
        hX = torch.empty(n2, dtype=cdt).pin_memory()
        hx.copy_(x)
        torch.cuda.synchronize()

        def step_host():
            lib.exec_3Dtrans(pf, hx, hX, 0, single=single)   # H2D real field, 3 stages, D2H spectrum
            lib.exec_3Dtrans(pb, hX, hx, 1, single=single)   # H2D spectrum, 3 stages, D2H real field
        step_host()
        hx.div_(N)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            step_host()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        dt = max_over_ranks(dt)
        hb = n1 * rb + n2 * cb
        e2e = {"value": 2 * flops_3d(n) / dt / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": hb * world, "d2h_bytes_per_step": hb * world,
               "ms_per_step": dt * 1e3, "steps": args.e2e_steps, "host_memory": "pinned"}
        del hx, hX

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = cpu_sample_size(n)
        r = cpu_reference_roundtrip(sample, reps=1)
        if r is not None:
            dt, cores = r
            kind = "reference"
            what = "reference build/*.C + sample/C++/test3D_r2c.C on mini-MPI + plain-C FFT shim (not FFTW)"
        else:
            dt, cores = cpu_port_roundtrip(sample, reps=2)
            kind = "port"
            what = "oracle port: scipy pocketfft rfftn/irfftn"
        cpu = {"value": 2 * flops_3d(sample) / dt / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": kind,
               "sample": f"{sample[0]}x{sample[1]}x{sample[2]} double R2C+C2R round trip ({flops_3d(sample) / flops_3d(n):.4f} of the "
                         f"workload's flops), {dt:.2f} s; {what}"}

    if rank == 0:
        line = {"metric": "3D R2C+C2R GFLOP/s (5N log2N)", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32" if single else "f64", "data": "synthetic",
                "config": {"workload": f"{n[0]}x{n[1]}x{n[2]} {'single' if single else 'double'} R2C+C2R round trip "
                                       f"(X-pencil mo 012 -> Z-pencil mo 120)", "proc_grid": pdims,
                           "l2": "inputs larger than L2 (per-GPU arrays >= 1 GB vs 126 MB L2)", "roundtrip_rel_l2": rt_err},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    lib.free_data_grid(g1)
    lib.free_data_grid(g2)
    del x, X, y
    lib.cleanup()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
