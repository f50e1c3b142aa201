#!/usr/bin/env python
"""bench.py -- the BASELINE.json benchmarks through the P3DFFT++ C API (default: 1024^3 double R2C + C2R round trip).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c3|c4|c5] [--grid slab|pencil]
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Configs (BASELINE.json `configs`):
  c1  128^3 double R2C/C2R (the reference's own sample size; 2x2 pencil grid on 4 ranks)
  c2  512^3 single C2C forward + backward, default memory order
  c3  1024^3 double R2C/C2R, X-pencil mo 012 -> Z-pencil mo 120 (default; the metric's configuration)
  c4  512 x 512 x 513 double R2C(x) C2C(y) DCT-I(z) with the fused derivative (exec_3Dderiv), output order 120, + backward
  c5  2048^3 single R2C/C2R (8 GPUs: 17 GB per GPU; one GPU: 137 GB with the backward transform in place)
One step = one forward plus one backward transform of the same synthetic field.  metric = GFLOP/s with the nominal
5*N*log2(N) flops per 3D transform.  Prints ONE JSON line on rank 0.

  value          arrays resident in HBM, device pointers through p3dfft_exec_3Dtrans_*, CUDA-event timed, max over ranks
  e2e            the same calls with HOST arrays: H2D of the input and D2H of the result inside the timed region; pinned arrays
                 (headline) and, as e2e.pageable, ordinary pageable arrays as the reference's users pass them
  parity         before timing, at ANY N: the reference sample's known-answer check at the bench size (sine field -> +-N/8 i at
                 the modes (1, +-1, +-1), sample/C++/test3D_r2c.C:281-331; every rank checks its own block), the round trip,
                 and a comparison with the CPU oracle on a reduced grid with the SAME processor grid and overlap machinery.
                 The run fails when any of them is out of tolerance.
  roofline       the dominant stage: N = 1 the slowest stage kernel against the measured HBM peak; N > 1 the fused exchange
                 stage (with the local stages overlapped with it) against NVLink
  cpu_baseline   the reference's host code (oracle/_ref) or the oracle port on the host cores, bounded sample, at every N
  --impl reference   times the reference's CPU path on the same configuration (one launch, Nrep repetitions inside)
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

KEY = 20240  # globally indexed Philox field
# development only: dry-run this script's logic without a GPU on the CPU-thread emulation build of the library (tools/cuda_emu),
# with tiny grids (--n 32).  Prints {"dry_run": true, ...} instead of a metric line; the driver never sets this.
DRY = os.environ.get("P3DFFT_BENCH_DRYRUN") == "1"
DEV = "cpu" if DRY else "cuda"


def _dry_run_shims(torch):
    import types

    class _Stream:
        cuda_stream = 0

    class _Event:
        def __init__(self, enable_timing=False):
            self.t = 0.0

        def record(self, stream=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *a: None
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.current_stream = lambda *a: _Stream()
    torch.cuda.Event = _Event
    torch.cuda.empty_cache = lambda: None
    torch.cuda.get_device_properties = lambda *a: types.SimpleNamespace(total_memory=int(180e9))
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.pin_memory = lambda self, *a, **k: self


def flops_3d(n):
    N = float(n[0]) * n[1] * n[2]
    return 5.0 * N * math.log2(N)


CONFIGS = {
    "c1": dict(n=(128, 128, 128), single=False, kind="r2c", desc="128^3 double R2C+C2R round trip (X-pencil mo 012 -> Z-pencil mo 120)"),
    "c2": dict(n=(512, 512, 512), single=True, kind="c2c", desc="512^3 single C2C forward+backward round trip, default memory order"),
    "c3": dict(n=(1024, 1024, 1024), single=False, kind="r2c",
               desc="1024x1024x1024 double R2C+C2R round trip (X-pencil mo 012 -> Z-pencil mo 120)"),
    "c4": dict(n=(512, 512, 513), single=False, kind="cheb",
               desc="512x512x513 double R2C(x) C2C(y) DCT-I(z) with fused d/dx (exec_3Dderiv) + backward, output mo 120"),
    "c5": dict(n=(2048, 2048, 2048), single=True, kind="r2c",
               desc="2048^3 single R2C+C2R round trip (X-pencil mo 012 -> Z-pencil mo 120)"),
}


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


def choose_grid(nranks, mode, cfg):
    if nranks == 1:
        return [1, 1, 1]
    if cfg in ("c1", "c4") and nranks == 4:
        return [1, 2, 2]  # the configuration's own 2x2 pencil grid
    if mode == "pencil":
        p1 = 2 if nranks % 2 == 0 and nranks > 2 else 1
        return [1, p1, nranks // p1]
    return [1, 1, nranks]


def bind_to_gpu_numa_node(local_rank):
    """run this rank's host threads (and first-touch its host arrays) on the NUMA node of its GPU: with 8 ranks copying at once
    the host memory system, not PCIe, bounds the end-to-end leg when arrays sit on the far socket"""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = [x for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip()]
        idx = int(vis[local_rank]) if local_rank < len(vis) and vis[local_rank].isdigit() else local_rank
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"node": None}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed)}
    except Exception as e:  # sysfs / NVML not available: leave the affinity alone
        return {"node": None, "why": type(e).__name__}


# ------------------------------------------------------------------------------------------------ synthetic fields
def philox_plane(key, z, ny, nx):
    """plane z of the globally indexed random field: every decomposition sees the same values"""
    import numpy as np
    return np.random.Generator(np.random.Philox(key=[key, z])).standard_normal((ny, nx))


def fill_philox_block(host, start, ldims, gdims, complex_, workers):
    """host: numpy view [lz][ly][lx] (x fastest; real or complex) of the X-pencil block that starts at `start`"""
    from concurrent.futures import ThreadPoolExecutor

    def one(lz):
        z = start[2] + lz
        sl = (slice(start[1], start[1] + ldims[1]), slice(start[0], start[0] + ldims[0]))
        p = philox_plane(KEY, z, gdims[1], gdims[0])[sl]
        if complex_:
            host[lz].real = p
            host[lz].imag = philox_plane(KEY + 1, z, gdims[1], gdims[0])[sl]
        else:
            host[lz] = p

    with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:
        list(ex.map(one, range(ldims[2])))


def philox_global(gdims, complex_):
    """the same field as a global logical array [x][y][z] (oracle comparison on the reduced grid)"""
    import numpy as np
    planes = np.stack([philox_plane(KEY, z, gdims[1], gdims[0]) for z in range(gdims[2])])
    if complex_:
        planes = planes + 1j * np.stack([philox_plane(KEY + 1, z, gdims[1], gdims[0]) for z in range(gdims[2])])
    return np.ascontiguousarray(planes.transpose(2, 1, 0))


def logical_view(t, ldims, mo):
    """torch view of a local array with axes in LOGICAL order (i0, i1, i2); storage extent sd[mo[i]] = ldims[i], sd[0] fastest"""
    sd = [0, 0, 0]
    for i in range(3):
        sd[mo[i]] = ldims[i]
    return t.view(sd[2], sd[1], sd[0]).permute(2 - mo[0], 2 - mo[1], 2 - mo[2])


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_sample_size(n):
    """bounded sample: the same transform on a grid whose edges are halved until the largest is <= 512 (<= 1/8 of the work)"""
    m = list(n)
    while max(m) > 512:
        m = [max(x // 2, 1) for x in m]
    return tuple(m)


def cpu_port_roundtrip(n, kind, single, reps=1):
    """oracle port: scipy pocketfft forward + backward of the configuration's transform, all host cores; s per round trip"""
    import numpy as np
    import scipy.fft as sfft
    cores = os.cpu_count() or 1
    rng = np.random.Generator(np.random.Philox(key=KEY))
    shape = (n[2], n[1], n[0])  # storage order of the X-pencil array: x fastest
    rdt = np.float32 if single else np.float64
    x = rng.standard_normal(shape).astype(rdt)
    if kind == "c2c":
        x = (x + 1j * rng.standard_normal(shape)).astype(np.complex64 if single else np.complex128)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        if kind == "r2c":
            X = sfft.rfftn(x, axes=(2, 1, 0), workers=cores)
            y = sfft.irfftn(X, s=shape, axes=(2, 1, 0), workers=cores)
        elif kind == "c2c":
            y = sfft.ifftn(sfft.fftn(x, workers=cores), workers=cores)
        else:  # R2C(x) C2C(y) DCT-I(z) and back
            X = sfft.dct(sfft.fft(sfft.rfft(x, axis=2, workers=cores), axis=1, workers=cores), type=1, axis=0, workers=cores)
            y = sfft.irfft(sfft.ifft(sfft.dct(X, type=1, axis=0, workers=cores), axis=1, workers=cores), n=n[0], axis=2, workers=cores)
            y = y / (2 * (n[2] - 1))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    err = float(np.abs(y - x).max())
    assert err < (1e-3 if single else 1e-9), err
    return best, cores


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "test3D_r2c_ref")
    return p if os.path.exists(p) else None


def cpu_reference_roundtrip(n, nrep=1, timeout=3000):
    """the reference's own host code (oracle/_ref: build/*.C + sample/C++/test3D_r2c.C compiled unmodified against the
    mini-MPI and the plain-C FFT shim), one rank per host core (at most 8), ONE launch with `nrep` forward+backward
    repetitions inside (sample/C++/test3D_r2c.C:230-266: plans, first touch and the field set-up are outside its timer);
    returns (s per round trip, ranks) or None"""
    exe = ref_binary()
    if not exe:
        return None
    import tempfile
    cores = os.cpu_count() or 1
    ranks = 1
    while ranks * 2 <= min(cores, 8):
        ranks *= 2
    while ranks > 1 and (n[1] % ranks or n[2] % ranks):
        ranks //= 2
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "stdin"), "w") as f:
            f.write(f"{n[0]} {n[1]} {n[2]} 2 {nrep}\n")
        p1 = 2 if ranks >= 4 else 1
        with open(os.path.join(td, "dims"), "w") as f:
            f.write(f"{p1} {ranks // p1}\n")
        try:
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mpirun.py"), "-np", str(ranks), exe], cwd=td,
                                 capture_output=True, text=True, timeout=timeout)
        except subprocess.TimeoutExpired:
            sys.stderr.write("reference run timed out\n")
            return None
    t = None
    ok = "Results are correct" in out.stdout
    for line in out.stdout.splitlines():
        if line.startswith("Transform time"):
            t = float(line.split(":")[1].split()[2])  # max over ranks, averaged over the nrep repetitions
    if t is None or not ok:
        sys.stderr.write("reference run failed:\n" + out.stdout[-2000:] + out.stderr[-2000:])
        return None
    return t, ranks


REF_WHAT = "reference build/*.C + sample/C++/test3D_r2c.C on mini-MPI + plain-C FFT shim (NOT FFTW)"


def cpu_baseline_leg(cfg):
    """bounded CPU sample for the `cpu_baseline` object (about 10-30 s): the reference's host code when the configuration is
    the R2C sample it implements, else the oracle port"""
    c = CONFIGS[cfg]
    sample = cpu_sample_size(c["n"])
    r = cpu_reference_roundtrip(sample, nrep=2) if (c["kind"] == "r2c" and not c["single"]) else None
    if r is not None:
        dt, cores = r
        kind, what = "reference", REF_WHAT + ", 2 repetitions in one launch"
    else:
        dt, cores = cpu_port_roundtrip(sample, c["kind"], c["single"], reps=2)
        kind, what = "port", "oracle port: scipy pocketfft (the reference's sample program built under oracle/_ref covers double R2C)"
    return {"value": 2 * flops_3d(sample) / dt / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": kind,
            "sample": f"{sample[0]}x{sample[1]}x{sample[2]} forward+backward ({flops_3d(sample) / flops_3d(c['n']):.4f} of the "
                      f"workload's flops), {dt:.2f} s per round trip; {what}"}


def run_reference_arm(args, cfg, rank):
    """the reference's CPU implementation on THIS configuration's grid (same_config) when it fits the host and a few minutes:
    one launch, warmup+steps (at most 4) repetitions inside; otherwise a bounded sample"""
    if rank != 0:
        return
    c = CONFIGS[cfg]
    n = c["n"]
    kind, cores, same = "port", os.cpu_count() or 1, False
    use_ref = c["kind"] == "r2c" and not c["single"] and ref_binary()
    nrep = max(1, min(args.steps + args.warmup, 3))
    dt = None
    sample = n
    if use_ref:
        # 1024^3 double: 8 ranks x (3 arrays + exchange buffers) ~ 45 GB of host memory, ~20 s per round trip on 8 cores
        try:
            avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
        except (ValueError, OSError):
            avail = 0
        need = 6.0 * float(n[0]) * n[1] * n[2] * 8
        if avail < need or n[0] > 1024:
            sample = cpu_sample_size(n)
        r = cpu_reference_roundtrip(sample, nrep=nrep)
        if r is None and sample == n:
            sample = cpu_sample_size(n)
            r = cpu_reference_roundtrip(sample, nrep=nrep)
        if r is not None:
            dt, cores = r
            kind = "reference"
    if dt is None:
        sample = cpu_sample_size(n)
        dt, cores = cpu_port_roundtrip(sample, c["kind"], c["single"], reps=max(1, min(args.steps, 3)))
    same = tuple(sample) == tuple(n)
    gflops = 2 * flops_3d(sample) / dt / 1e9
    desc = (f"{sample[0]}x{sample[1]}x{sample[2]} forward+backward per step = {flops_3d(sample) / flops_3d(n):.4f} of the workload's "
            f"flops; " + (f"{REF_WHAT}, {nrep} repetitions inside one launch (the sample's own timer: plans, first touch "
                          f"and field set-up excluded)" if kind == "reference" else "oracle port: scipy pocketfft"))
    line = {"impl": "reference", "metric": "3D R2C+C2R GFLOP/s (5N log2N)", "value": gflops, "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if c["single"] else "f64", "data": "synthetic",
            "config": {"workload": c["desc"], "config": cfg, "sample": list(sample), "same_config": same, "repetitions_timed": nrep},
            "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class Problem:
    """grids, plans and step functions of one configuration on one processor grid"""

    def __init__(self, lib, cfg_kind, n, single, pdims, idir=-1):
        self.lib, self.kind, self.n, self.single, self.pdims, self.idir = lib, cfg_kind, tuple(n), single, list(pdims), idir
        S = "S" if single else "D"
        self.pg = lib.init_proc_grid(self.pdims)
        if cfg_kind == "c2c":
            self.n2 = self.n
            self.mo2, self.dmap2, cs2 = [0, 1, 2], [0, 1, 2], -1
            tf, tb = [f"CFFT_FORWARD_{S}"] * 3, [f"CFFT_BACKWARD_{S}"] * 3
            self.dt_in = 2
            self.norm = float(n[0]) * n[1] * n[2]
        else:
            self.n2 = (n[0] // 2 + 1, n[1], n[2])
            self.mo2, self.dmap2, cs2 = [1, 2, 0], [1, 2, 0], 0
            self.dt_in = 1
            if cfg_kind == "cheb":
                tf = [f"R2CFFT_{S}", f"CFFT_FORWARD_{S}", f"DCT1_COMPLEX_{S}"]
                tb = [f"C2RFFT_{S}", f"CFFT_BACKWARD_{S}", f"DCT1_COMPLEX_{S}"]
                self.norm = float(n[0]) * n[1] * 2 * (n[2] - 1)
            else:
                tf = [f"R2CFFT_{S}", f"CFFT_FORWARD_{S}", f"CFFT_FORWARD_{S}"]
                tb = [f"C2RFFT_{S}", f"CFFT_BACKWARD_{S}", f"CFFT_BACKWARD_{S}"]
                self.norm = float(n[0]) * n[1] * n[2]
        self.types_f, self.types_b = tf, tb
        self.g1 = lib.init_data_grid(self.n, -1, self.pg, [0, 1, 2], [0, 1, 2])
        self.g2 = lib.init_data_grid(self.n2, cs2, self.pg, self.dmap2, self.mo2)
        self.pf = lib.plan_3Dtrans(self.g1, self.g2, lib.init_3Dtype(tf))
        self.pb = lib.plan_3Dtrans(self.g2, self.g1, lib.init_3Dtype(tb))
        self.dfw, self.dbw = lib.describe_plan3d(self.pf), lib.describe_plan3d(self.pb)
        assert self.dfw["ok"] and self.dbw["ok"], (self.dfw, self.dbw)
        self.ld1, self.gs1 = list(self.g1.contents.Ldims), list(self.g1.contents.GlobStart)
        self.ld2, self.gs2 = list(self.g2.contents.Ldims), list(self.g2.contents.GlobStart)
        self.n1e = self.ld1[0] * self.ld1[1] * self.ld1[2]  # local elements of the input array (real, or complex for C2C)
        self.n2e = self.ld2[0] * self.ld2[1] * self.ld2[2]  # local complex elements of the spectrum

    def forward(self, a, b, ow=0, deriv=False):
        if deriv and self.idir >= 0:
            self.lib.exec_3Dderiv(self.pf, a, b, self.idir, ow, single=self.single)
        else:
            self.lib.exec_3Dtrans(self.pf, a, b, ow, single=self.single)

    def backward(self, a, b, ow=1):
        self.lib.exec_3Dtrans(self.pb, a, b, ow, single=self.single)

    def free(self):
        self.lib.free_data_grid(self.g1)
        self.lib.free_data_grid(self.g2)

    def overlap_summary(self):
        out = {}
        for name, d in (("fwd", self.dfw), ("bwd", self.dbw)):
            st = d["stages"]
            out[name] = {"pair": any(s.get("pair") for s in st), "persistent_flag_kernels": any(s.get("pair_sync") for s in st),
                         "triple": any(s.get("triple") for s in st)}
        return out


def known_answer_check(torch, prob, x, X, rdt, cdt, tol):
    """the reference sample's own check at the bench size: x = sin(2 pi i/Nx) sin(2 pi j/Ny) {sin(2 pi k/Nz) | cos(pi k/(Nz-1))}
    -> the spectrum is zero except at kx = 1 (and Nx-1 for C2C), ky = 1 | Ny-1, kz = 1 | Nz-1 (DCT-I: kz = 1), where it is
    +-(N/8) i  (sample/C++/test3D_r2c.C:281-331, 343-369; test1D_cos.C:297-306 for the cosine).  Every rank builds and checks
    its own block from GlobStart.  Returns max |X - expected| / peak over all ranks' blocks (this rank's part)."""
    n, kind = prob.n, prob.kind
    dev = x.device

    def axis(i, cos=False):
        idx = torch.arange(prob.gs1[i], prob.gs1[i] + prob.ld1[i], device=dev, dtype=torch.float64)
        return (torch.cos(math.pi * idx / (n[i] - 1)) if cos else torch.sin(2 * math.pi * idx / n[i])).to(rdt)

    sx, sy, sz = axis(0), axis(1), axis(2, cos=(kind == "cheb"))
    if kind == "c2c":
        xv = torch.view_as_real(x).view(prob.ld1[2], prob.ld1[1], prob.ld1[0], 2)
        xv.zero_()
        torch.mul((sz[:, None] * sy[None, :])[:, :, None], sx[None, None, :], out=xv[..., 0])
    else:
        torch.mul((sz[:, None] * sy[None, :])[:, :, None], sx[None, None, :], out=x.view(prob.ld1[2], prob.ld1[1], prob.ld1[0]))
    prob.forward(x, X, 0)
    torch.cuda.synchronize()
    XL = logical_view(X, prob.ld2, prob.mo2)  # [kx][ky][kz] local
    # expected non-zeros: FFT of sin(2 pi j/N) = (N / 2i) (delta_1 - delta_{N-1}); DCT-I of cos(pi j/(N-1)) = (N-1) delta_1
    def modes(i, half_only=False, cos=False):
        if cos:
            return [(1, float(n[i] - 1))]
        m = [(1, n[i] / 2.0)]
        if not half_only and n[i] > 2:
            m.append((n[i] - 1, -n[i] / 2.0))
        return m
    mx = modes(0, half_only=(kind != "c2c"))
    my = modes(1)
    mz = modes(2, cos=(kind == "cheb"))
    nsin = 2 if kind == "cheb" else 3
    unit = (1 / 1j) ** nsin
    peak = abs(mx[0][1] * my[0][1] * mz[0][1])
    worst = 0.0
    for kx, ax in mx:
        for ky, ay in my:
            for kz, az in mz:
                l = [kx - prob.gs2[0], ky - prob.gs2[1], kz - prob.gs2[2]]
                if all(0 <= l[i] < prob.ld2[i] for i in range(3)):
                    want = complex(unit * ax * ay * az)
                    got = complex(XL[l[0], l[1], l[2]].item())
                    worst = max(worst, abs(got - want) / peak)
                    XL[l[0], l[1], l[2]] = 0  # what remains must be zero
    rest = float(torch.linalg.vector_norm(torch.view_as_real(X).reshape(-1), ord=float("inf")).item()) / peak if X.numel() else 0.0
    return max(worst, rest)


def oracle_check(torch, lib, cfg_kind, single, pdims, rank, rdt, cdt, idir):
    """reduced grid, same processor grid, same planner (overlapped pairs / persistent flag kernels where they are planned):
    forward (with the fused derivative for the Chebyshev configuration) and backward against the CPU oracle on the globally
    indexed Philox field; every rank compares its own block"""
    import numpy as np
    import __graft_entry__ as ge
    orc = ge.load_oracle()
    n = (128, 128, 129) if cfg_kind == "cheb" else (256, 256, 256)
    if DRY:
        n = (16, 16, 17) if cfg_kind == "cheb" else (32, 16, 16)
    world = pdims[0] * pdims[1] * pdims[2]
    if world > 1 and not DRY:  # the production overlap needs the same local extents per rank as the tests' pair-forcing shapes
        n = (n[0], n[1], max(n[2], 32 * world)) if cfg_kind != "cheb" else n
    prob = Problem(lib, cfg_kind, n, single, pdims, idir)
    G = philox_global(n, cfg_kind == "c2c")
    og1 = orc.OGrid(list(n), [0, 1, 2], [0, 1, 2], pdims, rank, -1)
    og2 = orc.OGrid(list(prob.n2), prob.dmap2, prob.mo2, pdims, rank, -1 if cfg_kind == "c2c" else 0)
    assert list(og1.Ldims) == prob.ld1 and list(og2.Ldims) == prob.ld2 and list(og2.GlobStart) == prob.gs2
    npc = np.complex64 if single else np.complex128
    npr = np.float32 if single else np.float64
    a = orc.local_of(G, og1).astype(npc if cfg_kind == "c2c" else npr)
    x = torch.from_numpy(np.ascontiguousarray(a).ravel()).cuda()
    X = torch.empty(max(prob.n2e, 1), device=DEV, dtype=cdt)[:prob.n2e]
    y = torch.empty_like(x)
    res = {"grid": list(n), "proc_grid": list(pdims), "overlap": prob.overlap_summary()}
    tol = 1e-5 if single else 1e-12
    want = orc.local_of(orc.transform_global(G, prob.types_f, list(prob.n2)), og2)
    prob.forward(x, X, 0)
    torch.cuda.synchronize()
    res["fwd_rel_l2"] = orc.rel_l2(X.cpu().numpy(), want.ravel()) if want.size else 0.0
    if idir >= 0:
        wantd = orc.local_of(orc.transform_global(G, prob.types_f, list(prob.n2), deriv_dim=idir), og2)
        prob.forward(x, X, 0, deriv=True)
        torch.cuda.synchronize()
        res["fwd_deriv_rel_l2"] = orc.rel_l2(X.cpu().numpy(), wantd.ravel()) if wantd.size else 0.0
        prob.forward(x, X, 0)
    # backward of the oracle's spectrum, against the input field times the normalisation the unnormalised pair leaves
    Xo = torch.from_numpy(np.ascontiguousarray(want.astype(npc)).ravel()).cuda()
    prob.backward(Xo, y, 1)
    torch.cuda.synchronize()
    res["bwd_rel_l2"] = orc.rel_l2(y.cpu().numpy() / prob.norm, a.ravel()) if a.size else 0.0
    res["tolerance"] = tol
    bad = {k: v for k, v in res.items() if k.endswith("rel_l2") and not v < tol}
    prob.free()
    del x, X, y, Xo
    return res, bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=os.environ.get("P3DFFT_BENCH_CONFIG", "c3"), choices=sorted(CONFIGS))
    ap.add_argument("--edge", "--n", dest="n", type=int, default=0, help="override the cube edge of the configuration")
    ap.add_argument("--grid", default=os.environ.get("P3DFFT_BENCH_GRID", "slab"), choices=["slab", "pencil"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-host-array e2e leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the known-answer and oracle checks (the round trip stays)")
    ap.add_argument("--single", action="store_true", help="single precision variant of the configuration")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = args.config
    conf = dict(CONFIGS[cfg])
    if args.n:
        conf["n"] = (args.n, args.n, args.n + 1 if conf["kind"] == "cheb" else args.n)
        conf["desc"] = conf["desc"] + f" [edge overridden: {args.n}]"
        CONFIGS[cfg] = conf
    if args.single:
        conf["single"] = True
    n, single, kind = conf["n"], conf["single"], conf["kind"]

    if args.impl == "reference":
        run_reference_arm(args, cfg, rank)
        return

    import numpy as np
    import torch
    import __graft_entry__ as ge

    if DRY:
        _dry_run_shims(torch)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else {"node": None}
    if world > 1:
        import torch.distributed as dist
        if DRY:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idle = dist.new_group(backend="gloo")  # ranks that wait for rank 0's CPU leg block on a socket instead of spinning
    mod = ge.load_package()
    lib = mod.load(emulated=DRY).setup()
    assert DRY or lib.have_device()

    rdt, cdt = (torch.float32, torch.complex64) if single else (torch.float64, torch.complex128)
    rb, cb = (4, 8) if single else (8, 16)
    pdims = choose_grid(world, args.grid, cfg)
    idir = 0 if kind == "cheb" else -1
    stream = torch.cuda.current_stream()
    lib.set_stream(stream.cuda_stream)
    tol = 1e-5 if single else 1e-12

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(v, op="max"):
        if world == 1:
            return v
        t = torch.tensor([v], device=DEV, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())

    # ---- parity (1): oracle comparison on a reduced grid with the same processor grid, before the big arrays exist
    parity = {}
    if not args.no_parity:
        res, bad = oracle_check(torch, lib, kind, single, pdims, rank, rdt, cdt, idir)
        for k in [k for k in res if k.endswith("rel_l2")]:
            res[k] = reduce_ranks(res[k])
        parity["oracle"] = res
        nbad = reduce_ranks(float(len(bad)))
        if nbad:
            raise SystemExit(f"bench.py: parity against the oracle FAILED on the reduced grid: {res}")

    prob = Problem(lib, kind, n, single, pdims, idir)
    n1, n2 = prob.n1e, prob.n2e
    in_dt = cdt if kind == "c2c" else rdt
    inb = cb if kind == "c2c" else rb
    # one GPU holding the whole 2048^3 problem: the backward transform runs in place on the spectrum's buffer
    dev_bytes = n1 * inb * 2 + n2 * cb + 2 * prob.dfw["work_bytes"]
    inplace_back = dev_bytes > 0.9 * torch.cuda.get_device_properties(local_rank).total_memory
    x = torch.empty(max(n1, 1), device=DEV, dtype=in_dt)[:n1]
    nX = max(n2, (n1 * inb + cb - 1) // cb if inplace_back else 0, 1)
    Xbuf = torch.empty(nX, device=DEV, dtype=cdt)
    X = Xbuf[:n2]
    y = torch.view_as_real(Xbuf).reshape(-1)[:n1] if inplace_back else torch.empty(max(n1, 1), device=DEV, dtype=in_dt)[:n1]

    # ---- parity (2): the reference sample's known-answer check at the bench size
    if not args.no_parity:
        ka = reduce_ranks(known_answer_check(torch, prob, x, X, rdt, cdt, tol))
        parity["known_answer"] = {"grid": list(n), "max_abs_err_over_peak": ka, "tolerance": tol * 10,
                                  "what": "sine field -> +-N/8 i at (1, +-1, +-1), zero elsewhere (sample/C++/test3D_r2c.C:281-331)"}
        if not ka < tol * 10:
            raise SystemExit(f"bench.py: known-answer check FAILED at the bench size: {parity}")

    # ---- the synthetic field: globally indexed Philox planes, generated on the host into the (pinned) array the e2e leg uses
    workers = max(1, (os.cpu_count() or 1) // max(1, min(world, 8)))
    try:
        hx = torch.empty(max(n1, 1), dtype=in_dt).pin_memory()[:n1]
    except RuntimeError:
        hx = torch.empty(max(n1, 1), dtype=in_dt)[:n1]
    t0 = time.perf_counter()
    if n1:
        fill_philox_block(hx.numpy().reshape(prob.ld1[2], prob.ld1[1], prob.ld1[0]), prob.gs1, prob.ld1, n, kind == "c2c", workers)
    gen_s = time.perf_counter() - t0
    x.copy_(hx, non_blocking=False)

    def step_device():
        prob.forward(x, X, 0, deriv=True)
        prob.backward(X, y, 1)

    # ---- parity (3): the round trip must return norm * input (without the derivative)
    prob.forward(x, X, 0)
    prob.backward(X, y, 1)
    torch.cuda.synchronize()
    d2 = x2 = 0.0
    for c0 in range(0, n1, 1 << 27):  # chunked: no temporaries of the arrays' size
        xc, yc = x[c0:c0 + (1 << 27)], y[c0:c0 + (1 << 27)]
        d2 += float((torch.linalg.vector_norm(yc / prob.norm - xc) ** 2).item())
        x2 += float((torch.linalg.vector_norm(xc) ** 2).item())
    if n1:
        del xc, yc  # (views: they would keep the arrays alive past the `del` before the host-array leg)
    rt_err = math.sqrt(reduce_ranks(d2, "sum") / reduce_ranks(x2, "sum"))
    parity["roundtrip_rel_l2"] = rt_err
    assert rt_err < tol, f"round-trip error {rt_err}"

    # ---- device-resident timing
    small = max(n1 * inb, n2 * cb) < (256 << 20)  # working set that fits the 126 MB L2: flush between steps
    flush = torch.empty(256 << 20, device=DEV, dtype=torch.uint8) if small else None
    sampler = ClockSampler(local_rank)
    if rank == 0 and not DRY:
        sampler.start()
        sampler.ready.wait(10)
    for _ in range(args.warmup):
        step_device()
    lib.enable_timers(True)
    lib.stage_times(prob.pf)
    lib.stage_times(prob.pb)  # (drop what the warm-up recorded)
    barrier()
    sampler.t0 = time.perf_counter()
    l0 = lib.kernel_launches()
    if small:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for e0, e1 in evs:
            flush.zero_()
            e0.record(stream)
            step_device()
            e1.record(stream)
        barrier()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs) / args.steps
        launches = lib.kernel_launches() - l0
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / args.steps
        launches = lib.kernel_launches() - l0
    sampler.t1 = time.perf_counter()
    # per-stage durations averaged over the steps of the timed region (CUDA events recorded by the library on the launch
    # stream around every stage kernel of every exec; read back only now, so nothing synchronises inside the region)
    stage_f = np.array(lib.stage_times(prob.pf))
    stage_b = np.array(lib.stage_times(prob.pb))
    clocks = sampler.stop() if rank == 0 and not DRY else None
    lib.enable_timers(False)
    ms = reduce_ranks(ms)
    gflops = 2 * flops_3d(n) / (ms * 1e-3) / 1e9

    # ---- roofline: per-stage figures, then the dominant stage
    peak, peak_src = measured_peaks()
    stages = []
    for d, t, name in ((prob.dfw, stage_f, "fwd"), (prob.dbw, stage_b, "bwd")):
        base = len(stages)
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
        # overlapped groups (exchange stage + the local stage(s) running beside it as persistent kernels) are timed as a
        # whole; the library books the duration on the group's first stage.  Every member is labelled with the group, the
        # exchange member carries the NVLink figure over the whole group's duration, local members carry no figure of their own
        i = 0
        S = len(d["stages"])
        while i < S:
            s = d["stages"][i]
            if s.get("pair"):
                tri = i + 2 < S and d["stages"][i + 1].get("triple") and s.get("pair_sync") and i + 3 == S and s["pair"] == 1
                members = stages[base + i: base + i + (3 if tri else 2)]
                gms = sum(m["ms"] for m in members)
                gname = "+".join(m["stage"] for m in members)
                for m in members:
                    m["ms"] = gms
                    m["overlap_group"] = gname
                    m["gbs"] = None
                    if m["exchange"]:
                        m["nvlink_gbs"] = m["nvlink_bytes"] / (gms * 1e-3) / 1e9 if gms > 0 else None
                members[0]["group_ms"] = gms
                i += len(members)
            else:
                stages[base + i]["group_ms"] = stages[base + i]["ms"]
                i += 1
    local = [s for s in stages if not s["exchange"] and s["gbs"]]
    hbm = None
    if local:
        dom = max(local, key=lambda s: s["ms"])
        hbm = {"bound": "hbm", "achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": dom["gbs"] / peak, "traffic": None,
               "kernel": f"{dom['variant']} ({dom['stage']}, dim {dom['dim']})", "peak_source": peak_src,
               "alg_bytes_per_launch": dom["alg_bytes"], "ms_per_launch": dom["ms"]}
    total_alg = sum(s["alg_bytes"] for s in stages)
    xs = [s for s in stages if s["exchange"]]
    if xs:
        # N > 1: the step is bound by the fused exchange stages (each overlapped with its neighbouring local stages):
        # NVLink roofline, 900 GB/s per direction per GPU nominal (rank 0's figures; SM-issued all-to-all stores measure
        # 669 GB/s on 8 GPUs: tools/microbench/a2a_bench.cu, profiles/r02_a2a_store_microbench.txt)
        xd = max(xs, key=lambda s: s["ms"])
        tot_b = sum(s["nvlink_bytes"] for s in xs)
        tot_ms = sum(s["ms"] for s in xs)
        roofline = {"bound": "nvlink", "achieved": xd["nvlink_gbs"], "peak": 900.0, "unit": "GB/s",
                    "frac": xd["nvlink_gbs"] / 900.0 if xd["nvlink_gbs"] else None, "traffic": None,
                    "kernel": f"{xd['variant']} ({xd.get('overlap_group', xd['stage'])}: fused FFT + all-to-all over NVLink, "
                              f"overlapped with the neighbouring local stages; duration of the whole group)",
                    "peak_source": "NVLink 5 nominal, per direction per GPU", "peak_measured_sm_store_all_to_all": 669.0,
                    "bytes_per_launch": xd["nvlink_bytes"], "ms_per_launch": xd["ms"],
                    "all_exchanges": {"bytes": tot_b, "ms": tot_ms, "achieved": tot_b / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else None,
                                      "frac": tot_b / (tot_ms * 1e-3) / 1e9 / 900.0 if tot_ms > 0 else None},
                    "whole_step_frac_of_exchange_bound": (tot_b / 900e9) / (ms * 1e-3),
                    "hbm_slowest_unoverlapped_local_stage": hbm}
    else:
        roofline = hbm or {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None}
    roofline["traffic_note"] = ("DRAM bytes need a profiler pass: see profiles/ (ncu --set full of the same command); "
                                "not copied into a timed run")
    roofline["stages"] = stages
    roofline["whole_step_hbm_frac"] = total_alg / (ms * 1e-3) / 1e9 / peak

    # ---- end to end: host arrays through the same C ABI calls (H2D of the input and D2H of the result inside each call)
    e2e = None
    if not args.no_e2e:
        del y
        y = None
        hb = n1 * inb + n2 * cb
        hb_all = int(reduce_ranks(float(hb), "sum"))

        def run_host(ha, hA, steps):
            def step_host():
                prob.forward(ha, hA, 0, deriv=True)   # H2D field, 3 stages, D2H spectrum
                prob.backward(hA, ha, 1)              # H2D spectrum, 3 stages, D2H field
            step_host()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                step_host()
            torch.cuda.synchronize()
            return reduce_ranks((time.perf_counter() - t0) / steps)

        # the device-resident arrays are not needed any more: give their memory to the library's staging buffers
        del x, X, Xbuf
        torch.cuda.empty_cache()
        if not DRY:
            free_b, total_b = torch.cuda.mem_get_info()
            sys.stderr.write(f"[bench rank {rank}] before the host-array leg: {free_b / 1e9:.1f} GB of {total_b / 1e9:.1f} GB device memory free, "
                             f"torch holds {torch.cuda.memory_reserved() / 1e9:.1f} GB; staging needs {hb / 1e9:.1f} GB\n")
        try:
            if not DRY and hb > free_b:
                raise RuntimeError(f"device staging buffers of {hb / 1e9:.1f} GB do not fit next to the work buffers ({free_b / 1e9:.1f} GB free)")
            hX = torch.empty(max(n2, 1), dtype=cdt).pin_memory()[:n2]
            dt = run_host(hx, hX, args.e2e_steps)
            e2e = {"value": 2 * flops_3d(n) / dt / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": hb_all, "d2h_bytes_per_step": hb_all,
                   "ms_per_step": dt * 1e3, "steps": args.e2e_steps, "host_memory": "pinned (cudaHostAlloc by the caller)",
                   "numa": numa}
            del hX
        except RuntimeError as e:
            e2e = {"value": None, "unit": "GFLOP/s", "why": f"host-array leg skipped: {e}"[:300]}
        if not args.no_pageable and e2e.get("value"):
            # what a user of the reference passes: plain heap arrays.  "ring" (the library's default): pinned staging ring with a
            # multi-threaded CPU copy; "register": the library page-locks the arrays on first use (the untimed first call pays)
            import ctypes
            pa = np.empty(max(n1, 1), dtype=hx.numpy().dtype)[:n1]
            pa[...] = hx.numpy()
            del hx
            pA = np.empty(max(n2, 1), dtype=np.complex64 if single else np.complex128)[:n2]
            for mode in ("ring", "register"):
                lib.dll.p3dfft_b200_set_host_staging(mode.encode())
                t0 = time.perf_counter()
                prob.forward(pa, pA, 0, deriv=True)
                first = time.perf_counter() - t0
                dt = run_host(pa, pA, args.e2e_steps)
                e2e["pageable" if mode == "ring" else "pageable_registered"] = {
                    "value": 2 * flops_3d(n) / dt / 1e9, "unit": "GFLOP/s", "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
                    "host_memory": "pageable (numpy heap arrays)", "staging": mode + (" (library default)" if mode == "ring" else
                                                                                      " (opt-in: P3DFFT_B200_HOST_STAGING=register)"),
                    "first_forward_call_ms": first * 1e3, "ratio_to_pinned": e2e["ms_per_step"] / (dt * 1e3)}
            lib.dll.p3dfft_b200_host_release(ctypes.c_void_p(pa.ctypes.data))
            lib.dll.p3dfft_b200_host_release(ctypes.c_void_p(pA.ctypes.data))
            lib.dll.p3dfft_b200_set_host_staging(b"ring")
            del pa, pA

    cpu = None
    if not args.no_cpu:
        if rank == 0 and not DRY:
            cpu = cpu_baseline_leg(cfg)
        if world > 1:
            dist.barrier(group=idle)

    if rank == 0:
        line = {"metric": "3D R2C+C2R GFLOP/s (5N log2N)", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32" if single else "f64", "data": "synthetic",
                "config": {"workload": conf["desc"], "config": cfg, "proc_grid": pdims,
                           "field": f"globally indexed Philox planes (key {KEY}, one stream per z plane), generated in {gen_s:.1f} s",
                           "l2": ("working set fits the L2: 256 MB written between timed steps" if small else
                                  "inputs larger than L2 (per-GPU arrays >= 1 GB vs 126 MB L2)"),
                           "backward_in_place": bool(inplace_back), "overlap": prob.overlap_summary()},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "parity": parity}
        if DRY:
            line = {"dry_run": True, "emulated_library": True, "would_print": line}
        print(json.dumps(line), flush=True)
    prob.free()
    lib.cleanup()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
