#!/bin/bash
# tuning sweep on one GPU: bench (device-resident leg only) under different kernel-shape overrides; one summary line per setting
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
for P in ${SWEEP:-0 2 4 8 16}; do
  if [ "$P" = 0 ]; then unset P3DFFT_B200_POW2_PENCILS; else export P3DFFT_B200_POW2_PENCILS=$P; fi
  timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu --no-e2e ${BENCH_ARGS:-} > gpurun_out/sweep_$P.json 2> gpurun_out/sweep_$P.err
  python - "$P" <<'PY' | tee -a gpurun_out/sweep.txt
import json, sys
p = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/sweep_{p}.json").read().strip().splitlines()[-1])
    st = " ".join(f"{s['stage']}:{s['gbs']:.0f}({s['variant']})" for s in d["roofline"]["stages"])
    print(f"pencils={p} ms={d['ms_per_step']:.2f} gflops={d['value']:.0f} | {st}")
except Exception as e:
    print(f"pencils={p} FAILED {e}", open(f"gpurun_out/sweep_{p}.err").read()[-500:])
PY
done
