#!/bin/bash
# tuning sweep on one GPU: bench (device-resident leg only) under different kernel-shape overrides; one summary line per setting.
# SWEEP entries: "<pencils>" (0 = default) optionally suffixed with ":nopipe"
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
if [ -n "$PRETEST" ]; then timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$PRETEST" 2>&1 | tail -5 | tee -a gpurun_out/sweep.txt; fi
for S in ${SWEEP:-0 0:nopipe 4 8 16}; do
  P=${S%%:*}
  unset P3DFFT_B200_POW2_PENCILS P3DFFT_B200_NO_PIPE
  [ "$P" != 0 ] && export P3DFFT_B200_POW2_PENCILS=$P
  [[ "$S" == *nopipe* ]] && export P3DFFT_B200_NO_PIPE=1
  timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu --no-e2e ${BENCH_ARGS:-} > gpurun_out/sweep_$S.json 2> gpurun_out/sweep_$S.err
  python - "$S" <<'PY' | tee -a gpurun_out/sweep.txt
import json, sys
p = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/sweep_{p}.json").read().strip().splitlines()[-1])
    st = " ".join(f"{s['stage']}:{s['gbs']:.0f}" for s in d["roofline"]["stages"])
    print(f"cfg={p} ms={d['ms_per_step']:.2f} gflops={d['value']:.0f} | {st} | {d['roofline']['stages'][1]['variant']}")
except Exception as e:
    print(f"cfg={p} FAILED {e}", open(f"gpurun_out/sweep_{p}.err").read()[-800:])
PY
done
