#!/bin/bash
# N-GPU sweep of the CTA cap of the fused exchange stages (P3DFFT_B200_XGRID): bench line summary per setting
N=${1:-2}
mkdir -p gpurun_out
: > gpurun_out/xgrid_$N.txt
for G in ${XGRIDS:-off 74 60 90}; do
  # off: no overlap; otherwise the SMs given to the exchange stage of an overlapped pair
  if [ "$G" = off ]; then export P3DFFT_B200_OVERLAP=0; else export P3DFFT_B200_OVERLAP=1 P3DFFT_B200_OVERLAP_XSMS=$G; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
     bench.py --gpus $N --no-cpu --no-e2e --steps 5 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/xgrid_${N}_$G.json 2> gpurun_out/xgrid_${N}_$G.err
  python - $N $G <<'PY' | tee -a gpurun_out/xgrid_$N.txt
import json, sys
n, g = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/xgrid_{n}_{g}.json").read().strip().splitlines()[-1])
    st = " ".join(f"{s['stage']}:{s['ms']:.2f}" + (f"({s['nvlink_gbs']:.0f})" if s.get('nvlink_gbs') else "") + ("*" if s.get("overlapped_with") else "") for s in d["roofline"]["stages"])
    print(f"xgrid={g} ms={d['ms_per_step']:.2f} gflops={d['value']:.0f} | {st}")
except Exception as e:
    print(f"xgrid={g} FAILED {e}", open(f"gpurun_out/xgrid_{n}_{g}.err").read()[-600:])
PY
done
