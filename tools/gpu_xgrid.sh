#!/bin/bash
# N-GPU sweep of the overlapped-pair settings: entries "off" (no overlap) or "<SMs for the exchange stage>:<chunks>[:trace]"
N=${1:-2}
mkdir -p gpurun_out
: > gpurun_out/xgrid_$N.txt
for G in ${XGRIDS:-off 74:4 74:8}; do
  unset P3DFFT_B200_OVERLAP P3DFFT_B200_OVERLAP_XSMS P3DFFT_B200_OVERLAP_CHUNKS P3DFFT_B200_OVERLAP_TRACE
  unset P3DFFT_B200_XGRID
  if [ "$G" = off ]; then export P3DFFT_B200_OVERLAP=0;
  elif [[ "$G" == offx* ]]; then export P3DFFT_B200_OVERLAP=0 P3DFFT_B200_XGRID=${G#offx};   # no overlap, exchange stages capped to N CTAs
  else
    IFS=: read -r XS CH TR <<< "$G"
    export P3DFFT_B200_OVERLAP=1 P3DFFT_B200_OVERLAP_XSMS=$XS P3DFFT_B200_OVERLAP_CHUNKS=$CH
    [ -n "$TR" ] && export P3DFFT_B200_OVERLAP_TRACE=1
  fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
     bench.py --gpus $N --no-cpu --no-e2e --steps ${STEPS:-10} --warmup 3 --grid ${GRID:-slab} ${BENCH_ARGS:-} > gpurun_out/xgrid_${N}_$G.json 2> gpurun_out/xgrid_${N}_$G.err
  grep "pair trace" gpurun_out/xgrid_${N}_$G.err | tail -2 | cut -c1-700 | tee -a gpurun_out/xgrid_$N.txt
  python - $N $G <<'PY' | tee -a gpurun_out/xgrid_$N.txt
import json, sys
n, g = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/xgrid_{n}_{g}.json").read().strip().splitlines()[-1])
    st = " ".join(f"{s['stage']}:{s['ms']:.2f}" + (f"({s['nvlink_gbs']:.0f})" if s.get('nvlink_gbs') else "") + ("*" if s.get("overlapped_with") else "") for s in d["roofline"]["stages"])
    print(f"cfg={g} ms={d['ms_per_step']:.2f} gflops={d['value']:.0f} | {st}")
except Exception as e:
    print(f"cfg={g} FAILED {e}", open(f"gpurun_out/xgrid_{n}_{g}.err").read()[-600:])
PY
done
