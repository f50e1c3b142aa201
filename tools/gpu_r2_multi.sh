#!/bin/bash
# round 2, multi-GPU session on N GPUs of one box (N from the environment, default: all visible): bench lines (slab with the
# full parity / e2e / cpu legs; triple off; pencil; C5 / C4 where the grid fits), then the N-rank gpu tests
N=${N:-$(nvidia-smi -L | wc -l)}
TAG=${TAG:-r02m}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpus_$N.txt 2>&1; nproc >> gpurun_out/${TAG}_gpus_$N.txt
run() {  # name, env assignments..., -- bench args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
     bench.py --gpus $N "$@" > gpurun_out/${TAG}_bench_${name}_${N}gpu.json 2> gpurun_out/${TAG}_bench_${name}_${N}gpu.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_bench_${name}_${N}gpu.json") if l.startswith("{")][-1])
    r = d["roofline"]
    st = " ".join(f"{s['stage']}:{s['ms']:.2f}" + ("*" if s.get("overlap_group") else "") for s in r["stages"])
    e = d.get("e2e") or {}
    print(f"N=$N ${name}: {d['ms_per_step']:.3f} ms {d['value']:.0f} GF | {st} | frac_of_exchange_bound {r.get('whole_step_frac_of_exchange_bound')} | e2e {e.get('ms_per_step')} pageable {(e.get('pageable') or {}).get('ms_per_step')} | parity {json.dumps(d.get('parity'))[:300]}")
except Exception as ex:
    print("N=$N ${name}: FAILED", ex)
PY
  tail -2 gpurun_out/${TAG}_bench_${name}_${N}gpu.err | cut -c1-300
}
run c3_slab P3DFFT_B200_OVERLAP_TRACE=0 -- 
if [ "${QUICK:-0}" != "1" ]; then
run c3_slab_notriple P3DFFT_B200_TRIPLE=0 -- --no-e2e --no-cpu --no-parity
run c3_slab_trace P3DFFT_B200_OVERLAP_TRACE=1 -- --no-e2e --no-cpu --no-parity --steps 3
if [ $N -ge 4 ]; then run c3_pencil P3DFFT_B200_OVERLAP_TRACE=0 -- --grid pencil --no-e2e --no-cpu; fi
fi
if [ $N -eq 4 ]; then run c4_2x2 P3DFFT_B200_OVERLAP_TRACE=0 -- --config c4 --no-cpu; run c1_2x2 P3DFFT_B200_OVERLAP_TRACE=0 -- --config c1 --no-cpu; fi
if [ $N -eq 8 ]; then run c5_slab P3DFFT_B200_OVERLAP_TRACE=0 -- --config c5 --no-cpu --e2e-steps 3; fi
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== pytest multirank on $N GPUs"; (time timeout 1500 python -m pytest tests/test_multirank.py -m gpu -x -q -k "[$N]") 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_multi_${N}gpu.log
fi
