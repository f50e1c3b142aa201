#!/bin/bash
# multi-GPU session on N = $1 GPUs: (optionally) parity on N ranks, then the bench line at N (slab, and pencil when N >= 4 and PENCIL=1)
N=${1:-2}
TAG=${TAG:-r01}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== pytest multirank gpu"; timeout 1500 python -m pytest tests/test_multirank.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_multi_$N.log
fi
for grid in slab pencil; do
  if [ "$grid" = pencil ] && { [ "$N" -lt 4 ] || [ "${PENCIL:-0}" != "1" ]; }; then continue; fi
  echo "== bench N=$N $grid"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --grid $grid ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench_${N}gpu_$grid.json 2> gpurun_out/${TAG}_bench_${N}gpu_$grid.err
  python - gpurun_out/${TAG}_bench_${N}gpu_$grid.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    st = " ".join(f"{s['stage']}:{s['ms']:.2f}" + (f"({s['nvlink_gbs']:.0f})" if s.get('nvlink_gbs') else "") + ("*" if s.get("overlapped_with") else "") for s in d["roofline"]["stages"])
    print(f"n_gpus={d['n_gpus']} grid={d['config']['proc_grid']} ms={d['ms_per_step']:.2f} gflops={d['value']:.0f} e2e={d['e2e'] and round(d['e2e']['value'])} clocks={d['clocks']} | {st}")
    print("   nvlink:", {k: v for k, v in d["roofline"].get("nvlink", {}).items() if k in ("achieved", "frac", "whole_step_frac_of_exchange_bound")})
except Exception as e:
    print("FAILED", e, open(sys.argv[1].replace(".json", ".err")).read()[-800:])
PY
done
