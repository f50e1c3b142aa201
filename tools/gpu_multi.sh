#!/bin/bash
# multi-GPU session: N = $1.  parity on N ranks, then bench at N (slab and pencil)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
echo "== pytest multirank gpu"; timeout 1500 python -m pytest tests/test_multirank.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_multi_$N.log
for grid in slab pencil; do
  if [ "$grid" = pencil ] && [ "$N" -lt 4 ]; then continue; fi
  echo "== bench N=$N $grid"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --grid $grid ${BENCH_ARGS:-} > gpurun_out/bench_${N}_$grid.json 2> gpurun_out/bench_${N}_$grid.err
  tail -c 2500 gpurun_out/bench_${N}_$grid.json; tail -5 gpurun_out/bench_${N}_$grid.err
done
