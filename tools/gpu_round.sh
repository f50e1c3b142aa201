#!/bin/bash
# one 1-GPU session: smoke, full bench line, ncu launch list of the same command, ncu --set full of the stage kernels
# (csv pages exported on the box), optionally the gpu tests.  Everything lands in gpurun_out/ with the prefix $TAG
TAG=${TAG:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
echo "== bench"; timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err; tail -c 1500 gpurun_out/${TAG}_bench_1gpu.json; tail -3 gpurun_out/${TAG}_bench_1gpu.err
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>&1; tail -c 600 gpurun_out/${TAG}_bench_reference.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_ncu_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch_bench.log 2>&1
tail -2 gpurun_out/ncu_launch_bench.log | cut -c1-200
echo "== ncu full"
KREGEX="pipe_kernel|stage_kernel" SKIP=6 COUNT=6 OUT=${TAG}_ncu_full bash tools/gpu_ncu.sh
fi
if [ "${SKIP_TESTS:-1}" != "1" ]; then
echo "== pytest gpu"; timeout ${PYTEST_TIMEOUT:-1200} python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_gpu.log
fi
