#!/bin/bash
# one GPU session: smoke, bench, ncu launch list + full capture of the stage kernels, gpu tests.  Logs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch_bench.log 2>&1
tail -3 gpurun_out/ncu_launch_bench.log
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 12 -c 6 -f -o gpurun_out/prof \
   python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full_bench.log 2>&1
tail -3 gpurun_out/ncu_full_bench.log
fi
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== pytest gpu"; timeout ${PYTEST_TIMEOUT:-1200} python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
fi
ls -la gpurun_out
