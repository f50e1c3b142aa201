#!/bin/bash
# round 2, third 1-GPU session: full gpu test suite, shapes table, C5 on one GPU with the host-array leg
TAG=${TAG:-r02c}
mkdir -p gpurun_out
echo "== pytest gpu"; (time timeout ${PYTEST_TIMEOUT:-1800} python -m pytest tests -m gpu -q ${PYTEST_ARGS:-}) 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== other shapes"; timeout 600 python tools/gpu_configs.py > gpurun_out/${TAG}_other_configs_1gpu.txt 2>&1; cat gpurun_out/${TAG}_other_configs_1gpu.txt | cut -c1-420
echo "== bench c5 (2048^3 single on one GPU, 137 GB)"; (time timeout 900 python bench.py --config c5 --steps 5 --e2e-steps 2 --no-pageable --no-cpu) > gpurun_out/${TAG}_bench_c5_1gpu.json 2> gpurun_out/${TAG}_bench_c5_1gpu.err; tail -c 400 gpurun_out/${TAG}_bench_c5_1gpu.json; tail -5 gpurun_out/${TAG}_bench_c5_1gpu.err
echo "== bench c3"; timeout 900 python bench.py > gpurun_out/${TAG}_bench_c3_1gpu.json 2> gpurun_out/${TAG}_bench_c3_1gpu.err; tail -c 300 gpurun_out/${TAG}_bench_c3_1gpu.json; tail -3 gpurun_out/${TAG}_bench_c3_1gpu.err
echo "== ncu full 768^3"
NCU_CMD="python tools/gpu_configs.py 768" KREGEX="pipe_kernel|stage_kernel" SKIP=6 COUNT=6 OUT=${TAG}_ncu_full_768 bash tools/gpu_ncu.sh
rm -f gpurun_out/*.source.csv.gz gpurun_out/*.raw.csv
