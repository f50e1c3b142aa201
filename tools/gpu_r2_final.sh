#!/bin/bash
# round 2, final 1-GPU session on the committed code: smoke, full gpu test suite, default bench line, reference arm
TAG=${TAG:-r02f}
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
echo "== pytest gpu"; (time timeout 1800 python -m pytest tests -m gpu -q) 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/${TAG}_bench_c3_1gpu.json 2> gpurun_out/${TAG}_bench_c3_1gpu.err; tail -c 300 gpurun_out/${TAG}_bench_c3_1gpu.json; tail -3 gpurun_out/${TAG}_bench_c3_1gpu.err
echo "== other shapes (768, 640, C4)"; for k in 768 640 C4; do timeout 300 python tools/gpu_configs.py $k; done 2>&1 | tee gpurun_out/${TAG}_other_configs_1gpu.txt | cut -c1-420
echo "== bench --impl reference"; (time timeout 900 python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/${TAG}_bench_reference.json 2>&1; tail -c 700 gpurun_out/${TAG}_bench_reference.json
