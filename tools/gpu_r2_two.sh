#!/bin/bash
# round 2, second 1-GPU session: gpu tests, C4 / other shapes with the compile-time DCT-I and the mixed-radix kernels,
# staging-ring thread sweep, C5 on one GPU, compute-sanitizer on a subset of the gpu tests, ncu of C4 / 768^3 kernels
TAG=${TAG:-r02b}
mkdir -p gpurun_out
echo "== pytest gpu"; (time timeout ${PYTEST_TIMEOUT:-1500} python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-}) 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== other shapes"; timeout 600 python tools/gpu_configs.py > gpurun_out/${TAG}_other_configs_1gpu.txt 2>&1; cat gpurun_out/${TAG}_other_configs_1gpu.txt | cut -c1-420
echo "== bench c4"; timeout 600 python bench.py --config c4 > gpurun_out/${TAG}_bench_c4_1gpu.json 2> gpurun_out/${TAG}_bench_c4_1gpu.err; tail -c 300 gpurun_out/${TAG}_bench_c4_1gpu.json; tail -3 gpurun_out/${TAG}_bench_c4_1gpu.err
echo "== bench c2"; timeout 600 python bench.py --config c2 --no-cpu > gpurun_out/${TAG}_bench_c2_1gpu.json 2> gpurun_out/${TAG}_bench_c2_1gpu.err; tail -c 300 gpurun_out/${TAG}_bench_c2_1gpu.json; tail -3 gpurun_out/${TAG}_bench_c2_1gpu.err
echo "== staging ring threads"
for t in 2 8 16; do
P3DFFT_B200_HOST_THREADS=$t timeout 600 python bench.py --no-cpu --no-parity --steps 3 > gpurun_out/${TAG}_bench_c3_ring$t.json 2> gpurun_out/${TAG}_bench_c3_ring$t.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_c3_ring$t.json')); e=d['e2e']; print('threads $t: pinned', round(e['ms_per_step'],1), 'ring', round(e['pageable']['ms_per_step'],1), 'registered', round(e['pageable_registered']['ms_per_step'],1), 'device', round(d['ms_per_step'],2))"
done
echo "== bench c5 (2048^3 single on one GPU, 137 GB)"; (time timeout 900 python bench.py --config c5 --steps 5 --e2e-steps 2 --no-pageable --no-cpu) > gpurun_out/${TAG}_bench_c5_1gpu.json 2> gpurun_out/${TAG}_bench_c5_1gpu.err; tail -c 400 gpurun_out/${TAG}_bench_c5_1gpu.json; tail -5 gpurun_out/${TAG}_bench_c5_1gpu.err
echo "== compute-sanitizer racecheck / memcheck (subset of the gpu tests)"
SAN="tests/test_gpu_parity.py -k (test_config_c1_single_rank_128 or test_fused_derivative or (test_pow2_c2c_sizes and 1024 and pipe) or (test_r2r_kinds_on_pipe_kernel and 256) or (test_smooth_lengths and 768))"
(time timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest -m gpu -x -q tests/test_gpu_parity.py -k "test_config_c1_single_rank_128 or test_fused_derivative or (test_pow2_c2c_sizes and 1024 and pipe) or (test_r2r_kinds_on_pipe_kernel and 256) or (test_smooth_lengths and 768)") > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1; tail -8 gpurun_out/${TAG}_sanitizer_racecheck.log
(time timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest -m gpu -x -q tests/test_gpu_parity.py -k "test_config_c1_single_rank_128 or test_fused_derivative or (test_pow2_c2c_sizes and 1024 and pipe) or (test_r2r_kinds_on_pipe_kernel and 256) or (test_smooth_lengths and 768) or test_r2c_c2r_any_length") > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1; tail -8 gpurun_out/${TAG}_sanitizer_memcheck.log
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu full c4 (DCT stage)"
BENCH_ARGS="--no-parity --config c4" KREGEX="pipe_kernel|stage_kernel" SKIP=6 COUNT=6 OUT=${TAG}_ncu_full_c4 bash tools/gpu_ncu.sh
echo "== ncu full 768^3"
NCU_CMD="python tools/gpu_configs.py 768" KREGEX="pipe_kernel|stage_kernel" SKIP=6 COUNT=6 OUT=${TAG}_ncu_full_768 bash tools/gpu_ncu.sh
rm -f gpurun_out/*.source.csv.gz gpurun_out/*.raw.csv
fi
