#!/bin/bash
# round 2, one 1-GPU session: smoke, gpu tests, bench lines of every BASELINE config, the reference arm, A/B switches,
# ncu launch list + full capture.  Every step is independent; everything lands in gpurun_out/ with the prefix $TAG
TAG=${TAG:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt; free -g | head -2 >> gpurun_out/${TAG}_gpu.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
echo "== bench c3"; timeout 900 python bench.py > gpurun_out/${TAG}_bench_c3_1gpu.json 2> gpurun_out/${TAG}_bench_c3_1gpu.err; tail -c 600 gpurun_out/${TAG}_bench_c3_1gpu.json; tail -3 gpurun_out/${TAG}_bench_c3_1gpu.err
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== pytest gpu"; (time timeout ${PYTEST_TIMEOUT:-1500} python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-}) 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest_gpu.log
fi
for c in c4 c2 c1; do
echo "== bench $c"; timeout 600 python bench.py --config $c > gpurun_out/${TAG}_bench_${c}_1gpu.json 2> gpurun_out/${TAG}_bench_${c}_1gpu.err; tail -c 300 gpurun_out/${TAG}_bench_${c}_1gpu.json; tail -3 gpurun_out/${TAG}_bench_${c}_1gpu.err
done
echo "== other shapes"; timeout 600 python tools/gpu_configs.py > gpurun_out/${TAG}_other_configs_1gpu.txt 2>&1; cat gpurun_out/${TAG}_other_configs_1gpu.txt | cut -c1-400
echo "== A/B: 4 pencils per transposing tile (2 CTAs per SM)"
P3DFFT_B200_POW2_PENCILS=4 timeout 600 python tools/gpu_configs.py > gpurun_out/${TAG}_other_configs_p4.txt 2>&1; grep -E "1024|C2 |C5" gpurun_out/${TAG}_other_configs_p4.txt | cut -c1-400
P3DFFT_B200_POW2_PENCILS=4 timeout 600 python bench.py --no-cpu --no-e2e --no-parity > gpurun_out/${TAG}_bench_c3_p4.json 2>&1; tail -c 300 gpurun_out/${TAG}_bench_c3_p4.json
echo "== A/B: 8 pencils per transposing tile in single precision"
P3DFFT_B200_POW2_PENCILS=8 timeout 600 python tools/gpu_configs.py > gpurun_out/${TAG}_other_configs_p8.txt 2>&1; grep -E "single" gpurun_out/${TAG}_other_configs_p8.txt | cut -c1-400
echo "== A/B: DCT stage on fastcore"
P3DFFT_B200_NO_PIPE_R2R=1 timeout 600 python tools/gpu_configs.py 2>&1 | grep "C4" | tee gpurun_out/${TAG}_c4_fastcore.txt | cut -c1-400
echo "== bench --impl reference"; (time timeout 900 python bench.py --impl reference --steps 1 --warmup 0) > gpurun_out/${TAG}_bench_reference.json 2>&1; tail -c 900 gpurun_out/${TAG}_bench_reference.json
if [ "${SKIP_C5:-0}" != "1" ]; then
echo "== bench c5 (2048^3 single on one GPU, 137 GB)"; (time timeout 900 python bench.py --config c5 --steps 5 --e2e-steps 2 --no-pageable --no-cpu) > gpurun_out/${TAG}_bench_c5_1gpu.json 2> gpurun_out/${TAG}_bench_c5_1gpu.err; tail -c 400 gpurun_out/${TAG}_bench_c5_1gpu.json; tail -5 gpurun_out/${TAG}_bench_c5_1gpu.err
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_ncu_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/ncu_launch_bench.log 2>&1
tail -2 gpurun_out/ncu_launch_bench.log | cut -c1-200
echo "== ncu full c3"
BENCH_ARGS="--no-parity" KREGEX="pipe_kernel|stage_kernel" SKIP=6 COUNT=6 OUT=${TAG}_ncu_full bash tools/gpu_ncu.sh
echo "== ncu full c4 (DCT stage)"
BENCH_ARGS="--no-parity --config c4" KREGEX="pipe_kernel|stage_kernel" SKIP=6 COUNT=6 OUT=${TAG}_ncu_full_c4 bash tools/gpu_ncu.sh
rm -f gpurun_out/*.source.csv.gz gpurun_out/*.raw.csv
fi
