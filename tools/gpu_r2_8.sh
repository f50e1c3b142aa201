#!/bin/bash
# round-2 8-GPU session: all-to-all store ceiling, parity on 8 real GPUs, step time of the persistent pair kernels by SM split
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpus8.txt
timeout 120 tools/microbench/_build/a2a_bench 8 > gpurun_out/r02_a2a_bench_8.txt 2>&1
timeout 120 tools/microbench/_build/a2a_bench 4 >> gpurun_out/r02_a2a_bench_8.txt 2>&1
cat gpurun_out/r02_a2a_bench_8.txt
( time timeout 900 python -m pytest tests/test_multirank.py -m gpu -x -q -k "parity and 8" 2>&1 | tail -5 ) > gpurun_out/r02_pytest_multi_8gpu.log 2>&1
cat gpurun_out/r02_pytest_multi_8gpu.log
SKIP_TESTS=1 SYNCS="1" XSMS="${XSMS:-56 74 92}" bash tools/gpu_r2_pairs.sh 8
