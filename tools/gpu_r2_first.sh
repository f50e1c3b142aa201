#!/bin/bash
# round-2 first 2-GPU session: store-path microbench, then multi-rank parity with ranks sharing the two GPUs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpus.txt
timeout 300 tools/microbench/_build/xchg_bench > gpurun_out/r02_xchg_bench.txt 2>&1
tail -70 gpurun_out/r02_xchg_bench.txt
for n in 2 4 8; do
  ( time timeout 900 python -m pytest "tests/test_multirank.py::test_gpu_multirank_parity[$n]" -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r02_pytest_multi_shared_$n.log 2>&1
  cat gpurun_out/r02_pytest_multi_shared_$n.log
done
