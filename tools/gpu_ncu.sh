#!/bin/bash
# ncu --set full capture of stage kernels: KREGEX (kernel name regex), SKIP, COUNT; report -> gpurun_out/$OUT.ncu-rep
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-stage_kernel} -s ${SKIP:-6} -c ${COUNT:-3} -f -o gpurun_out/${OUT:-prof} \
   python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e ${BENCH_ARGS:-} > gpurun_out/ncu_${OUT:-prof}.log 2>&1
tail -3 gpurun_out/ncu_${OUT:-prof}.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
