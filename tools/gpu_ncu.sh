#!/bin/bash
# ncu --set full capture of stage kernels: KREGEX (kernel name regex), SKIP, COUNT.  The report itself is too large to
# travel back (64 MiB cap), so the csv pages are exported on the box: gpurun_out/$OUT.{raw,details,source}.csv(.gz)
mkdir -p gpurun_out
OUT=${OUT:-prof}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-stage_kernel} -s ${SKIP:-6} -c ${COUNT:-3} -f -o /tmp/$OUT \
   ${NCU_CMD:-python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e ${BENCH_ARGS:-}} > gpurun_out/ncu_$OUT.log 2>&1
tail -3 gpurun_out/ncu_$OUT.log | cut -c1-300
ncu -i /tmp/$OUT.ncu-rep --page raw --csv > gpurun_out/$OUT.raw.csv 2>/dev/null
ncu -i /tmp/$OUT.ncu-rep --page details --csv > gpurun_out/$OUT.details.csv 2>/dev/null
ncu -i /tmp/$OUT.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/$OUT.source.csv.gz
python tools/ncu_summary.py /tmp/$OUT.ncu-rep gpurun_out/$OUT.summary.csv
ls -la /tmp/$OUT.ncu-rep gpurun_out/$OUT.*
