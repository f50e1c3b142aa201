#!/bin/bash
# round 2, last 1-GPU session: the tensor-load kernel (pow2_tload.cuh) on hardware -- parity first, then A/B timing against the
# plain-load kernel, one ncu capture, then as much of the rest of the gpu suite as the remaining box time allows
TAG=${TAG:-r02t}
mkdir -p gpurun_out
T0=$(date +%s)
left() { echo $(( ${BUDGET_S:-470} - ($(date +%s) - T0) )); }
echo "== tload parity"; (time timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_load") 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_tload.log
echo "== tload timing (left $(left) s)"; timeout 150 python tools/gpu_configs.py TLOAD 2>&1 | tee gpurun_out/${TAG}_tload_timing.txt | cut -c1-400
echo "== load-side transposition plan at the bench size, checked (left $(left) s)"
P3DFFT_B200_COST_TLOAD=1.0 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "known_answer_at_bench_size_1024 or roundtrip_1024 or config_c2_512 or config_c1_single" 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_loadside.log
echo "== ncu (left $(left) s)"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:pow2_tload -c 3 -f -o /tmp/${TAG}_tload \
  python tools/gpu_configs.py "TLOAD 1024^3 R2C double mo 102->120, tensor" > gpurun_out/${TAG}_ncu_tload.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_tload.log | cut -c1-300
ncu -i /tmp/${TAG}_tload.ncu-rep --page details --csv > gpurun_out/${TAG}_ncu_tload.details.csv 2>/dev/null
python tools/ncu_summary.py /tmp/${TAG}_tload.ncu-rep gpurun_out/${TAG}_ncu_tload.summary.csv 2>&1 | tail -2
echo "== gpu tests that used the plain-load kernel before (left $(left) s)"
L=$(left); [ $L -gt 40 ] && (time timeout $((L - 10)) python -m pytest tests -m gpu -q -x -k "memory_order or golden or 1d_r2c or samples or strided or smoke") 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_related.log
echo "== rest of the gpu suite (left $(left) s)"
L=$(left); [ $L -gt 60 ] && (time timeout $((L - 10)) python -m pytest tests -m gpu -q -x -k "not (tensor_load or memory_order or golden or 1d_r2c or samples or strided or smoke)") 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_rest.log
echo "== done (left $(left) s)"
