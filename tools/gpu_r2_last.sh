#!/bin/bash
# round 2, closing 1-GPU call (under 3 minutes of box time): smoke() of the final library, the gpu tests of the kernels added
# last (smooth lengths 9 x 2^k / 15 x 2^k / 64-point cores, tensor-load kernel, kernel-size golden vectors incl. strided user
# arrays), A/B timing of the new smooth lengths against Bluestein
TAG=${TAG:-r02u}
mkdir -p gpurun_out
T0=$(date +%s)
left() { echo $(( ${BUDGET_S:-165} - ($(date +%s) - T0) )); }
echo "== smoke"; timeout 50 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
echo "== gpu tests of the last kernels (left $(left) s)"
L=$(left); (time timeout $((L > 90 ? 90 : L - 5)) python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "smooth_lengths or tensor_load_kernel or golden_vectors_at_kernel_sizes or fastcore_kinds") 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_last.log
echo "== smooth lengths timing (left $(left) s)"
L=$(left); [ $L -gt 25 ] && timeout $((L - 5)) python tools/gpu_configs.py SMOOTH 2>&1 | tee gpurun_out/${TAG}_smooth_timing.txt | cut -c1-330
echo "== done (left $(left) s)"
