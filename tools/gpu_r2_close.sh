#!/bin/bash
# round 2, last minute of box time: smoke() and the mixed-radix / kernel-size golden gpu tests on the final library
TAG=${TAG:-r02w}
mkdir -p gpurun_out
echo "== smoke"; timeout 25 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
echo "== gpu tests"; timeout 38 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "smooth_lengths or golden_vectors_at_kernel_sizes or 768_cubed" 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_final.log
