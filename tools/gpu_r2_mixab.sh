#!/bin/bash
# round 2, last seconds of box time: the mixed-radix kernel's radix-Q step as register butterflies (library build B,
# p3dfft.3_b200/lib_b) against the by-definition form (build A, p3dfft.3_b200/lib) -- parity of B, then both timed.
# How the two builds were made (lib_b travels gzipped: the snapshot limit is 512 MiB):
#   make -j LIBDIR=p3dfft.3_b200/lib_b EXTRA_NVFLAGS='-DP3B_MIX_BFLY_MINQ_F64=3 -DP3B_MIX_BFLY_MINQ_F32=3' && gzip -1 -k p3dfft.3_b200/lib_b/libp3dfft.3.so
#   make -j EXTRA_NVFLAGS='-DP3B_MIX_BFLY_MINQ_F64=99 -DP3B_MIX_BFLY_MINQ_F32=99'
TAG=${TAG:-r02v}
mkdir -p gpurun_out
T0=$(date +%s)
left() { echo $(( ${BUDGET_S:-110} - ($(date +%s) - T0) )); }
gunzip -k -f p3dfft.3_b200/lib_b/libp3dfft.3.so.gz
B=$PWD/p3dfft.3_b200/lib_b/libp3dfft.3.so
echo "== build B: gpu tests of the mixed-radix kernel"
(time P3DFFT_B200_LIB=$B timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "smooth_lengths or golden_vectors_at_kernel_sizes or 768_cubed") 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_bfly.log
echo "== build B timing (left $(left) s)"; L=$(left); [ $L -gt 20 ] && P3DFFT_B200_LIB=$B timeout $((L > 45 ? 45 : L - 5)) python tools/gpu_configs.py MIXAB 2>&1 | tee gpurun_out/${TAG}_mixab_butterflies.txt | cut -c1-330
echo "== build A timing (left $(left) s)"; L=$(left); [ $L -gt 20 ] && timeout $((L - 5)) python tools/gpu_configs.py MIXAB 2>&1 | tee gpurun_out/${TAG}_mixab_by_definition.txt | cut -c1-330
echo "== done (left $(left) s)"
