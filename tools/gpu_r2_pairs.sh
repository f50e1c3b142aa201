#!/bin/bash
# round-2 session: persistent pair kernels -- parity (ranks may share GPUs), then step time with and without them at N = $1
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpus.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
( time timeout 1500 python -m pytest tests/test_multirank.py -m gpu -x -q -k "overlapped_pairs" 2>&1 | tail -15 ) > gpurun_out/r02_pytest_pairs_$N.log 2>&1
cat gpurun_out/r02_pytest_pairs_$N.log
fi
for SY in ${SYNCS:-1 0}; do
  export P3DFFT_B200_PAIR_SYNC=$SY P3DFFT_B200_PEER_TIMEOUT_S=60
  for XS in ${XSMS:-default}; do
    if [ "$XS" = default ]; then unset P3DFFT_B200_OVERLAP_XSMS; else export P3DFFT_B200_OVERLAP_XSMS=$XS; fi
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
       bench.py --gpus $N --no-cpu --no-e2e --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/r02_pairs_${N}_sync${SY}_$XS.json 2> gpurun_out/r02_pairs_${N}_sync${SY}_$XS.err
    grep "pair-sync trace" gpurun_out/r02_pairs_${N}_sync${SY}_$XS.err | tail -2 | cut -c1-400 | tee -a gpurun_out/r02_pairs_$N.txt
    python - $N $SY $XS <<'PY' | tee -a gpurun_out/r02_pairs_$N.txt
import json, sys
n, sy, xs = sys.argv[1:4]
f = f"gpurun_out/r02_pairs_{n}_sync{sy}_{xs}"
try:
    d = json.loads(open(f + ".json").read().strip().splitlines()[-1])
    st = " ".join(f"{s['stage']}:{s['ms']:.2f}" + (f"({s['nvlink_gbs']:.0f})" if s.get('nvlink_gbs') else "") + ("*" if s.get("overlapped_with") else "") for s in d["roofline"]["stages"])
    print(f"N={n} sync={sy} xsms={xs} ms={d['ms_per_step']:.3f} gflops={d['value']:.0f} | {st}")
except Exception as e:
    print(f"N={n} sync={sy} xsms={xs} FAILED {e}", open(f + ".err").read()[-1500:])
PY
  done
done
