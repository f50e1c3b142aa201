#!/bin/bash
# SM split of the overlapped groups (P3DFFT_B200_OVERLAP_XSMS = SMs of the exchange stage) on N GPUs, bench without the e2e / cpu / parity legs
N=${N:-2}
mkdir -p gpurun_out
for XS in ${XSMS_LIST:-40 56 64 74}; do
P3DFFT_B200_OVERLAP_XSMS=$XS timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
   bench.py --gpus $N --no-e2e --no-cpu --no-parity --steps ${STEPS:-10} ${BENCH_ARGS:-} > gpurun_out/xsms_${N}_$XS.json 2> gpurun_out/xsms_${N}_$XS.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/xsms_${N}_$XS.json") if l.startswith("{")][-1])
    st = " ".join(f"{s['stage']}:{s['ms']:.2f}" + ("*" if s.get("overlap_group") else "") for s in d["roofline"]["stages"])
    print(f"N=$N xsms=$XS ${BENCH_ARGS:-}: {d['ms_per_step']:.3f} ms {d['value']:.0f} GF | {st}")
except Exception as ex:
    print("N=$N xsms=$XS FAILED", ex)
PY
done | tee -a gpurun_out/r02m_xsms_sweep_${N}gpu.txt
