#!/bin/bash
# multi-GPU bench lines only: tools/gpu_scale2.sh TAG N...
mkdir -p gpurun_out
tag=$1; shift
for n in "$@"; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
    grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" gpurun_out/${tag}_bench_n$n.err | tail -5
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_n$n.json"))
    print("N=$n ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), "parity", d["parity"]["digest_ok"], d["details"]["rank0_exchange"].get("phase_ms_per_rank"))
except Exception as e:
    print("N=$n bench failed:", e)
PY
done
