#!/bin/bash
# Runs on the GPU box (gpurun --gpus N): multi-GPU parity tests, then bench.py under torchrun on N GPUs.
n=${1:-2}; tag=${2:-scale}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "slab or distributed or nccl" 2>&1 | tail -4 > gpurun_out/${tag}_n${n}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
tail -3 gpurun_out/${tag}_n${n}_tests.log
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_n${n}.json"))
    print("n", d["n_gpus"], "ms/step", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3))
    print(d["config"]["slab_planes"], d["config"]["rank0_exchange"].get("phase_ms_per_rank"))
except Exception as e:
    print("bench failed:", e)
PY
tail -3 gpurun_out/${tag}_bench_n${n}.err
