#!/bin/bash
# communicator tests on the box's GPUs + slab tests + multi-GPU bench lines
mkdir -p gpurun_out
tag=${1:-r2c}
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_comm.py tests/test_gpu_slabs.py -x -q 2>&1 | tail -30 > gpurun_out/${tag}_comm_tests.log
tail -12 gpurun_out/${tag}_comm_tests.log
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
    tail -5 gpurun_out/${tag}_bench_n$n.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_n$n.json"))
    print("N=$n ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), "parity", d["parity"]["digest_ok"], d["details"]["rank0_exchange"].get("phase_ms_per_rank"))
except Exception as e:
    print("N=$n bench failed:", e)
PY
  fi
done
