#!/bin/bash
# Runs on the GPU box (via gpurun), session 4: the whole GPU suite, the fracture bench with and without the
# inertial-property updater, one full ncu capture of k_moments_non_uniform, and the N=1 bench line.
tag=${1:-s4}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_gputests.log
tail -4 gpurun_out/${tag}_gputests.log
timeout 200 python tools/bench_fracture.py --steps 32 > gpurun_out/${tag}_fracture_plain.json 2> gpurun_out/${tag}_fracture_plain.err
timeout 200 python tools/bench_fracture.py --steps 32 --inertial > gpurun_out/${tag}_fracture_inertial.json 2> gpurun_out/${tag}_fracture_inertial.err
python - <<PY
import json
for n in ("plain", "inertial"):
    try:
        d = json.load(open("gpurun_out/${tag}_fracture_%s.json" % n))
        print(n, "ms/step", round(d["ms_per_step"], 3), "absorb", round(d["absorb_ms"], 3), "remesh", round(d["remesh_ms"], 3), d.get("inertial"))
    except Exception as e:
        print(n, "failed:", e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_moments_non_uniform -c 1 -f -o gpurun_out/${tag}_k_moments_non_uniform \
    python tools/bench_inertia.py --reps 1 > gpurun_out/${tag}_ncu_k_moments.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"], "roofline", d["roofline"], "cpu", d["cpu_baseline"])
except Exception as e:
    print("bench failed:", e)
PY
tail -3 gpurun_out/${tag}_bench.err
ls -la gpurun_out | tail -6
