#!/bin/bash
# Runs on the GPU box (via gpurun), session 4: the whole GPU suite, the inertial-moment timing, the fracture bench with
# and without the inertial-property updater, the mutual absorption benchmark and the N=1 bench line.
tag=${1:-s4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_gputests.log
tail -4 gpurun_out/${tag}_gputests.log
timeout 90 python tools/bench_inertia.py > gpurun_out/${tag}_inertia_bench.json 2> gpurun_out/${tag}_inertia_bench.err
cat gpurun_out/${tag}_inertia_bench.json | cut -c1-420
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
timeout 150 python tools/bench_mutual.py --radius 200 > gpurun_out/${tag}_mutual_r200.json 2> gpurun_out/${tag}_mutual_r200.err
cat gpurun_out/${tag}_mutual_r200.json | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "roofline frac", d["roofline"]["frac"], "launches", d.get("gpu_launches"))
except Exception as e:
    print("bench failed:", e)
PY
tail -3 gpurun_out/${tag}_bench.err
