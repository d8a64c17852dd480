#!/bin/bash
# round 2 measurements on one B200: full GPU test suite, the N = 1 bench line, the other BASELINE configurations,
# config 5 (fracture) with and without split handling
mkdir -p gpurun_out
tag=${1:-r2z}
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${tag}_gputests.log
cat gpurun_out/${tag}_gputests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_asteroid1024.json 2> gpurun_out/${tag}_bench.err
tail -2 gpurun_out/${tag}_bench.err
for w in sphere64 sphere202 noisybox256 asteroid512; do
  timeout 600 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/${tag}_bench_$w.json 2>> gpurun_out/${tag}_bench.err
done
timeout 900 python tools/bench_fracture.py --steps 32 --synced-mesh > gpurun_out/${tag}_fracture_plain.json 2>> gpurun_out/${tag}_bench.err
timeout 900 python tools/bench_fracture.py --steps 32 --synced-mesh --split > gpurun_out/${tag}_fracture_split.json 2>> gpurun_out/${tag}_bench.err
timeout 900 python tools/bench_aux.py > gpurun_out/${tag}_aux.json 2>> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_aux.json
python - <<PY
import json
for w in ("asteroid1024","sphere64","sphere202","noisybox256","asteroid512"):
    try:
        d = json.load(open("gpurun_out/${tag}_bench_%s.json" % w))
        print(w, "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), "parity", d["parity"].get("digest_ok"), "launches/step", d["gpu_launches"]/d["steps"])
        if w == "asteroid1024": print("  ", {k: round(v,3) for k,v in d["kernel_ms_per_step"].items()}, d.get("fp32_pipe",{}).get("frac"), d["roofline"]["frac"])
    except Exception as e:
        print(w, "failed:", e)
for f in ("plain","split"):
    try:
        d = json.load(open("gpurun_out/${tag}_fracture_%s.json" % f))
        print("fracture", f, "ms/step", round(d["ms_per_step"],3), "absorb", round(d["absorb_ms"],3), "remesh", round(d["remesh_ms"],3), "split", d["split_ms"], "medians", d.get("median_active_absorb_ms"), d.get("median_active_remesh_ms"), d.get("median_active_split_ms"))
    except Exception as e:
        print("fracture", f, "failed:", e)
PY
