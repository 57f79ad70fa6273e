#!/bin/bash
# 500x512x512 headline bench (device-resident only) after the CG head kernels moved to the fixed-piece mapping.
set -u
mkdir -p gpurun_out
timeout 80 python bench.py --shape 500,512,512 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02q_bench_500.json 2> gpurun_out/r02q_bench_500.err
python - <<'PY' | tee gpurun_out/r02q_summary.txt
import json
d = json.loads(open("gpurun_out/r02q_bench_500.json").read().strip().splitlines()[-1])
print(round(d["value"], 2), "Mvoxels/s", round(d["ms_per_step"], 1), "ms/step", {k: round(v["ms_per_step"], 1) for k, v in d["roofline"]["classes"].items()})
PY
