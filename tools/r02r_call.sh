#!/bin/bash
# Headline cube, device-resident only, pieces of 8192 elements (PST_RED_CH=8192) against the 32768 of the last full run.
set -u
mkdir -p gpurun_out
PST_RED_CH=8192 timeout 75 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err
python - <<'PY' | tee gpurun_out/r02r_summary.txt
import json
d = json.loads(open("gpurun_out/r02r_bench.json").read().strip().splitlines()[-1])
print(round(d["value"], 2), "Mvoxels/s", round(d["ms_per_step"], 1), "ms/step", {k: round(v["ms_per_step"], 1) for k, v in d["roofline"]["classes"].items()})
PY
