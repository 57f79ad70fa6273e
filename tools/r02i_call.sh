#!/bin/bash
# Sixth single-GPU call of round 2: full GPU suite after (a) the device-copy cache was removed, (b) every sum of the dip
# solve became canonical (partition-independent), (c) the warp-per-trace prediction kernel and the block-per-trace pwd3
# kernels; benches of the workloads those touch.
set -u
mkdir -p gpurun_out
O=gpurun_out
S=$O/r02i_summary.txt
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print(round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 1), "| roofline", r.get("kernel_class"), r.get("bound"), "frac", round(r.get("frac", 0), 4),
          "| e2e", round(d.get("e2e", {}).get("value", 0), 2), "| cpu", d.get("cpu_baseline", {}).get("value"),
          {k: round(v["ms_per_step"], 1) for k, v in r.get("classes", {}).items()})
except Exception as e:
    print("no JSON line:", e)
PY
}
echo "== smoke + full GPU suite" | tee $S
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02i_smoke.log 2>&1; echo "smoke rc $?: $(tail -1 $O/r02i_smoke.log)" | tee -a $S
timeout 1200 python -m pytest tests -q -m gpu --durations=6 -s > $O/r02i_pytest.log 2>&1; echo "pytest -m gpu rc $?" | tee -a $S
grep -E "passed|failed|rel-L2|panel\]|Error" $O/r02i_pytest.log | tail -10 | tee -a $S
echo "== spray / interpolation tests with the warp-per-trace prediction kernel forced on (1) and off (0)" | tee -a $S
for v in 1 0; do
    PST_PREDICT_WARP=$v timeout 600 python -m pytest tests -q -m gpu -k "spray or somf or somean or svmf or paint or sint or soint or shim or golden" > $O/r02i_pytest_warp$v.log 2>&1; echo "PST_PREDICT_WARP=$v rc $?: $(tail -1 $O/r02i_pytest_warp$v.log)" | tee -a $S
done
echo "== benches (JSON lines kept)" | tee -a $S
timeout 500 python bench.py --steps 5 --warmup 3 > $O/r02i_bench_dip3d_somf3d_n1.json 2> $O/r02i_bench_dip3d_somf3d_n1.err; echo "headline rc $?: $(line $O/r02i_bench_dip3d_somf3d_n1.json)" | tee -a $S
timeout 300 python bench.py --shape 500,512,512 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02i_bench_dip3d_somf3d_500_n1.json 2> /dev/null; echo "500x512x512 rc $?: $(line $O/r02i_bench_dip3d_somf3d_500_n1.json)" | tee -a $S
timeout 400 python bench.py --workload dip2d_somf2d --steps 5 --warmup 3 > $O/r02i_bench_dip2d_somf2d_n1.json 2> /dev/null; echo "dip2d_somf2d 3000x860 rc $?: $(line $O/r02i_bench_dip2d_somf2d_n1.json)" | tee -a $S
PST_PREDICT_WARP=0 timeout 400 python bench.py --workload dip2d_somf2d --steps 5 --warmup 3 --no-cpu-baseline > $O/r02i_bench_dip2d_somf2d_nowarp.json 2> /dev/null; echo "dip2d_somf2d 3000x860, PST_PREDICT_WARP=0 rc $?: $(line $O/r02i_bench_dip2d_somf2d_nowarp.json)" | tee -a $S
timeout 400 python bench.py --workload dip2d_somf2d --shape 30000,1280 --steps 3 --warmup 3 --no-cpu-baseline > $O/r02i_bench_dip2d_somf2d_30000x1280_n1.json 2> /dev/null; echo "dip2d_somf2d 30000x1280 rc $?: $(line $O/r02i_bench_dip2d_somf2d_30000x1280_n1.json)" | tee -a $S
timeout 400 python bench.py --workload somean3d --steps 5 --warmup 3 > $O/r02i_bench_somean3d_n1.json 2> /dev/null; echo "somean3d rc $?: $(line $O/r02i_bench_somean3d_n1.json)" | tee -a $S
timeout 400 python bench.py --workload soint3d --steps 3 --warmup 3 > $O/r02i_bench_soint3d_n1.json 2> /dev/null; echo "soint3d rc $?: $(line $O/r02i_bench_soint3d_n1.json)" | tee -a $S
timeout 600 python bench.py --workload sint3d --steps 2 --warmup 3 > $O/r02i_bench_sint3d_n1.json 2> /dev/null; echo "sint3d rc $?: $(line $O/r02i_bench_sint3d_n1.json)" | tee -a $S
du -sh $O | tee -a $S
