#!/bin/bash
# Fourth GPU call of round 2 (1 GPU): new smoother defaults (L2 kernel on axis 1, systolic on axes 2/3), predict_fast_kernel,
# every bench workload, launch list + captures for profiles/.  gpurun --timeout 2400 -- 'bash tools/r02d_call.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
S=$O/r02d_summary.txt
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
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02d_smoke.log 2>&1; echo "smoke rc $?: $(tail -1 $O/r02d_smoke.log)" | tee -a $S
timeout 1200 python -m pytest tests -q -m gpu --durations=6 -s > $O/r02d_pytest.log 2>&1; echo "pytest -m gpu rc $?" | tee -a $S
grep -E "passed|failed|rel-L2|panel\]|Error" $O/r02d_pytest.log | tail -8 | tee -a $S
echo "== headline bench, new defaults" | tee -a $S
timeout 400 python bench.py --steps 3 --warmup 3 > $O/r02d_bench_default.json 2> $O/r02d_bench_default.err; echo "rc $?: $(line $O/r02d_bench_default.json)" | tee -a $S
for v in "PST_PREDICT_FAST=0" "PST_PREDICT_FAST=2" "PST_TRI_L2=7 PST_TRI_SYS=0" "PST_TRI_L2=0" "PST_CG_DEVSCALARS=0"; do
    tag=$(echo "$v" | tr ' =' '__')
    env $v timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02d_bench_$tag.json 2> $O/r02d_bench_$tag.err
    echo "$tag rc $?: $(line $O/r02d_bench_$tag.json)" | tee -a $S
done
echo "== 500x512x512: smoother micro-benchmarks and the bench" | tee -a $S
timeout 100 tools/mb_tri_sys.bin benchonly 500 512 512 2>&1 | grep -E "^bench|CUDA" | grep -v in-place | tee -a $S
timeout 100 tools/mb_tri_l2.bin benchonly 500 512 512 2>&1 | grep -E "^bench|CUDA" | grep -v in-place | tee -a $S
timeout 100 tools/mb_tri_stream.bin benchonly 500 512 512 2>&1 | grep -E "^bench|CUDA" | tee -a $S
timeout 300 python bench.py --shape 500,512,512 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02d_bench_500.json 2> $O/r02d_bench_500.err; echo "500x512x512 rc $?: $(line $O/r02d_bench_500.json)" | tee -a $S
PST_TRI_L2=7 PST_TRI_SYS=0 timeout 300 python bench.py --shape 500,512,512 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02d_bench_500_l2.json 2> $O/r02d_bench_500_l2.err; echo "500x512x512 L2 everywhere rc $?: $(line $O/r02d_bench_500_l2.json)" | tee -a $S
echo "== the other BASELINE.json configurations" | tee -a $S
timeout 400 python bench.py --workload dip2d_somf2d --steps 5 --warmup 3 > $O/r02d_bench_dip2d.json 2> $O/r02d_bench_dip2d.err; echo "dip2d_somf2d 3000x860 rc $?: $(line $O/r02d_bench_dip2d.json)" | tee -a $S
timeout 400 python bench.py --workload dip2d_somf2d --shape 30000,1280 --steps 3 --warmup 3 --no-cpu-baseline > $O/r02d_bench_dip2d_big.json 2> $O/r02d_bench_dip2d_big.err; echo "dip2d_somf2d 30000x1280 rc $?: $(line $O/r02d_bench_dip2d_big.json)" | tee -a $S
timeout 400 python bench.py --workload somean3d --steps 5 --warmup 3 > $O/r02d_bench_somean3d.json 2> $O/r02d_bench_somean3d.err; echo "somean3d rc $?: $(line $O/r02d_bench_somean3d.json)" | tee -a $S
PST_PREDICT_FAST=0 timeout 300 python bench.py --workload somean3d --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02d_bench_somean3d_old.json 2> $O/r02d_bench_somean3d_old.err; echo "somean3d, first prediction kernel rc $?: $(line $O/r02d_bench_somean3d_old.json)" | tee -a $S
timeout 400 python bench.py --workload soint3d --steps 3 --warmup 3 > $O/r02d_bench_soint3d.json 2> $O/r02d_bench_soint3d.err; echo "soint3d rc $?: $(line $O/r02d_bench_soint3d.json)" | tee -a $S
timeout 600 python bench.py --workload sint3d --steps 2 --warmup 3 > $O/r02d_bench_sint3d.json 2> $O/r02d_bench_sint3d.err; echo "sint3d rc $?: $(line $O/r02d_bench_sint3d.json)" | tee -a $S
echo "== ncu: launch list of one headline step at 500x512x512; full captures of the prediction kernels at bench size" | tee -a $S
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02d_launches.csv python bench.py --shape 500,512,512 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/r02d_launches.log 2>&1; echo "launch list rc $?" | tee -a $S
cat > /tmp/spray_big.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import pyseistr_b200 as ps
from pyseistr_b200 import synth
n1, n2, n3 = 1000, 1024, 60
d = synth.cube(n1, n2, n3, seed=3)
di, dx = synth.smooth_dips(n1, n2, n3, seed=3)
ps.somf3dc(d, di, dx, 2, 2, 0.01, 2, verb=0, ctx=ps.default_context(0))
PY
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"predict_fast_kernel|slot_median" -c 6 -o $O/r02d_predict python /tmp/spray_big.py > $O/r02d_ncu_predict.log 2>&1; echo "ncu predict rc $?" | tee -a $S
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tri_sys_kernel|tri_l2_kernel|allpass_kernel" -s 30 -c 6 -o $O/r02d_tri python bench.py --shape 500,512,512 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/r02d_ncu_tri.log 2>&1; echo "ncu tri rc $?" | tee -a $S
