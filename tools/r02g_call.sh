#!/bin/bash
# Fifth single-GPU call of round 2: full GPU suite on the final defaults, every bench workload (JSON lines kept), launch list,
# ncu captures exported as CSV on the box (reports stay there: gpurun_out is limited to 64 MiB).
set -u
mkdir -p gpurun_out
O=gpurun_out
S=$O/r02g_summary.txt
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
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02g_smoke.log 2>&1; echo "smoke rc $?: $(tail -1 $O/r02g_smoke.log)" | tee -a $S
timeout 1200 python -m pytest tests -q -m gpu --durations=6 -s > $O/r02g_pytest.log 2>&1; echo "pytest -m gpu rc $?" | tee -a $S
grep -E "passed|failed|rel-L2|panel\]|resident\]|Error" $O/r02g_pytest.log | tail -10 | tee -a $S
echo "== benches (JSON lines kept)" | tee -a $S
timeout 500 python bench.py --steps 5 --warmup 3 > $O/r02g_bench_dip3d_somf3d_n1.json 2> $O/r02g_bench_dip3d_somf3d_n1.err; echo "headline rc $?: $(line $O/r02g_bench_dip3d_somf3d_n1.json)" | tee -a $S
PST_RESIDENT=0 timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02g_bench_noresident.json 2> $O/r02g_bench_noresident.err; echo "headline, PST_RESIDENT=0 rc $?: $(line $O/r02g_bench_noresident.json)" | tee -a $S
timeout 300 python bench.py --shape 500,512,512 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02g_bench_dip3d_somf3d_500_n1.json 2> /dev/null; echo "500x512x512 rc $?: $(line $O/r02g_bench_dip3d_somf3d_500_n1.json)" | tee -a $S
timeout 400 python bench.py --workload dip2d_somf2d --steps 5 --warmup 3 > $O/r02g_bench_dip2d_somf2d_n1.json 2> /dev/null; echo "dip2d_somf2d 3000x860 rc $?: $(line $O/r02g_bench_dip2d_somf2d_n1.json)" | tee -a $S
timeout 400 python bench.py --workload dip2d_somf2d --shape 30000,1280 --steps 3 --warmup 3 --no-cpu-baseline > $O/r02g_bench_dip2d_somf2d_30000x1280_n1.json 2> /dev/null; echo "dip2d_somf2d 30000x1280 rc $?: $(line $O/r02g_bench_dip2d_somf2d_30000x1280_n1.json)" | tee -a $S
timeout 400 python bench.py --workload somean3d --steps 5 --warmup 3 > $O/r02g_bench_somean3d_n1.json 2> /dev/null; echo "somean3d rc $?: $(line $O/r02g_bench_somean3d_n1.json)" | tee -a $S
timeout 400 python bench.py --workload soint3d --steps 3 --warmup 3 > $O/r02g_bench_soint3d_n1.json 2> /dev/null; echo "soint3d rc $?: $(line $O/r02g_bench_soint3d_n1.json)" | tee -a $S
timeout 600 python bench.py --workload sint3d --steps 2 --warmup 3 > $O/r02g_bench_sint3d_n1.json 2> /dev/null; echo "sint3d rc $?: $(line $O/r02g_bench_sint3d_n1.json)" | tee -a $S
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02g_bench_reference_arm.json 2> /dev/null; echo "reference arm rc $?: $(head -c 400 $O/r02g_bench_reference_arm.json)" | tee -a $S
echo "== ncu: launch list of one headline step at 500x512x512" | tee -a $S
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02g_launches.csv python bench.py --shape 500,512,512 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/r02g_launches.log 2>&1; echo "launch list rc $?, $(wc -l < $O/r02g_launches.csv) lines" | tee -a $S
python - <<'PY' | tee -a $S
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02g_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
t = collections.Counter(); n = collections.Counter()
for r in rows[1:]:
    k = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    t[k] += float(r[vi].replace(",", "")); n[k] += 1
tot = sum(t.values())
for k, v in t.most_common(16):
    print(f"  {k:58s} {n[k]:6d} launches {v/1e6:9.2f} ms {100*v/tot:5.1f} %")
print(f"  total {tot/1e6:.1f} ms in {sum(n.values())} launches")
PY
gzip -f $O/r02g_launches.csv
echo "== ncu --set full: prediction kernels at bench size, smoothing + stencil kernels in the bench; exported as CSV" | tee -a $S
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
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"predict_fast_kernel|slot_median" -c 5 -o /tmp/r02g_predict python /tmp/spray_big.py > $O/r02g_ncu_predict.log 2>&1; echo "ncu predict rc $?" | tee -a $S
ncu -i /tmp/r02g_predict.ncu-rep --page raw --csv > $O/r02g_predict_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tri_sys_kernel|tri_l2_kernel|allpass_kernel" -s 30 -c 6 -o /tmp/r02g_tri python bench.py --shape 500,512,512 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/r02g_ncu_tri.log 2>&1; echo "ncu tri rc $?" | tee -a $S
ncu -i /tmp/r02g_tri.ncu-rep --page raw --csv > $O/r02g_tri_raw.csv 2>/dev/null
ls -la $O | tail -30 >> $S
du -sh $O | tee -a $S
