#!/bin/bash
# Third GPU call of round 2: the reworked L2-resident smoother (no divisions in the ring bookkeeping, direct coalesced stores
# on the strided axes, shared products, up to 8 warps per SM).  gpurun --timeout 1500 -- 'bash tools/r02c_call.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
S=$O/r02c_summary.txt
echo "== L2 smoother v2: bit-exactness + micro-benchmark (defaults: 8 warps, 4 slots)" | tee $S
timeout 300 tools/mb_tri_l2.bin bench > $O/r02c_mb_l2.log 2>&1; echo "mb_tri_l2 rc $?" | tee -a $S
grep -E "correctness|^bench|MISMATCH|CUDA|not eligible" $O/r02c_mb_l2.log | tail -14 | tee -a $S
if grep -q "correctness: 0 failing" $O/r02c_mb_l2.log; then
  for cfg in "PST_TRI_L2_WARPS=4" "PST_TRI_L2_WARPS=5" "PST_TRI_L2_WARPS=6" "PST_TRI_L2_WARPS=7" "PST_TRI_L2_WARPS=8 PST_TRI_L2_SLOTS=2" "PST_TRI_L2_WARPS=8 PST_TRI_L2_SLOTS=3" "PST_TRI_L2_WARPS=8 PST_TRI_L2_HINTS=0" "PST_TRI_L2_WARPS=6 PST_TRI_L2_SLOTS=6"; do
    echo "-- $cfg" | tee -a $S
    env $cfg timeout 120 tools/mb_tri_l2.bin benchonly 2>&1 | grep -E "^bench|CUDA" | grep -v in-place | tee -a $S
  done
  echo "-- 500x512x512 (defaults)" | tee -a $S
  timeout 120 tools/mb_tri_l2.bin benchonly 500 512 512 2>&1 | grep -E "^bench|CUDA" | grep -v in-place | tee -a $S
  echo "== ncu: DRAM traffic of the strided-axis kernel at 8 / 6 / 4 warps" | tee -a $S
  for w in 8 6 4; do
    PST_TRI_L2_WARPS=$w timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__cycles_active.avg,lts__t_sector_hit_rate.pct --clock-control none -k regex:tri_l2_kernel -s 2 -c 12 --csv --log-file $O/r02c_ncu_w$w.csv tools/mb_tri_l2.bin benchonly 1000 1024 256 > /dev/null 2>&1
    python - <<PY | tee -a $S
import csv
rows = [r for r in csv.reader(open("$O/r02c_ncu_w$w.csv")) if len(r) > 10]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
idi = hdr.index("ID")
agg = {}
for r in rows[1:]:
    agg.setdefault((r[idi], r[ki][:40]), {})[r[mi]] = float(r[vi].replace(",", ""))
seen = set()
for (i, k), m in agg.items():
    if k in seen: continue
    seen.add(k)
    print("  warps $w", k, {a: round(b, 3) for a, b in m.items()})
PY
  done
  echo "== bench.py with the L2 smoother on the strided axes (6) / on every axis (7)" | tee -a $S
  for v in "PST_TRI_L2=6" "PST_TRI_L2=7"; do
    tag=$(echo "$v" | tr ' =' '__')
    env $v timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02c_bench_$tag.json 2> $O/r02c_bench_$tag.err
    echo "$tag rc $?: $(python - <<PY
import json
try:
    d = json.loads(open("$O/r02c_bench_$tag.json").read().strip().splitlines()[-1])
    print(d["value"], d["unit"], "ms/step", d["ms_per_step"], "tri frac", d.get("roofline", {}).get("frac"), "e2e", d.get("e2e", {}).get("value"))
except Exception as e:
    print("no JSON line:", e)
PY
)" | tee -a $S
  done
  echo "== GPU tests that smooth, L2 smoother on every axis" | tee -a $S
  PST_TRI_L2=7 timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "smooth or dip3d or divne or dip2d" > $O/r02c_pytest_l2.log 2>&1; echo "pytest (PST_TRI_L2=7) rc $?" | tee -a $S
  tail -3 $O/r02c_pytest_l2.log | tee -a $S
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:tri_l2_kernel -s 6 -c 1 -o $O/r02c_tri_l2 tools/mb_tri_l2.bin benchonly 1000 1024 256 > $O/r02c_ncu_l2.log 2>&1; echo "ncu tri_l2 (strided, full) rc $?" | tee -a $S
fi
echo "== default path: bench with e2e, new tests (SVMF, sint2d, device-side CG scalars)" | tee -a $S
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02c_bench_default.json 2> $O/r02c_bench_default.err
python - <<PY | tee -a $S
import json
try:
    d = json.loads(open("$O/r02c_bench_default.json").read().strip().splitlines()[-1])
    print("default:", d["value"], "ms/step", d["ms_per_step"], "tri frac", d["roofline"]["frac"], "e2e", d.get("e2e", {}).get("value"), {k: round(v["ms_per_step"]) for k, v in d["roofline"]["classes"].items()})
except Exception as e:
    print("no JSON line:", e)
PY
timeout 600 python -m pytest tests -q -m gpu -k "not 30000 and not 200x128" > $O/r02c_pytest.log 2>&1; echo "pytest -m gpu (quick subset) rc $?" | tee -a $S
tail -5 $O/r02c_pytest.log | tee -a $S
