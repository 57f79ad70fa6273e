#!/bin/bash
# Second GPU call of round 2: the L2-resident checkpoint + recompute smoother (pst_tri_l2.cu), the rest of the GPU suite,
# the prediction-kernel capture.  gpurun --timeout 2400 -- 'bash tools/r02b_call.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
S=$O/r02b_summary.txt
echo "== L2 smoother: bit-exactness + micro-benchmark (4 warps per SM, all L2 hints)" | tee $S
timeout 300 tools/mb_tri_l2.bin bench > $O/r02b_mb_l2.log 2>&1; echo "mb_tri_l2 rc $?" | tee -a $S
grep -E "correctness|^bench|MISMATCH|CUDA|not eligible" $O/r02b_mb_l2.log | tail -14 | tee -a $S
if grep -q "correctness: 0 failing" $O/r02b_mb_l2.log; then
  for cfg in "PST_TRI_L2_WARPS=3" "PST_TRI_L2_WARPS=5" "PST_TRI_L2_WARPS=6" "PST_TRI_L2_WARPS=2" "PST_TRI_L2_HINTS=0" "PST_TRI_L2_HINTS=1" "PST_TRI_L2_HINTS=5" "PST_TRI_L2_WARPS=4 PST_TRI_L2_SLOTS=4" "PST_TRI_L2_WARPS=6 PST_TRI_L2_HINTS=0"; do
    echo "-- $cfg" | tee -a $S
    env $cfg timeout 120 tools/mb_tri_l2.bin benchonly 2>&1 | grep -E "^bench|CUDA" | grep -v in-place | tee -a $S
  done
  echo "-- 500x512x512" | tee -a $S
  timeout 120 tools/mb_tri_l2.bin benchonly 500 512 512 2>&1 | grep -E "^bench|CUDA" | grep -v in-place | tee -a $S
  echo "== bench.py with the L2 smoother on the strided axes (6) / on every axis (7)" | tee -a $S
  for v in "PST_TRI_L2=6" "PST_TRI_L2=7" "PST_TRI_L2=6 PST_TRI_SYS=0"; do
    tag=$(echo "$v" | tr ' =' '__')
    env $v timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02b_bench_$tag.json 2> $O/r02b_bench_$tag.err
    echo "$tag rc $?: $(python - <<PY
import json
try:
    d = json.loads(open("$O/r02b_bench_$tag.json").read().strip().splitlines()[-1])
    print(d["value"], d["unit"], "ms/step", d["ms_per_step"], "tri frac", d.get("roofline", {}).get("frac"), "e2e", d.get("e2e", {}).get("value"))
except Exception as e:
    print("no JSON line:", e)
PY
)" | tee -a $S
  done
  echo "== GPU tests that smooth, L2 smoother on every axis" | tee -a $S
  PST_TRI_L2=7 timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "smooth or dip3d or divne or dip2d" > $O/r02b_pytest_l2.log 2>&1; echo "pytest (PST_TRI_L2=7) rc $?" | tee -a $S
  tail -3 $O/r02b_pytest_l2.log | tee -a $S
  echo "== ncu: the L2 smoother (axis 2 then axis 3 launches of the micro-benchmark)" | tee -a $S
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:tri_l2_kernel -s 10 -c 2 -o $O/r02b_tri_l2 tools/mb_tri_l2.bin benchonly 1000 1024 256 > $O/r02b_ncu_l2.log 2>&1; echo "ncu tri_l2 rc $?" | tee -a $S
fi
echo "== the full GPU suite (default path)" | tee -a $S
timeout 1200 python -m pytest tests -q -m gpu --durations=8 -s > $O/r02b_pytest.log 2>&1; echo "pytest -m gpu rc $?" | tee -a $S
grep -E "passed|failed|rel-L2|panel\]" $O/r02b_pytest.log | tail -8 | tee -a $S
echo "== ncu: the prediction kernels + slot median" | tee -a $S
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"predict_kernel|slot_median" -c 12 -o $O/r02b_predict python tools/spray_once.py > $O/r02b_ncu_predict.log 2>&1; echo "ncu predict rc $?" | tee -a $S
