#!/bin/bash
# First GPU call of round 2: measure the two smoothers that were written (and verified on the host only) after round 1's
# GPU minutes were spent, then the full GPU suite and the captures round 1 lacked.  Run from the repo root on a B200:
#   gpurun --timeout 2400 -- 'bash tools/r02_first_call.sh'
# Everything lands in gpurun_out/r02a_*.  Each step has its own timeout: a protocol bug must not eat the call.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== systolic smoother: bit-exactness + micro-benchmark" | tee $O/r02a_summary.txt
timeout 300 tools/mb_tri_sys.bin bench > $O/r02a_mb_tri_sys.log 2>&1; echo "mb_tri_sys rc $?" | tee -a $O/r02a_summary.txt
grep -E "correctness|bench|MISMATCH|CUDA" $O/r02a_mb_tri_sys.log | tail -12 | tee -a $O/r02a_summary.txt
echo "== systolic smoother, outputs stored from inside the backward chain (PST_TRI_SYS_ILS=1)" | tee -a $O/r02a_summary.txt
PST_TRI_SYS_ILS=1 timeout 200 tools/mb_tri_sys.bin bench > $O/r02a_mb_tri_sys_ils.log 2>&1; echo "mb_tri_sys ILS rc $?" | tee -a $O/r02a_summary.txt
grep -E "correctness|bench|MISMATCH|CUDA" $O/r02a_mb_tri_sys_ils.log | tail -8 | tee -a $O/r02a_summary.txt
echo "== systolic smoother, next tile's t pre-built in place (PST_TRI_SYS_PRE=1)" | tee -a $O/r02a_summary.txt
PST_TRI_SYS_PRE=1 timeout 200 tools/mb_tri_sys.bin bench > $O/r02a_mb_tri_sys_pre.log 2>&1; echo "mb_tri_sys PRE rc $?" | tee -a $O/r02a_summary.txt
grep -E "correctness|bench|MISMATCH|CUDA" $O/r02a_mb_tri_sys_pre.log | tail -8 | tee -a $O/r02a_summary.txt
echo "== checkpoint + recompute smoother" | tee -a $O/r02a_summary.txt
timeout 200 tools/mb_tri_rc.bin bench > $O/r02a_mb_tri_rc.log 2>&1; echo "mb_tri_rc rc $?" | tee -a $O/r02a_summary.txt
grep -E "correctness|bench|MISMATCH|CUDA" $O/r02a_mb_tri_rc.log | tail -14 | tee -a $O/r02a_summary.txt
echo "== streaming smoother (the default) for reference" | tee -a $O/r02a_summary.txt
timeout 200 tools/mb_tri_stream.bin benchonly > $O/r02a_mb_tri_stream.log 2>&1
grep -E "^bench " $O/r02a_mb_tri_stream.log | tee -a $O/r02a_summary.txt
echo "== bench.py: default, PST_TRI_SYS=1 (+PRE), PST_TRI_SYS=2, PST_TRI_RC=1" | tee -a $O/r02a_summary.txt
for v in "" "PST_TRI_SYS=1" "PST_TRI_SYS=1 PST_TRI_SYS_PRE=1" "PST_TRI_SYS=2" "PST_TRI_RC=1"; do
    tag=$(echo "${v:-default}" | tr ' =' '__')
    env $v timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02a_bench_$tag.json 2> $O/r02a_bench_$tag.err
    echo "$tag rc $?: $(python - <<PY
import json
try:
    d = json.loads(open("$O/r02a_bench_$tag.json").read().strip().splitlines()[-1])
    print(d["value"], d["unit"], "ms/step", d["ms_per_step"], "tri frac", d.get("roofline", {}).get("frac"))
except Exception as e:
    print("no JSON line:", e)
PY
)" | tee -a $O/r02a_summary.txt
done
echo "== smoke + the full GPU suite (default path)" | tee -a $O/r02a_summary.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02a_smoke.log 2>&1; echo "smoke rc $?" | tee -a $O/r02a_summary.txt
tail -2 $O/r02a_smoke.log | tee -a $O/r02a_summary.txt
timeout 900 python -m pytest tests -x -q -m gpu --durations=8 > $O/r02a_pytest.log 2>&1; echo "pytest -m gpu rc $?" | tee -a $O/r02a_summary.txt
tail -14 $O/r02a_pytest.log | tee -a $O/r02a_summary.txt
echo "== parity of the dip path with the systolic smoother switched on (GPU tests that smooth)" | tee -a $O/r02a_summary.txt
PST_TRI_SYS=1 timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "smooth or dip3d or divne" > $O/r02a_pytest_sys.log 2>&1; echo "pytest (PST_TRI_SYS=1) rc $?" | tee -a $O/r02a_summary.txt
tail -3 $O/r02a_pytest_sys.log | tee -a $O/r02a_summary.txt
echo "== ncu: the systolic kernel, the prediction kernels + slot median" | tee -a $O/r02a_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tri_sys_kernel -s 4 -c 3 -o $O/r02a_tri_sys tools/mb_tri_sys.bin benchonly 1000 1024 256 > $O/r02a_ncu.log 2>&1; echo "ncu tri_sys rc $?" | tee -a $O/r02a_summary.txt
cat > /tmp/spray_once.py <<'PY'
import numpy as np, pyseistr_b200 as ps
from pyseistr_b200 import synth
n1, n2, n3 = 1000, 256, 96
d = synth.cube(n1, n2, n3, seed=3)
di, dx = synth.smooth_dips(n1, n2, n3, seed=3)
ctx = ps.default_context(0)
ps.somf3dc(d, di, dx, 2, 2, 0.01, 2, verb=0, ctx=ctx)
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"predict_kernel|slot_median" -c 12 -o $O/r02a_predict python /tmp/spray_once.py > $O/r02a_ncu_predict.log 2>&1; echo "ncu predict rc $?" | tee -a $O/r02a_summary.txt
