#!/bin/bash
# Single-GPU call: the kernels of the multi-GPU axis-3 pass on one GPU -- bit-exactness of the register kernels for every
# rank role (pst_selftest_axis3_slabs), time per pass of the tile / register kernels on the slab of 8 ranks.
set -u
mkdir -p gpurun_out
O=gpurun_out
S=$O/r02k_summary.txt
echo "== GPU tests: smoothing (incl. the multi-rank kernels on one GPU)" | tee $S
timeout 600 python -m pytest tests -q -m gpu -k "smooth or axis3" > $O/r02k_pytest.log 2>&1; echo "pytest rc $?: $(tail -1 $O/r02k_pytest.log)" | tee -a $S
grep -E "FAILED|Error" $O/r02k_pytest.log | head -20 | tee -a $S
echo "== time per axis-3 pass, 1000x1024x128 (the slab of 8 ranks), r3 = 5" | tee -a $S
timeout 600 python tools/mb_axis3_tiles.py 2>&1 | tee -a $S
echo "== ncu: the register kernels (interior-rank role, timing mode)" | tee -a $S
PST_TRI3_SOLO=2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__average_warp_latency_issue_stalled_long_scoreboard.pct --clock-control none -k regex:"tri3_reg" -c 4 --csv --log-file $O/r02k_ncu_reg.csv python tools/mb_axis3_tiles.py 2 1000 1024 128 5 2 > $O/r02k_ncu_reg.log 2>&1; echo "ncu rc $?" | tee -a $S
python - <<'PY' | tee -a $S
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02k_ncu_reg.csv")) if len(r) > 10]
if rows:
    h = rows[0]; ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    out = {}
    for r in rows[1:]:
        out.setdefault((r[ii], r[ki][:40]), {})[r[mi]] = r[vi]
    for k, v in out.items():
        print(" ", k, v)
PY
